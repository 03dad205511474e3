"""Whole training step (forward + all active losses + backward) as one replayable CUDA graph.

The reference's step (utils/utils_init.py:199-239) issues ~8k ATen ops from Python; here the step is a fixed
sequence of yvb200 kernels, so it is captured once per batch shape and replayed with a single launch.  Inputs are
copied into static device buffers (the host->device boundary of utils/utils_init.py:201-204), the dropout RNG step
counter is advanced inside the graph, and every weight is re-split to bf16 planes inside the graph (weights change
every optimiser step).  Gradients land in the parameters' ``.grad`` (static storage owned by the graph pool).
"""
from typing import List, Optional

import torch

from . import fused, lib, ops, synth

_FLOAT_FIELDS = (1, 2, 4)          # image_features, image_locations, image_targets


class GradientExchange:
    """Data-parallel gradient averaging overlapped with backward.

    A post-accumulate hook on every parameter collects finished gradients into ~``bucket_mb`` buckets; a full
    bucket is all-reduced (NCCL ``AVG``, one grouped launch for all its tensors) on a communication stream that
    only waits for the backward work issued so far.  Under ``torch.cuda.graph`` capture the collectives become
    graph nodes on a parallel branch, so every replay overlaps them with the remaining backward kernels.  The
    reference gets the same effect from ``DistributedDataParallel`` bucket hooks (utils/distributed.py:97-99).
    """

    def __init__(self, model: torch.nn.Module, group=None, bucket_mb: float = 64.0):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.bucket_bytes = int(bucket_mb * 2 ** 20)
        self.pending: List[torch.Tensor] = []
        self.pending_bytes = 0
        self.device = next(model.parameters()).device
        self.comm = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.avg = dist.ReduceOp.AVG if dist.get_backend(group) == "nccl" else dist.ReduceOp.SUM
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in model.parameters()
                        if p.requires_grad]
        self.launched = 0
        self.defer = False          # True: hooks do nothing, ``reduce_all`` exchanges every gradient after backward

    def reduce_all(self, model: torch.nn.Module):
        """Bucketed exchange of all gradients (used when the collectives are not part of a captured graph)."""
        for p in model.parameters():
            if p.grad is not None:
                self.pending.append(p.grad)
                self.pending_bytes += p.grad.numel() * p.grad.element_size()
                if self.pending_bytes >= self.bucket_bytes:
                    self._flush()
        self.finish()

    def _on_grad(self, p: torch.Tensor):
        g = p.grad
        if g is None or self.defer:
            return
        self.pending.append(g)
        self.pending_bytes += g.numel() * g.element_size()
        if self.pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self):
        if not self.pending:
            return
        grads, self.pending, self.pending_bytes = self.pending, [], 0
        if self.comm is not None:
            cur = torch.cuda.current_stream(self.device)
            self.comm.wait_stream(cur)
            with torch.cuda.stream(self.comm):
                self._reduce(grads)
        else:
            self._reduce(grads)
        self.launched += 1

    def _reduce(self, grads):
        dist = self.dist
        if self.avg == dist.ReduceOp.AVG:           # NCCL: one grouped launch, averaging inside the collective
            with dist._coalescing_manager(group=self.group, device=self.device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=self.avg, group=self.group)
            return
        flat = torch.cat([g.reshape(-1) for g in grads])        # gloo (CPU tests): flatten, sum, scatter back
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(float(self.world))
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def finish(self):
        """Reduce the last partial bucket and make the current stream wait for all communication."""
        self._flush()
        if self.comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)

    def remove(self):
        for h in self.handles:
            h.remove()


class GraphedStep:
    def __init__(self, model: torch.nn.Module, args, example_batch: List[torch.Tensor], use_graph: bool = True,
                 refresh_weights_each_step: bool = True, warmup: int = 2, exchange: Optional[GradientExchange] = None):
        self.model, self.args = model, args
        self.exchange = exchange
        # NCCL collectives inside the captured graph overlap the exchange with backward; opt-in because a capture
        # with live communicator threads needs thread-local capture mode (YVB200_CAPTURE_NCCL=1).  Default: the
        # bucketed exchange runs right after the replay on the communication stream.
        import os as _os
        self.capture_exchange = exchange is not None and _os.environ.get("YVB200_CAPTURE_NCCL", "0") == "1"
        if exchange is not None and not self.capture_exchange:
            exchange.defer = True
        self.device = next(model.parameters()).device
        self.rt = ops.rt(self.device)
        self.refresh = refresh_weights_each_step
        self.static = [t.to(self.device).clone() if torch.is_tensor(t) else t for t in example_batch]
        if not bool(self.static[13].all()):
            raise RuntimeError("GraphedStep needs batches without padded candidates (opt_mask all ones)")
        self.loss = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.use_graph = use_graph
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero_grads()
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if use_graph:
            self._zero_grads()
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            mode = "thread_local" if self.capture_exchange else "global"
            with torch.cuda.graph(self.graph, capture_error_mode=mode):
                self._body()
            self.launches_per_step = lib.launch_count() - n0
        else:
            n0 = lib.launch_count()
            self._zero_grads()
            self._body()
            self.launches_per_step = lib.launch_count() - n0

    def _zero_grads(self):
        for p in self.model.parameters():
            p.grad = None

    def _body(self):
        self.rt.advance_rng()
        if self.refresh:
            self.rt.arena.refresh_all(force=True)
        b = self.static
        co = b[11]
        inputs = (b[6].flatten(0, 1), b[1].flatten(0, 1), b[2].flatten(0, 1), b[10].flatten(0, 1), b[7].flatten(0, 1),
                  b[3].flatten(0, 1), co.reshape(-1, co.size(2), co.size(3)), b[9].flatten(0, 1), b[15])
        out = self.model(*inputs)
        ld = fused.step_losses(b, out, self.args, training=True, flat=True)
        tot = 0.0
        for k in ("vision", "language", "ranking"):
            if k in ld:
                tot = tot + ld[k]
        if "traj" in ld:
            tot = tot + self.args.traj_loss_scale * ld["traj"]
        tot.backward()
        if self.exchange is not None and self.capture_exchange:
            self.exchange.finish()
        self.loss = tot.detach()

    def load(self, batch: List[torch.Tensor]):
        """Copy a (pinned host or device) batch into the static buffers (async on the current stream)."""
        for dst, src in zip(self.static, batch):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.static if torch.is_tensor(t))

    def run(self) -> torch.Tensor:
        """One step on the data currently in the static buffers; returns the (device) total loss."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._zero_grads()
            self._body()
        if self.exchange is not None and not self.capture_exchange:
            self.exchange.reduce_all(self.model)
        return self.loss
