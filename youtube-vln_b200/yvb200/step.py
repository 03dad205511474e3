"""Whole training step (forward + all active losses + backward) as one replayable CUDA graph.

The reference's step (utils/utils_init.py:199-239) issues ~8k ATen ops from Python; here the step is a fixed
sequence of yvb200 kernels, so it is captured once per batch shape and replayed with a single launch.  Inputs are
copied into static device buffers (the host->device boundary of utils/utils_init.py:201-204), the dropout RNG step
counter is advanced inside the graph, and every weight is re-split to bf16 planes inside the graph (weights change
every optimiser step).  Gradients land in the parameters' ``.grad`` (static storage owned by the graph pool).
"""
import os
from typing import List, Optional

import torch

from . import fused, lib, ops

_FLOAT_FIELDS = (1, 2, 4)          # image_features, image_locations, image_targets


class GradientExchange:
    """Data-parallel gradient averaging overlapped with backward.

    A post-accumulate hook on every parameter collects finished gradients; every ``segment_mb`` of them closes a
    *segment*: the stream is joined with the helper streams and an **external** CUDA event is recorded (inside a
    captured step this is an event-record node of the graph).  After the step is launched, the communication stream
    waits for each segment's event in turn and all-reduces that segment (NCCL ``AVG``, one grouped launch per
    segment) while the rest of the backward graph is still executing.  NCCL itself stays outside the graph.
    The reference gets the same overlap from ``DistributedDataParallel`` bucket hooks (utils/distributed.py:97-99).
    """

    def __init__(self, model: torch.nn.Module, group=None, segment_mb: float = None, overlap: bool = True):
        import os
        import torch.distributed as dist
        if segment_mb is None:
            segment_mb = float(os.environ.get("YVB200_SEGMENT_MB", "320"))
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.segment_bytes = int(segment_mb * 2 ** 20)
        self.overlap = overlap
        self.device = next(model.parameters()).device
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.nccl = dist.get_backend(group) == "nccl"
        # YVB200_EXCHANGE_DTYPE=bf16 (opt-in): half the bytes over NVLink, at bf16 rounding (2^-9 relative) of every
        # averaged gradient element -- NOT the reference's fp32 DistributedDataParallel arithmetic
        self.payload = os.environ.get("YVB200_EXCHANGE_DTYPE", "fp32")
        if self.payload not in ("fp32", "bf16"):
            raise RuntimeError("YVB200_EXCHANGE_DTYPE must be fp32 or bf16")
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in model.parameters()
                        if p.requires_grad]
        self.recording = False
        self.segments = []          # [(event or None, [grads])] of the step being issued / captured
        self.pending: List[torch.Tensor] = []
        self.pending_params: List[torch.Tensor] = []
        self.segment_params: List[List[torch.Tensor]] = []
        self.pending_bytes = 0
        self.launched = 0
        self._flat = None           # (bucket, [views per segment], [slice per segment]) of the flat exchange buffer
        # YVB200_EXCHANGE_FLAT=0: one all-reduce per gradient tensor, grouped per segment (round-1 behaviour)
        self.flat = os.environ.get("YVB200_EXCHANGE_FLAT", "1") != "0"

    # ---- called while the step is being issued (eagerly or under capture)
    def begin(self):
        self.recording = True
        self.segments, self.pending, self.pending_bytes = [], [], 0
        self.pending_params, self.segment_params = [], []
        self._flat = None

    def _on_grad(self, p: torch.Tensor):
        if not self.recording or p.grad is None:
            return
        self.pending.append(p.grad)
        self.pending_params.append(p)
        self.pending_bytes += p.grad.numel() * p.grad.element_size()
        if self.overlap and self.pending_bytes >= self.segment_bytes:
            self._close_segment(with_event=True)

    def _close_segment(self, with_event: bool):
        if not self.pending:
            return
        evs = None
        if with_event and self.cuda:
            # the gradients of this segment were produced on the issuing stream and on the runtime's helper / fork /
            # branch streams (trailing weight gradients): one external event per stream, recorded where each stream
            # stands now; nothing is joined into the dependency chain of the backward pass
            from . import ops
            r = ops.rt(self.device)
            cur = torch.cuda.current_stream(self.device)
            capturing = torch.cuda.is_current_stream_capturing()
            evs = []
            for s_ in [cur, r.branch_stream] + list(r._helpers.values()):
                if s_ != cur and capturing:                   # only streams that belong to this capture
                    with torch.cuda.stream(s_):
                        if not torch.cuda.is_current_stream_capturing():
                            continue
                ev = torch.cuda.Event(external=True)
                ev.record(s_)
                evs.append(ev)
        self.segments.append((evs, self.pending))
        self.segment_params.append(self.pending_params)
        self.pending, self.pending_bytes, self.pending_params = [], 0, []

    def end(self):
        """Close the last segment (it is covered by the completion of the step itself)."""
        self._close_segment(with_event=False)
        self.recording = False

    def _build_flat(self):
        """One contiguous exchange buffer in backward order: each segment is a slice, each gradient a view of it.  NCCL
        then sees ONE large all-reduce per segment (full-bandwidth protocol; with many per-tensor operations grouped into
        one launch it falls back to the low-latency protocols that move half the payload per byte on the wire)."""
        dt = torch.bfloat16 if self.payload == "bf16" else torch.float32
        total = sum(g.numel() for _, grads in self.segments for g in grads)
        bucket = torch.empty(total, dtype=dt, device=self.device)
        views, slices, off = [], [], 0
        for _, grads in self.segments:
            start, vs = off, []
            for g in grads:
                vs.append(bucket[off:off + g.numel()].view_as(g))
                off += g.numel()
            views.append(vs)
            slices.append(bucket[start:off])
        self._flat = (bucket, views, slices)

    # ---- called after the step has been launched (after graph.replay() or the eager body)
    def exchange(self):
        if self.cuda and self.nccl and self.flat:
            if self._flat is None:
                self._build_flat()
            _, views, slices = self._flat
            main = torch.cuda.current_stream(self.device)
            for i, (evs, grads) in enumerate(self.segments):
                if evs is not None:
                    for ev in evs:
                        self.comm.wait_event(ev)
                else:
                    self.comm.wait_stream(main)
                with torch.cuda.stream(self.comm):
                    torch._foreach_copy_(views[i], grads)               # gather (and cast) into the flat slice
                    self.dist.all_reduce(slices[i], op=self.dist.ReduceOp.AVG, group=self.group)
                    if self.payload == "bf16":
                        torch._foreach_copy_(grads, views[i])           # back to the fp32 gradients
                self.launched += 1
            main.wait_stream(self.comm)
            if self.payload != "bf16":
                # the averaged gradients live in the flat buffer: hand those views to the parameters (no copy back)
                for ps, vs in zip(self.segment_params, views):
                    for p, v in zip(ps, vs):
                        p.grad = v
            return
        if self.cuda:
            main = torch.cuda.current_stream(self.device)
            for evs, grads in self.segments:
                if evs is not None:
                    for ev in evs:
                        self.comm.wait_event(ev)
                else:
                    self.comm.wait_stream(main)
                with torch.cuda.stream(self.comm):
                    self._reduce(grads)
                self.launched += 1
            main.wait_stream(self.comm)
        else:
            for _, grads in self.segments:
                self._reduce(grads)
                self.launched += 1

    def _reduce(self, grads):
        dist = self.dist
        if self.nccl and self.payload == "bf16":
            total = sum(g.numel() for g in grads)
            bucket = torch.empty(total, dtype=torch.bfloat16, device=self.device)
            views, off = [], 0
            for g in grads:
                views.append(bucket[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(views, grads)
            dist.all_reduce(bucket, op=dist.ReduceOp.AVG, group=self.group)
            torch._foreach_copy_(grads, views)
            return
        if self.nccl:                               # one grouped launch, averaging inside the collective
            with dist._coalescing_manager(group=self.group, device=self.device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            return
        flat = torch.cat([g.reshape(-1) for g in grads])        # gloo (CPU tests): flatten, sum, scatter back
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(float(self.world))
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def remove(self):
        for h in self.handles:
            h.remove()


class _Consumed(torch.autograd.Function):
    """Identity whose backward records an (external) CUDA event: placed on the logits that enter the two big losses,
    it fires right after the loss-gradient kernels have read the target tensors of the batch -- the last reads of the
    step's static input buffers.  ``GraphedStep.load`` lets the next batch's host-to-device copy wait for these events
    instead of for the end of the step."""

    @staticmethod
    def forward(ctx, x, owner):
        ctx.owner = owner
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ev = torch.cuda.Event(external=True)
        ev.record(torch.cuda.current_stream(g.device))
        ctx.owner._consumed.append(ev)
        return g, None


class GraphedStep:
    #: static inputs read in place by the step (everything else is small and cloned into step-private memory first):
    #: image_features (read once by the plane split at the start of forward) and image_targets (read by the KL loss
    #: and its gradient kernel at the start of backward)
    BIG_INPUTS = (1, 4)

    def __init__(self, model: torch.nn.Module, args, example_batch: List[torch.Tensor], use_graph: bool = True,
                 refresh_weights_each_step: bool = True, warmup: int = 2, exchange: Optional[GradientExchange] = None,
                 prefetch: bool = True, fast_heads: Optional[bool] = None):
        self.model, self.args = model, args
        self.prefetch = prefetch
        self._consumed: List[torch.cuda.Event] = []
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._loaded: Optional[torch.cuda.Event] = None
        self._launched = False
        self.exchange = exchange
        # (the exchange's NCCL calls stay outside the graph: inside it only per-segment external events are recorded)
        self.device = next(model.parameters()).device
        self.prefetch = self.prefetch and self.device.type == "cuda"
        self.rt = ops.rt(self.device)
        self.refresh = refresh_weights_each_step
        self.static = [t.to(self.device).clone() if torch.is_tensor(t) else t for t in example_batch]
        if fast_heads is None:
            import os
            fast_heads = os.environ.get("YVB200_FAST_HEADS", "1") != "0"
        # the compacted-head path needs the Lily layout (bert / cls / vil_logit / judge); anything else runs generically
        self.fast_heads = fast_heads and all(hasattr(model, a) for a in ("bert", "cls", "vil_logit", "judge", "dropout",
                                                                          "fusion_method"))
        if not self.fast_heads and not bool(self.static[13].all()):
            raise RuntimeError("GraphedStep without fast_heads needs batches without padded candidates (opt_mask all ones)")
        self.losses = {}
        self.loss = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self.use_graph = use_graph
        side = torch.cuda.Stream(device=self.device, priority=self.rt.main_priority)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._zero_grads()
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if use_graph and self.refresh:
            # the first forward may have added weights to the arena: rebuild the segment tables now, a capture
            # cannot do the host-to-device copy
            self.rt.arena.refresh_all(force=True)
            torch.cuda.synchronize(self.device)
        if use_graph:
            self._zero_grads()
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(self.graph, stream=side):
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
        # (the dropout step counter is advanced by BertModel.forward itself, once per training forward)
        if self.refresh:
            # chunk 0 (first layers) on this stream, the rest overlapped with the start of the forward pass
            self.rt.arena.refresh_all(force=True, overlap=self.rt.concurrent)
        self.rt.arena.skip_next_refresh = True      # the model's own start-of-forward refresh would repeat the work
        b = self.static
        if self.prefetch:
            # the next batch may be copied into the static buffers while this step is still running (see ``load``):
            # small inputs (token ids, locations, masks, targets: some are read again at the very end of backward) are
            # cloned into step-private memory, the two big ones are read in place and guarded by ``_consumed`` events
            self._consumed = []
            b = [t.clone() if (torch.is_tensor(t) and i not in self.BIG_INPUTS) else t for i, t in enumerate(b)]
        co = b[11]
        inputs = (b[6].flatten(0, 1), b[1].flatten(0, 1), b[2].flatten(0, 1), b[10].flatten(0, 1), b[7].flatten(0, 1),
                  b[3].flatten(0, 1), co.reshape(-1, co.size(2), co.size(3)), b[9].flatten(0, 1), b[15])
        if self.fast_heads:
            ld = self._forward_fast(b, inputs)
        else:
            ld = self._forward_generic(b, inputs)
        tot = 0.0
        for k in ("vision", "language", "ranking"):
            if k in ld:
                tot = tot + ld[k]
        if "traj" in ld:
            tot = tot + self.args.traj_loss_scale * ld["traj"]
        self.losses = {k: v.detach() for k, v in ld.items()}
        # metrics of the step (utils/utils_init.py:167-189) packed as [task, {loss, correct, batch size}] on the device:
        # no .item(), and ONE all-reduce for all tasks when ``metrics`` is asked for (the reference issues three per task)
        self.metric_tasks = [k for k in ("vision", "language", "ranking", "traj") if k in ld]
        bsz = float(b[13].shape[0])
        zero = torch.zeros((), device=self.device)
        corr = getattr(self, "_correct", None) or {}
        self.metric_pack = torch.stack([torch.stack([ld[k].detach().float(), corr.get(k, zero).float(),
                                                     torch.full((), bsz, device=self.device)])
                                        for k in self.metric_tasks])
        if self.exchange is not None:
            self.exchange.begin()
        # weight gradients may trail on the helper streams: they are joined at segment ends by the exchange and at the
        # end of the pass by the runtime's autograd callback
        self.rt.defer_wgrad = self.rt.defer_wgrad_allowed
        try:
            tot.backward()
        finally:
            self.rt.defer_wgrad = False
        if self.exchange is not None:
            self.exchange.end()
        self.loss = tot.detach()

    def _forward_generic(self, b, inputs):
        """Any model with the ``Lily`` call signature: full logits, fused loss kernels on all rows."""
        out = self.model(*inputs)
        self.rt.arena.join()
        if self.prefetch:
            out = dict(out)
            marked = False
            for k in ("vision", "language"):
                if k in out and out[k].requires_grad:
                    out[k] = _Consumed.apply(out[k], self)
                    marked = True
            if not marked or "vision" not in out:       # image_targets unused: the inputs are free after forward
                ev = torch.cuda.Event(external=True)
                ev.record(torch.cuda.current_stream(self.device))
                self._consumed.append(ev)
        return fused.step_losses(b, out, self.args, training=True, flat=True)

    def _forward_fast(self, b, inputs):
        """``Lily.forward`` (lily.py:58-129) + ``get_loss_correct`` (utils/utils_init.py:108-164) composed from the
        model's own sub-modules so that the two big heads only see their supervised rows (``fused.language_head_loss`` /
        ``vision_head_loss``): the [N, T, 30522] and [N, V, 1601] logits of un-supervised positions -- 85 % of them --
        are never computed.  Every candidate slot of the [bs, C] batch runs (static shapes); padded candidates
        (``opt_mask``, utils/utils_init.py:54-61) are masked out of every loss on the device."""
        m, args = self.model, self.args
        seq_t, seq_v, pooled_t, pooled_v, _ = m.bert(
            input_txt=inputs[0], input_imgs=inputs[1], image_loc=inputs[2], token_type_ids=inputs[3],
            attention_mask=inputs[4], image_attention_mask=inputs[5], co_attention_mask=inputs[6],
            output_all_encoded_layers=False)
        self.rt.arena.join()
        valid = b[13].flatten().bool()
        ld = {}
        cur = torch.cuda.current_stream(self.device)
        side = self.rt.branch_stream if (self.rt.concurrent and args.masked_vision and args.masked_language) else None
        if args.masked_vision:
            tgt = b[4].flatten(0, 1)
            rows = tgt.shape[0] * tgt.shape[1]
            if self.prefetch and fused.head_capacity(rows) >= rows:
                tgt = tgt.clone()                   # (no compaction: the loss backward would read the static buffer late)
            msk = b[5].flatten(0, 1) * valid[:, None].to(b[5].dtype)
            if side is not None:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ld["vision"] = fused.vision_head_loss(m.cls.imagePredictions, seq_v, tgt, msk)
            else:
                ld["vision"] = fused.vision_head_loss(m.cls.imagePredictions, seq_v, tgt, msk)
        if args.masked_language:
            lm_t = b[8].flatten(0, 1).masked_fill(~valid[:, None], -1)
            ld["language"] = fused.language_head_loss(m.cls.predictions, seq_t, lm_t)
        if side is not None:
            cur.wait_stream(side)
            ld["vision"].record_stream(cur)
        if self.prefetch:                           # both big static inputs have been read for the last time
            ev = torch.cuda.Event(external=True)
            ev.record(cur)
            self._consumed.append(ev)
        if m.fusion_method == "sum":
            pooled = pooled_t + pooled_v
        elif m.fusion_method == "mul":
            pooled = pooled_t * pooled_v
        else:
            assert False
        pooled = m.dropout(pooled)
        small = {}
        if args.ranking:
            small["ranking"] = m.vil_logit(pooled)
        if args.traj_judge:
            small["traj"] = m.judge(pooled)
        self._correct = {}
        ld.update(fused.small_losses(b, small, args, training=True, correct=self._correct))
        return ld

    def load(self, batch: List[torch.Tensor]):
        """Copy a (pinned host or device) batch into the static buffers.  With ``prefetch`` the copy runs on its own
        stream as soon as the step in flight has finished reading its inputs (start of its backward pass), i.e.
        overlapped with the rest of that step; the next ``run`` waits for the copy."""
        if not (self.prefetch and self.device.type == "cuda"):
            for dst, src in zip(self.static, batch):
                if torch.is_tensor(dst):
                    dst.copy_(src, non_blocking=True)
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        cur = torch.cuda.current_stream(self.device)
        if self._launched and self._consumed:
            for ev in self._consumed:
                cs.wait_event(ev)
        else:
            cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            for dst, src in zip(self.static, batch):
                if torch.is_tensor(dst):
                    dst.copy_(src, non_blocking=True)
            self._loaded = torch.cuda.Event()
            self._loaded.record(cs)

    def metrics(self, reduce: bool = True):
        """``reduced_metrics`` of the last step as the reference's ``compute_metrics_independent`` builds it
        (utils/utils_init.py:167-189): ``{"loss": {task: mean over ranks}, "accuracy": {ranking / traj: correct / batch}}``
        -- device scalars; with a gradient exchange attached the ranks are combined by a single packed all-reduce."""
        pack = self.metric_pack.clone()
        world = 1.0
        if reduce and self.exchange is not None and self.exchange.world > 1:
            world = float(self.exchange.world)
            self.exchange.dist.all_reduce(pack, op=self.exchange.dist.ReduceOp.SUM, group=self.exchange.group)
        out = {"loss": {}, "accuracy": {}}
        for i, k in enumerate(self.metric_tasks):
            out["loss"][k] = pack[i, 0] / world
            if k not in ("vision", "language"):
                out["accuracy"][k] = pack[i, 1] / pack[i, 2]
        return out

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.static if torch.is_tensor(t))

    def run(self) -> torch.Tensor:
        """One step on the data currently in the static buffers; returns the (device) total loss."""
        if self._loaded is not None:
            torch.cuda.current_stream(self.device).wait_event(self._loaded)
            self._loaded = None
        self._launched = True
        if self.graph is not None:
            self.graph.replay()
        else:
            self._zero_grads()
            self._body()
        if self.exchange is not None:
            self.exchange.exchange()
        return self.loss
