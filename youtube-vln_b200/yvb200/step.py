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


class _Plan:
    """Frozen layout of the gradient-exchange buffer: parameters in the order their gradients become final, cut into
    segments; every gradient has a fixed view of ONE flat buffer, every segment is a contiguous slice of it."""

    def __init__(self):
        self.order: List[torch.Tensor] = []         # parameters in arrival order
        self.seg_end = set()                        # indices into ``order`` after which a segment closes
        self.segments: List[List[torch.Tensor]] = []
        self.bucket = None
        self.symm = None                            # symmetric-memory handle of the bucket (ce / nvls)
        self.views: List[List[torch.Tensor]] = []   # per segment, per parameter (arrival order)
        self.slices: List[torch.Tensor] = []        # per segment (padded to world * 32 elements)
        self.sink_keys: List[int] = []              # runtime.grad_sink entries owned by this plan


class GradientExchange:
    """Data-parallel gradient averaging overlapped with backward.

    A post-accumulate hook on every parameter collects finished gradients; every ``segment_mb`` of them closes a
    *segment*: an **external** CUDA event is recorded on every stream that produced gradients (inside a captured step
    these are event-record nodes of the graph).  After the step is launched, the communication stream waits for each
    segment's events in turn and averages that segment across the ranks -- one NCCL all-reduce (AVG) of the segment's
    slice of a flat buffer -- while the rest of the backward graph is still executing.  NCCL stays outside the graph.
    The reference gets the same overlap from ``DistributedDataParallel`` bucket hooks (utils/distributed.py:97-99).

    The first pass only observes (arrival order and sizes) and is exchanged through a throw-away buffer; from it a
    *plan* is frozen: the segmentation and a fixed view of the flat buffer for every gradient; later passes copy their
    gradients into those views (one multi-tensor copy per segment) before the collective.  With ``direct`` (opt-in, CUDA)
    the plan also registers a *sink* for every GEMM weight: the weight-gradient kernels then write straight into the flat
    buffer (``ops._dw_out``), autograd adopts those views as ``.grad``, and only what still arrives elsewhere (biases,
    LayerNorm parameters, the tied embedding) is copied.  A pass whose arrival order deviates from the plan falls back to
    the observing path and re-plans.
    """

    def __init__(self, model: torch.nn.Module, group=None, segment_mb: float = None, overlap: bool = True,
                 direct: Optional[bool] = None):
        import torch.distributed as dist
        if segment_mb is None:
            segment_mb = float(os.environ.get("YVB200_SEGMENT_MB", "320"))
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.segment_bytes = int(segment_mb * 2 ** 20)
        # YVB200_SEGMENT_MIN_MB < YVB200_SEGMENT_MB makes planned segments shrink towards the end of the backward pass
        # (target = half of what is still to come, at least this much).  Measured on 2 and 8 GPUs the fixed size was as
        # good or better (profiles/README.md), so the default is "no shrinking".
        self.segment_min_bytes = int(float(os.environ.get("YVB200_SEGMENT_MIN_MB", str(segment_mb))) * 2 ** 20)
        self.overlap = overlap
        self.device = next(model.parameters()).device
        self.cuda = self.device.type == "cuda"
        self.nccl = dist.get_backend(group) == "nccl"
        # YVB200_EXCHANGE_DTYPE=bf16 (opt-in): half the bytes over NVLink, at bf16 rounding (2^-9 relative) of every
        # averaged gradient element -- NOT the reference's fp32 DistributedDataParallel arithmetic
        self.payload = os.environ.get("YVB200_EXCHANGE_DTYPE", "fp32")
        if self.payload not in ("fp32", "bf16"):
            raise RuntimeError("YVB200_EXCHANGE_DTYPE must be fp32 or bf16")
        if not self.cuda:
            self.payload = "fp32"
        # YVB200_EXCHANGE_FLAT=0: one all-reduce per gradient tensor, grouped per segment (round-1 behaviour)
        self.flat = os.environ.get("YVB200_EXCHANGE_FLAT", "1") != "0"
        # How a segment's slice of the flat buffer is averaged across ranks:
        #   nccl : one NCCL all-reduce (AVG) per segment -- its channel CTAs share the SMs with the backward pass
        #   ce   : the flat buffer lives in symmetric (peer-mapped) memory; reduce-scatter and all-gather are peer-to-peer
        #          copies on the copy engines over NVLink, the only SM work is the local mean of this rank's chunk
        #   nvls : symmetric memory + the in-switch (multimem) all-reduce kernel of torch's symm_mem library (4 CTAs)
        # ce / nvls fall back to nccl (with a warning on rank 0) when symmetric memory cannot be set up on this machine.
        # Measured (profiles/README.md): nccl is the fastest of the three on 2 and on 8 GPUs, hence the default.
        self.transport = os.environ.get("YVB200_EXCHANGE", "nccl")
        if self.transport not in ("nccl", "ce", "nvls"):
            raise RuntimeError("YVB200_EXCHANGE must be nccl, ce or nvls")
        if not (self.cuda and self.nccl and self.flat and self.payload == "fp32") or self.world == 1:
            self.transport = "nccl"
        # (copy-engine transport: its few tiny kernels -- barriers, the chunk mean -- must not queue behind the backward
        # pass; NCCL's channel CTAs keep the default priority)
        base = ops.rt(self.device).helper_priority if self.cuda else 0      # (the level of the trailing compute streams)
        self.comm = torch.cuda.Stream(device=self.device, priority=base - 1 if self.transport == "ce" else base) if self.cuda else None
        # direct (YVB200_EXCHANGE_DIRECT=1): planned passes write weight gradients straight into the flat buffer instead
        # of copying them there before the collective.  Opt-in: it removes ~0.8 ms of copy kernels per step, but on 2
        # GPUs the step was not faster (the collectives then start earlier and their channel CTAs overlap more of the
        # backward chain: 10.75 ms with copies, 10.8-11.2 ms without; profiles/README.md)
        if direct is None:
            direct = os.environ.get("YVB200_EXCHANGE_DIRECT", "0") != "0"
        self.direct = bool(direct) and self.cuda and self.payload == "fp32"
        self._symm_cache = {}       # ce / nvls: {elements: (flat buffer in symmetric memory, its handle)}
        self._stage = None          # ce: [world - 1, largest chunk] landing area of the reduce-scatter pulls
        self._copy2 = None          # ce: second copy stream (two copy engines in flight)
        self.barrier_timeout_ms = int(os.environ.get("YVB200_EXCHANGE_TIMEOUT_MS", "30000"))
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in model.parameters()
                        if p.requires_grad]
        self.recording = False
        self.segments = []          # [(events or None, [grads])] of the pass being issued / captured
        self.segment_params: List[List[torch.Tensor]] = []
        self.pending: List[torch.Tensor] = []
        self.pending_params: List[torch.Tensor] = []
        self.pending_bytes = 0
        self.closed_bytes = 0
        self.launched = 0
        self.plan: Optional[_Plan] = None        # frozen after the first observed pass
        self._pass_plan: Optional[_Plan] = None  # the plan the recorded pass followed (None: observed pass)
        self._cursor = 0
        self._copies = []           # planned pass: per segment ([views], [grads]) that still need a copy into place
        self._flat = None           # observed pass: (bucket, views, slices) built at its first exchange

    # ---- called while the step is being issued (eagerly or under capture)
    def begin(self):
        self.recording = True
        self.segments, self.pending, self.pending_bytes = [], [], 0
        self.pending_params, self.segment_params = [], []
        self.closed_bytes = 0
        self._flat = None
        self._pass_plan, self._cursor = self.plan, 0

    def _on_grad(self, p: torch.Tensor):
        if not self.recording or p.grad is None:
            return
        self.pending.append(p.grad)
        self.pending_params.append(p)
        self.pending_bytes += p.grad.numel() * p.grad.element_size()
        plan = self._pass_plan
        if plan is not None:
            i = self._cursor
            if i < len(plan.order) and plan.order[i] is p:
                self._cursor = i + 1
                if self.overlap and i in plan.seg_end:
                    self._close_segment(with_event=True)
                return
            self._pass_plan = None                  # arrival order changed: observe the rest of this pass, re-plan after
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
        self.closed_bytes += self.pending_bytes
        self.pending, self.pending_bytes, self.pending_params = [], 0, []

    def end(self):
        """Close the last segment (it is covered by the completion of the step itself); freeze or check the plan."""
        self._close_segment(with_event=False)
        self.recording = False
        plan = self._pass_plan
        if plan is not None and self._cursor == len(plan.order) and len(self.segments) == len(plan.segments):
            # planned pass: which gradients did not land in their slot of the flat buffer by themselves?
            self._copies = []
            for (_, grads), views in zip(self.segments, plan.views):
                todo = [(v, g) for v, g in zip(views, grads) if g.data_ptr() != v.data_ptr()]
                self._copies.append(([v for v, _ in todo], [g for _, g in todo]))
            return
        self._pass_plan = None
        capturing = self.cuda and torch.cuda.is_current_stream_capturing()
        if self.flat and not capturing and self.segments:
            self._freeze_plan()

    # ---- the plan
    def _planned_segments(self, params: List[torch.Tensor]) -> List[List[torch.Tensor]]:
        total = sum(p.numel() * 4 for p in params)
        segs, cur, cur_b, closed = [], [], 0, 0
        for p in params:
            cur.append(p)
            cur_b += p.numel() * 4
            target = min(self.segment_bytes, max(self.segment_min_bytes, (total - closed) // 2))
            if self.overlap and cur_b >= target:
                segs.append(cur)
                closed += cur_b
                cur, cur_b = [], 0
        if cur:
            segs.append(cur)
        return segs

    def _freeze_plan(self):
        """Collective (every rank gets here at the end of its first observed pass): lay out the flat buffer."""
        self._drop_sinks()
        order = [p for ps in self.segment_params for p in ps]
        if len({id(p) for p in order}) != len(order):
            return                                   # a gradient was reported twice (accumulation inside the pass)
        plan = _Plan()
        plan.order = order
        plan.segments = self._planned_segments(order)
        n = 0
        for seg in plan.segments[:-1]:
            n += len(seg)
            plan.seg_end.add(n - 1)
        # fused weights (Q|K|V of one projection: one arena entry, one weight-gradient GEMM) sit next to each other, in
        # the row order of that GEMM's output, at the position of their first member
        groups = {}
        if self.direct:
            arena = ops.rt(self.device).arena
            arena.prune()
            uses = {}
            for e in arena.entries.values():
                for q in e.params:
                    uses[id(q)] = uses.get(id(q), 0) + 1
            for e in arena.entries.values():
                if all(uses[id(q)] == 1 and q.requires_grad for q in e.params):
                    for q in e.params:
                        groups[id(q)] = e
        quantum = self.world * 32
        layout, lens, sinks = [], [], []             # layout: per segment [(param, offset in the bucket)]
        off = 0
        for seg in plan.segments:
            in_seg = {id(q) for q in seg}
            placed, start, where = set(), off, {}
            for q in seg:
                if id(q) in placed:
                    continue
                e = groups.get(id(q))
                members = list(e.params) if (e is not None and all(id(m) in in_seg for m in e.params)) else [q]
                first = off
                for m in members:
                    where[id(m)] = off
                    placed.add(id(m))
                    off += m.numel()
                if e is not None and len(members) == len(e.params):
                    sinks.append((e, first))
            layout.append([(q, where[id(q)]) for q in seg])
            off = start + (off - start + quantum - 1) // quantum * quantum
            lens.append(off - start)
        total = off
        dt = torch.bfloat16 if self.payload == "bf16" else torch.float32
        got = self._symmetric_bucket(total) if self.transport != "nccl" else None
        if got is None:
            self.transport = "nccl"
            bucket = torch.zeros(total, dtype=dt, device=self.device)
        else:
            bucket, plan.symm = got
        plan.bucket = bucket
        start = 0
        for seg_layout, ln in zip(layout, lens):
            plan.views.append([bucket[o:o + q.numel()].view_as(q) for q, o in seg_layout])
            plan.slices.append(bucket[start:start + ln])
            start += ln
        if self.transport == "ce":
            self._ce_buffers(max(lens))
        if self.direct:
            r = ops.rt(self.device)
            for e, o in sinks:
                key = e.planes.addr
                r.grad_sink[key] = (bucket[o:o + e.rows * e.cols].view(e.rows, e.cols), tuple(e.params))
                plan.sink_keys.append(key)
        self.plan = plan

    def _drop_sinks(self):
        if self.plan is not None and self.plan.sink_keys:
            r = ops.rt(self.device)
            for k in self.plan.sink_keys:
                r.grad_sink.pop(k, None)
        self.plan = None

    def _ce_buffers(self, longest: int):
        n = longest // self.world
        if self._stage is None or self._stage.shape[1] < n:
            self._stage = torch.empty(self.world - 1, n, dtype=torch.float32, device=self.device)
        if self._copy2 is None:
            self._copy2 = torch.cuda.Stream(device=self.device, priority=self.comm.priority)

    def _build_flat(self):
        """Observed pass (no plan yet, or the plan did not fit): a throw-away flat buffer in arrival order; every
        gradient is copied into it.  Slices are padded to ``world * 32`` elements so that every rank owns a 128-byte
        aligned chunk of each of them."""
        dt = torch.bfloat16 if self.payload == "bf16" else torch.float32
        quantum = self.world * 32
        lens = [(sum(g.numel() for g in grads) + quantum - 1) // quantum * quantum for _, grads in self.segments]
        total = sum(lens)
        got = self._symmetric_bucket(total) if self.transport != "nccl" else None
        handle = None
        if got is None:
            self.transport = "nccl"
            bucket = torch.zeros(total, dtype=dt, device=self.device)
        else:
            bucket, handle = got
        views, slices, start = [], [], 0
        for (_, grads), n in zip(self.segments, lens):
            off, vs = start, []
            for g in grads:
                vs.append(bucket[off:off + g.numel()].view_as(g))
                off += g.numel()
            views.append(vs)
            slices.append(bucket[start:start + n])
            start += n
        self._flat = (bucket, views, slices, handle)
        if self.transport == "ce":
            self._ce_buffers(max(lens))

    def _symmetric_bucket(self, total: int):
        """A flat buffer in symmetric memory (every rank maps every peer's copy) and its handle, or ``None`` if that
        cannot be had here.  Collective: all ranks get here together.  One buffer per size (the plan's; an observed
        pass of the same size shares it -- the two are never in flight together)."""
        hit = self._symm_cache.get(total)
        if hit is not None:
            return hit
        ok, bucket, handle, why = 1, None, None, ""
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else self.dist.group.WORLD
            bucket = symm.empty(total, dtype=torch.float32, device=self.device)
            handle = symm.rendezvous(bucket, group)
            self._group_name = group.group_name
            if self.transport == "nvls" and not handle.has_multicast_support:
                raise RuntimeError("no multicast support")
            bucket.zero_()
        except Exception as e:                       # noqa: BLE001 -- any failure means "not available on this machine"
            ok, why = 0, f"{type(e).__name__}: {e}"
        flag = torch.tensor([ok], device=self.device)
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN, group=self.group)     # all ranks or none
        if int(flag.item()) == 0:
            if self.dist.get_rank(self.group) == 0:
                import warnings
                warnings.warn(f"yvb200: YVB200_EXCHANGE={self.transport} needs symmetric memory, which is not available "
                              f"here ({why or 'failed on another rank'}); using NCCL")
            return None
        self._symm_cache[total] = (bucket, handle)
        return bucket, handle

    # The three data phases of the copy-engine exchange as pure functions of (rank, world, get_buffer): the CUDA path
    # below adds streams and barriers, tests/test_exchange_cpu.py drives them for several fake ranks in one process.
    @staticmethod
    def ce_pull_chunks(rank, world, get_buffer, base, n, stage, run=None):
        """Reduce-scatter, data movement: chunk ``rank`` of every peer's slice -> ``stage[step - 1]``."""
        for step in range(1, world):
            peer = (rank - step) % world
            src = get_buffer(peer, (n,), torch.float32, base + rank * n)
            (run or (lambda k, f: f()))(step, lambda d=stage[step - 1], s_=src: d.copy_(s_, non_blocking=True))

    @staticmethod
    def ce_reduce(rank, world, sl, n, stage):
        """Mean over the ranks, summed in rank order, of this rank's chunk (in place in the slice)."""
        own = sl[rank * n:(rank + 1) * n]
        order = [(rank - step) % world for step in range(1, world)]
        acc = None
        for r_ in range(world):
            term = own if r_ == rank else stage[order.index(r_)]
            acc = term.clone() if acc is None else acc.add_(term)
        own.copy_(acc.mul_(1.0 / world))

    @staticmethod
    def ce_pull_reduced(rank, world, get_buffer, base, n, sl, run=None):
        """All-gather: the reduced chunk ``peer`` of every peer's slice -> the same chunk of this rank's slice."""
        for step in range(1, world):
            peer = (rank - step) % world
            src = get_buffer(peer, (n,), torch.float32, base + peer * n)
            (run or (lambda k, f: f()))(step, lambda d=sl[peer * n:(peer + 1) * n], s_=src: d.copy_(s_, non_blocking=True))

    def _exchange_ce(self, sl: torch.Tensor, h):
        """Average a slice of the flat buffer over the ranks with peer-to-peer copies (issued on the communication
        stream and a second copy stream, executed by the copy engines): pull this rank's chunk of every peer's slice,
        take the mean locally (one rank reduces each chunk, in fixed rank order: every rank ends up with bit-identical
        gradients), pull every peer's reduced chunk.  Barriers (tiny signal kernels of the symmetric-memory handle)
        order the three phases."""
        world = self.world
        rank = h.rank
        n = sl.numel() // world
        base = sl.storage_offset()
        stage = self._stage[:, :n]
        c2, t = self._copy2, self.barrier_timeout_ms

        def run(step, copy):                                     # alternate the copies over two streams
            with torch.cuda.stream(c2 if step % 2 else self.comm):
                copy()

        def fork():
            ready = torch.cuda.Event()
            ready.record(self.comm)
            c2.wait_event(ready)

        h.barrier(0, t)                                          # every rank has its slice in place
        fork()
        self.ce_pull_chunks(rank, world, h.get_buffer, base, n, stage, run)
        self.comm.wait_stream(c2)
        lib.mean_chunks(sl[rank * n:(rank + 1) * n], self._stage, world, rank)      # (ce_reduce as one kernel)
        h.barrier(0, t)                                          # every chunk is reduced where it lives
        fork()
        self.ce_pull_reduced(rank, world, h.get_buffer, base, n, sl, run)
        self.comm.wait_stream(c2)
        h.barrier(0, t)                                          # nobody still reads a slice its owner may overwrite

    def _average(self, sl: torch.Tensor, handle=None):
        if self.transport == "ce":
            self._exchange_ce(sl, handle)
        elif self.transport == "nvls":
            torch.ops.symm_mem.multimem_all_reduce_(sl, "sum", self._group_name)
            sl.mul_(1.0 / self.world)
        elif self.nccl:
            self.dist.all_reduce(sl, op=self.dist.ReduceOp.AVG, group=self.group)
        else:                                                    # gloo (CPU tests)
            self.dist.all_reduce(sl, op=self.dist.ReduceOp.SUM, group=self.group)
            sl.div_(float(self.world))

    # ---- called after the step has been launched (after graph.replay() or the eager body)
    def exchange(self):
        if self.flat:
            plan = self._pass_plan
            if plan is not None:
                views, slices, copies, handle = plan.views, plan.slices, self._copies, plan.symm
            else:
                if self._flat is None:
                    self._build_flat()
                _, views, slices, handle = self._flat
                copies = [(vs, grads) for vs, (_, grads) in zip(views, self.segments)]
            main = torch.cuda.current_stream(self.device) if self.cuda else None
            for i, (evs, grads) in enumerate(self.segments):
                if self.cuda:
                    if evs is not None:
                        for ev in evs:
                            self.comm.wait_event(ev)
                    else:
                        self.comm.wait_stream(main)
                with (torch.cuda.stream(self.comm) if self.cuda else _null_context()):
                    if copies[i][0]:
                        torch._foreach_copy_(copies[i][0], copies[i][1])    # what is not in place yet (cast for bf16)
                    self._average(slices[i], handle)
                    if self.payload == "bf16":
                        torch._foreach_copy_(grads, views[i])               # back to the fp32 gradients
                self.launched += 1
            if self.cuda:
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

    def in_place_fraction(self) -> float:
        """Share of the gradient bytes of the last planned pass that needed no copy into the flat buffer."""
        if self._pass_plan is None:
            return 0.0
        total = sum(g.numel() for _, grads in self.segments for g in grads)
        copied = sum(g.numel() for _, gs in self._copies for g in gs)
        return 1.0 - copied / max(total, 1)

    def verify(self, samples: int = 16):
        """Largest relative error of the exchanged gradients against a plain all-gather + fp64 mean of the ranks' local
        gradients, over ``samples`` tensors spread across the segments -- only tensors that were *copied* into the flat
        buffer still have their local value (flat fp32 exchange only, else ``None``).  A diagnostic for new transports:
        collective, every rank must call it."""
        if not self.flat or self.payload != "fp32":
            return None
        if self._pass_plan is not None:
            pairs = [(g, v) for vs, gs in self._copies for v, g in zip(vs, gs)]
        elif self._flat is not None:
            pairs = [(g, v) for (_, grads), vs in zip(self.segments, self._flat[1]) for g, v in zip(grads, vs)]
        else:
            return None
        worst = 0.0
        for g, v in pairs[:: max(1, len(pairs) // samples)]:
            parts = [torch.empty_like(g) for _ in range(self.world)]
            self.dist.all_gather(parts, g.contiguous(), group=self.group)
            ref = torch.stack([x.double() for x in parts]).mean(0)
            worst = max(worst, float((v.double() - ref).norm() / ref.norm().clamp_min(1e-30)))
        return worst

    def remove(self):
        for h in self.handles:
            h.remove()
        self._drop_sinks()


class _null_context:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


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
