"""Fused building blocks of the two-stream ViLBERT path as ``torch.autograd.Function``s over the C ABI.

Each block mirrors one sub-module of the reference (vilbert/vilbert.py) and owns a hand-written backward:

  DenseAct       act(x W^T + b)                          BertIntermediate :351-354, poolers :827-848, decoders
  DenseResLN     LN(dropout(x W^T + b) + r)              Bert(Image)SelfOutput / Output :321-325,:364-368, BiOutput
  DenseActLN     LN(gelu(x W^T + b))                     Bert(Img)PredictionHeadTransform :863-886
  SelfAttention  QKV proj -> softmax(QK^T/sqrt(d)+m) V   Bert(Image)SelfAttention :284-311, :413-440
  BiAttention    both cross-stream directions            BertBiAttention :552-618
  AttnBlock      attention + out-proj + residual + LN    Bert(Image)Attention :328-337, :456-464 (one node)
  FFN            LN(dropout(W2 gelu(W1 x)) + x)          Bert(Image)Intermediate+Output (one node)
  TextEmbed      gather-sum -> LN -> dropout             BertEmbeddings :240-256
  ImageEmbed     2048->H GEMM + location terms -> LN     BertImageEmbeddings :1356-1370

All contractions run on ``yv_gemm`` (tcgen05); operands travel as bf16 hi/lo planes (bf16x3 by default so the
end-to-end result stays within 1e-3 of the fp32 reference; YVB200_PRECISION=bf16 selects plain bf16).
Nothing here falls back to torch math: tensors must live on a CUDA device and the library must load.
"""
from __future__ import annotations

import itertools
import math
import os
import types
import weakref
from typing import Dict, List, Optional, Tuple

import torch
from torch.autograd import Function

from . import lib as L
from .lib import Planes

LN_EPS = 1e-12

# ----------------------------------------------------------------------------------------------------
# runtime state
# ----------------------------------------------------------------------------------------------------
_site_counter = itertools.count(1)


def new_site() -> int:
    """Unique id of a dropout call site (one per module instance), part of the RNG key."""
    return next(_site_counter)


class Runtime:
    def __init__(self, device: torch.device):
        self.device = device
        mode = os.environ.get("YVB200_PRECISION", "bf16x3")
        if mode not in ("bf16x3", "bf16"):
            raise RuntimeError(f"YVB200_PRECISION must be bf16x3 or bf16, got {mode}")
        self.passes = 3 if mode == "bf16x3" else 1
        self.attn_passes = int(os.environ.get("YVB200_ATTN_PASSES", self.passes))
        # dropout RNG: {seed, step counter}.  ``master_rng`` is advanced once per training forward
        # (``begin_training_forward``); ``rng`` is the snapshot the nodes of the current forward -- and later their
        # backward, which regenerates the same masks -- read, so that a second forward issued before the first
        # backward cannot change the masks of the first.  Ranks of a data-parallel job draw different masks: the
        # default seed mixes in RANK (set YVB200_SEED to pin it).
        seed = os.environ.get("YVB200_SEED")
        seed = int(seed) if seed is not None else 20231117 + 7919 * int(os.environ.get("RANK", "0"))
        self.master_rng = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self.rng = self.master_rng.clone()
        self.arena = WeightArena(device)
        # independent work (dgrad vs wgrad, text vs vision stream) is issued on helper CUDA streams so that the
        # captured step graph has parallel branches; YVB200_CONCURRENT=0 serialises everything on one stream
        self.concurrent = os.environ.get("YVB200_CONCURRENT", "1") != "0"
        self._helpers: Dict[Tuple[int, int], torch.cuda.Stream] = {}
        self._helper_pool: Dict[int, List[torch.cuda.Stream]] = {}
        self._helper_next: Dict[int, int] = {}
        self._forks: Dict[Tuple[int, int], torch.cuda.Stream] = {}
        self.helpers_per_stream = int(os.environ.get("YVB200_HELPERS", "1"))   # measured: 1 best (2: +0.15 ms)
        # stream priorities (retained as kernel-node priorities when the step is captured): the two stream-level
        # branches (text / vision) carry the critical dependency chain and run at high priority; helper streams
        # carry work nobody waits for until the end of a block (weight gradients, bias gradients) and fill in behind
        # weight / bias gradients of the encoder layers are not joined back into the dependency chain of the backward
        # pass: they trail on the helper streams and are joined once at the end (YVB200_DEFER_WGRAD=0: join at once)
        # Off by default: a consumer that reads gradients DURING backward on its own stream (stock
        # DistributedDataParallel bucket hooks) would not see the helper streams.  yvb200.step.GraphedStep, whose
        # GradientExchange joins the helper streams itself, switches it on for its backward pass.
        self.defer_wgrad = False
        self.defer_wgrad_allowed = os.environ.get("YVB200_DEFER_WGRAD", "1") != "0"
        self.early_zero = os.environ.get("YVB200_EARLY_ZERO", "1") != "0"
        self._join_queued = False
        # fused attention kernels (yv_attn_fwd / yv_attn_bwd): used when the caller does not need the attention
        # probabilities.  ``want_probs`` is set by BertEncoder.forward for the duration of a call (True when
        # output_all_attention_masks); None = a sub-module called on its own, which returns probabilities like the
        # reference and therefore takes the un-fused chain.  YVB200_FUSED_ATTN=0 forces the un-fused chain everywhere.
        self.fused_attention = os.environ.get("YVB200_FUSED_ATTN", "1") != "0"
        self.ln_split = os.environ.get("YVB200_LN_SPLIT", "1") != "0"
        self.want_probs: Optional[bool] = None
        self._tickets: Dict[int, torch.Tensor] = {}
        # weight-gradient sinks registered by a GradientExchange plan (step.py): {address of the weight's planes:
        # ([rows, cols] view of the flat exchange buffer, the parameters it covers)} -- see ``_dw_out``
        self.grad_sink: Dict[int, Tuple[torch.Tensor, tuple]] = {}
        # YVB200_PRIORITIES=1: the backward chain on high-priority streams, trailing weight-gradient work on low-priority
        # ones.  Off by default since round 2: on one GPU the two settings measure the same (9.32 vs 9.35 ms, three
        # alternations), with a gradient exchange the low-priority weight gradients finish so late that the collectives
        # cannot start before ~70 % of the step (2 GPUs 11.0 -> 10.74 ms, 8 GPUs 11.70 -> 11.30 ms with equal priorities)
        prio = os.environ.get("YVB200_PRIORITIES", "0") != "0"
        # All compute streams sit one level above the default priority so that the start-of-step weight refresh (its
        # own stream, default priority, 10k-CTA streaming launches) cannot hold the first forward kernels back: with
        # everything at one level the work distributor dispatched whole refresh chunks ahead of them (split_multi
        # "running alone" 0.31 ms per step against 0.09 ms).  YVB200_BASE_PRIORITY=0 restores the single level.
        base = int(os.environ.get("YVB200_BASE_PRIORITY", "-1"))
        self.helper_priority = base
        self.main_priority = base - 1 if prio else base
        self.branch_stream = torch.cuda.Stream(device=device, priority=self.main_priority)

    # ---- deferred weight-gradient work -------------------------------------------------------------
    def defer_join(self):
        """Called by a backward that left work running on a helper stream: all helper streams are joined into the
        stream that called ``backward()`` once, by an autograd-engine callback at the end of the backward pass."""
        if not self._join_queued:
            self._join_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._final_join)

    def _final_join(self):
        self._join_queued = False
        cur = torch.cuda.current_stream(self.device)
        capturing = torch.cuda.is_current_stream_capturing()
        for s_ in [self.branch_stream] + list(self._helpers.values()):
            if s_ == cur:
                continue
            if capturing:                                   # only streams that belong to this capture can be joined
                with torch.cuda.stream(s_):
                    if not torch.cuda.is_current_stream_capturing():
                        continue
            cur.wait_stream(s_)

    def fork(self, slot: int = 0) -> torch.cuda.Stream:
        """Stream for a short branch that is joined again at once (slot 0: the second attention direction or a weight
        gradient that is not deferred; slot 1: dV / dK inside an attention backward): same priority as the chain,
        never shared with trailing work."""
        cur = torch.cuda.current_stream(self.device)
        key = (cur.cuda_stream, slot)
        f = self._forks.get(key)
        if f is None:
            f = self._forks[key] = torch.cuda.Stream(device=self.device, priority=self.main_priority)
            self._helpers[(cur.cuda_stream, -1 - slot)] = f   # joined with the helpers at segment / pass ends
        return f

    def helper(self) -> torch.cuda.Stream:
        """A trailing-work stream paired with the current stream (created on first use).  With deferred weight gradients
        the helpers of a stream are handed out round-robin so that several small trailing GEMMs can be in flight."""
        cur = torch.cuda.current_stream(self.device)
        pool = self._helper_pool.get(cur.cuda_stream)
        if pool is None:
            n = self.helpers_per_stream
            pool = self._helper_pool[cur.cuda_stream] = [
                torch.cuda.Stream(device=self.device, priority=self.helper_priority) for _ in range(n)]
            for i, h in enumerate(pool):
                self._helpers[(cur.cuda_stream, i)] = h
            self._helper_next[cur.cuda_stream] = 0
        i = self._helper_next[cur.cuda_stream]
        self._helper_next[cur.cuda_stream] = (i + 1) % len(pool)
        return pool[i]

    # ---- zero-initialised accumulators of the backward pass (dgamma / dbeta / dbias reductions) ------------------
    ZERO_SLAB = 1 << 20            # floats: one 4 MB memset per training forward instead of ~70 memset nodes on the chain

    def begin_zero_slab(self):
        """Called at the start of a forward pass that will be differentiated: the backward nodes carve their
        zero-initialised reduction buffers out of this slab (``zeros``) instead of issuing a memset each."""
        self._zslab = torch.zeros(self.ZERO_SLAB, dtype=torch.float32, device=self.device)
        self._zoff = 0

    def zeros(self, rows: int, cols: int) -> torch.Tensor:
        n = rows * cols
        n_al = (n + 63) // 64 * 64
        slab = getattr(self, "_zslab", None)
        if slab is None or self._zoff + n_al > slab.numel():
            return torch.zeros(rows, cols, dtype=torch.float32, device=self.device)
        t = slab[self._zoff:self._zoff + n].view(rows, cols)
        self._zoff += n_al
        return t

    def attn_tickets(self, n: int) -> torch.Tensor:
        """Zero-initialised ticket counters for one fused attention backward launch.  The kernel leaves them zero, so a
        buffer is reusable by later launches on the SAME stream; launches on different streams may overlap and get
        their own buffer (keyed by the issuing stream)."""
        key = torch.cuda.current_stream(self.device).cuda_stream
        t = self._tickets.get(key)
        if t is None or t.numel() < n:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("yvb200: attention ticket buffers must exist before graph capture (run a warm-up step)")
            t = self._tickets[key] = torch.zeros(max(n, 4096), dtype=torch.int32, device=self.device)
        return t

    def set_precision(self, mode: str):
        self.passes = 3 if mode == "bf16x3" else 1
        self.attn_passes = self.passes

    def advance_rng(self):
        """Advance the step counter and give the next forward a fresh snapshot of {seed, counter}."""
        L.rng_advance(self.master_rng)
        self.rng = self.master_rng.clone()

    begin_training_forward = advance_rng

    def rng_state(self) -> torch.Tensor:
        return self.master_rng.clone()

    def set_rng_state(self, state: torch.Tensor):
        """Restore a state returned by ``rng_state``: the next training forward draws the masks that followed it."""
        self.master_rng.copy_(state)
        self.rng = self.master_rng.clone()


_RT: Dict[int, Runtime] = {}


def rt(device) -> Runtime:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("yvb200 ops need CUDA tensors (no CPU fallback on this path)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    r = _RT.get(idx)
    if r is None:
        L.load()
        r = _RT[idx] = Runtime(torch.device("cuda", idx))
    return r


# ----------------------------------------------------------------------------------------------------
# weight arena: all GEMM weights as bf16 hi/lo planes in a few flat buffers, refreshed in one launch
# ----------------------------------------------------------------------------------------------------
class _Entry:
    """One GEMM weight (or a row-wise concatenation of several) in the arena.  Parameters are held weakly so that a
    model that goes away does not stay alive -- and is not re-split every step -- because of its planes."""
    __slots__ = ("_prefs", "rows", "cols", "off", "chunk", "versions", "ptrs", "planes")

    @property
    def params(self):
        ps = tuple(r() for r in self._prefs)
        return None if any(p is None for p in ps) else ps

    @params.setter
    def params(self, ps):
        self._prefs = tuple(weakref.ref(p) for p in ps)


class _Chunk:
    def __init__(self, cap: int, device):
        self.cap = cap
        self.used = 0
        self.buf = torch.empty((2, cap), dtype=torch.bfloat16, device=device)
        self.entries: List[_Entry] = []
        self.table = None
        self.total_blocks = 0
        self.nseg = 0
        self.ready = None              # event of an in-flight refresh on the refresh stream (None: nothing pending)
        self.joined = set()            # streams that already wait for `ready`


class WeightArena:
    CHUNK = 24 * 1024 * 1024  # elements per plane and chunk (one yv_split_multi launch, ~30 us, per chunk)

    def __init__(self, device):
        self.device = device
        self.chunks: List[_Chunk] = []
        self.entries: Dict[Tuple[int, ...], _Entry] = {}
        self.always_stale = False      # bench: convert weights every step as real training would
        self.refresh_stream = None     # created on first overlapped refresh
        self.skip_next_refresh = False  # set by a caller that has just refreshed everything itself (GraphedStep)
        self.bias_cats: Dict[Tuple[int, ...], Tuple[torch.Tensor, tuple, list]] = {}   # fused Q|K|V biases
        self.generation = 0            # bumped whenever an entry appears or disappears (FusedAdamW's pointer tables)

    def _alloc(self, n: int) -> Tuple[_Chunk, int]:
        n_al = (n + 63) // 64 * 64
        for c in self.chunks:
            if c.used + n_al <= c.cap:
                off = c.used
                c.used += n_al
                return c, off
        c = _Chunk(max(self.CHUNK if self.chunks or n_al > 8 * 1024 * 1024 else 8 * 1024 * 1024, n_al), self.device)
        self.chunks.append(c)
        c.used = n_al
        return c, 0

    def _drop(self, key, e: _Entry):
        e.chunk.entries.remove(e)
        e.chunk.table = None
        del self.entries[key]
        self.generation += 1

    def prune(self):
        """Forget the entries whose parameters no longer exist (their arena space is not reused)."""
        for key, e in list(self.entries.items()):
            if e.params is None:
                self._drop(key, e)

    @staticmethod
    def _fresh(e: _Entry) -> bool:
        return len(e.versions) == len(e.params) and all(
            p._version == v and p.data_ptr() == a for p, v, a in zip(e.params, e.versions, e.ptrs))

    def _mark(self, e: _Entry):
        e.versions = tuple(p._version for p in e.params)
        e.ptrs = tuple(p.data_ptr() for p in e.params)

    def get(self, params: Tuple[torch.Tensor, ...]) -> Planes:
        """Planes of the row-wise concatenation of ``params`` (each [rows_i, cols], fp32)."""
        key = tuple(id(p) for p in params)
        e = self.entries.get(key)
        if e is not None:
            live = e.params
            if live is None or any(a is not b for a, b in zip(live, params)):   # stale entry of a freed parameter
                self._drop(key, e)
                e = None
        if e is None:
            cols = params[0].shape[1]
            if cols % 8:
                raise RuntimeError(f"yvb200: GEMM weight with in_features={cols} (must be a multiple of 8)")
            e = _Entry()
            e.params = tuple(params)
            e.rows = sum(p.shape[0] for p in params)
            e.cols = cols
            e.chunk, e.off = self._alloc(e.rows * cols)
            e.versions = e.ptrs = ()
            e.planes = Planes(e.chunk.buf, e.chunk.buf.data_ptr() + 2 * e.off, e.rows, cols, cols, e.chunk.cap)
            e.chunk.entries.append(e)
            e.chunk.table = None
            self.entries[key] = e
            self.generation += 1
        c = e.chunk
        if c.ready is not None:        # this chunk is being re-split on the refresh stream: order this stream after it
            cur = torch.cuda.current_stream(self.device)
            if cur.cuda_stream not in c.joined:
                cur.wait_event(c.ready)
                c.joined.add(cur.cuda_stream)
        if not self._fresh(e):
            self._refresh_entry(e)
        return e.planes

    def _refresh_entry(self, e: _Entry):
        r0 = 0
        for p in e.params:
            if not (p.is_contiguous() and p.dtype == torch.float32):
                raise RuntimeError("yvb200: weights must be contiguous fp32")
            dst = Planes(e.chunk.buf, e.planes.addr + 2 * r0 * e.cols, p.shape[0], e.cols, e.cols, e.chunk.cap)
            L.split_planes(p.detach(), dst)
            r0 += p.shape[0]
        self._mark(e)

    def join(self):
        """Order the current stream after every in-flight refresh (end of the forward pass / of a captured step)."""
        if self.refresh_stream is not None and any(c.ready is not None for c in self.chunks):
            torch.cuda.current_stream(self.device).wait_stream(self.refresh_stream)
        for c in self.chunks:
            c.ready = None
            c.joined = set()

    def bias_cat(self, biases: Tuple[torch.Tensor, ...]) -> torch.Tensor:
        """fp32 concatenation of the biases of a fused projection (Q|K|V), kept next to the weight planes and refreshed
        with them by ``refresh_all`` -- instead of one ``torch.cat`` launch in front of every projection GEMM."""
        key = tuple(id(b) for b in biases)
        ent = self.bias_cats.get(key)
        if ent is not None:
            live = tuple(r() for r in ent[1])
            if any(a is not b for a, b in zip(live, biases)):
                ent = None
        if ent is None:
            buf = torch.cat([b.detach() for b in biases])
            views, off = [], 0
            for b in biases:
                views.append(buf[off:off + b.numel()])
                off += b.numel()
            ent = self.bias_cats[key] = (buf, tuple(weakref.ref(b) for b in biases), views)
            self._bias_versions = None
        return ent[0]

    def _refresh_biases(self, force: bool):
        if not self.bias_cats:
            return
        dst, src, vers = [], [], []
        for key, (buf, refs, views) in list(self.bias_cats.items()):
            live = tuple(r() for r in refs)
            if any(b is None for b in live):
                del self.bias_cats[key]
                continue
            dst.extend(views)
            src.extend(b.detach() for b in live)
            vers.extend((b._version, b.data_ptr()) for b in live)
        if not force and getattr(self, "_bias_versions", None) == vers:
            return
        if dst:
            torch._foreach_copy_(dst, src)
        self._bias_versions = vers

    def refresh_for_forward(self, force: bool):
        """Called at the start of every model forward.  ``force`` (training mode or autograd enabled) re-splits every
        weight: optimizers that update through ``p.data`` (the reference's AdamW, vilbert/optimization.py:176-187) do
        not bump ``p._version``, so version tracking alone would keep serving the initial planes."""
        if self.skip_next_refresh:
            self.skip_next_refresh = False
            return
        self.refresh_all(force=force)

    def refresh_all(self, force: bool = False, overlap: bool = False):
        """Re-split every chunk that holds a stale entry with one ``yv_split_multi`` launch.  With ``overlap`` the
        chunks after the first (the arena is filled in forward order, so they hold the later layers) are converted on
        a separate stream while the forward pass starts; ``get`` orders each consumer stream after its chunk."""
        force = force or self.always_stale
        self.prune()
        self._refresh_biases(force)
        cur = torch.cuda.current_stream(self.device)
        if overlap and len(self.chunks) > 1:
            if self.refresh_stream is None:
                self.refresh_stream = torch.cuda.Stream(device=self.device)
            self.refresh_stream.wait_stream(cur)
        for ci, c in enumerate(self.chunks):
            if not c.entries:
                continue
            if not force and all(self._fresh(e) for e in c.entries):
                continue
            ptrs = tuple(p.data_ptr() for e in c.entries for p in e.params)
            if c.table is None or c.table[1] != ptrs:
                rows = []
                blk = 0
                for e in c.entries:
                    off = e.off
                    for p in e.params:
                        if not (p.is_contiguous() and p.dtype == torch.float32):
                            raise RuntimeError("yvb200: weights must be contiguous fp32")
                        n = p.numel()
                        rows.append([p.data_ptr(), off, n, blk])
                        blk += (n + 2047) // 2048
                        off += n
                t = torch.tensor(rows, dtype=torch.int64).to(self.device, non_blocking=False)
                c.table = (t, ptrs)
                c.total_blocks = blk
                c.nseg = len(rows)
            if overlap and ci > 0 and self.refresh_stream is not None:
                with torch.cuda.stream(self.refresh_stream):
                    L.split_multi(c.table[0], c.nseg, c.total_blocks, c.buf, c.cap)
                    c.ready = torch.cuda.Event()
                    c.ready.record(self.refresh_stream)
                c.joined = set()
            else:
                L.split_multi(c.table[0], c.nseg, c.total_blocks, c.buf, c.cap)
            for e in c.entries:
                self._mark(e)


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def _c2d(x: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> contiguous 2-D view (copying only when the input is not contiguous)."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.dim() == 2 and x.stride(1) == 1 and x.stride(0) >= x.shape[1]:
        return x                       # row-strided 2-D views (e.g. hidden[:, 0]) are consumed in place
    if x.dim() > 2 and not x.is_contiguous():
        try:
            x2 = x.view(-1, x.shape[-1])            # padded-row views (logits with ld > cols) stay in place
            if x2.stride(1) == 1:
                return x2
        except RuntimeError:
            pass
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.is_contiguous() else x2.contiguous()


def attach_planes(t: torch.Tensor, p: Planes):
    t._yv_planes = (p, t._version, t.data_ptr())


def planes_of(x: torch.Tensor, x2: torch.Tensor) -> Planes:
    """bf16 hi/lo planes of the 2-D view ``x2`` of ``x`` -- reuses the producer's planes when present."""
    tag = getattr(x, "_yv_planes", None)
    if tag is not None and tag[1] == x._version and tag[2] == x.data_ptr() and tag[0].rows == x2.shape[0] \
            and tag[0].cols == x2.shape[1]:
        return tag[0]
    return L.split_planes(x2)


def _f32(*shape, device):
    return torch.empty(shape, dtype=torch.float32, device=device)


class _EarlyOut:
    """fp32 [M, N] output of a GEMM contracting over K.  If that launch will split K (partial sums are reduced into a
    zero-filled output) the buffer is zero-filled NOW on a forked stream, i.e. concurrently with whatever the caller
    issues before the GEMM; ``ready()`` orders the issuing stream after the fill right before the launch.  This takes
    the memset node that would otherwise sit in front of every split-K kernel off the dependency chain."""
    __slots__ = ("t", "zeroed", "_z", "_dev")

    def __init__(self, r: "Runtime", M: int, N: int, K: int, device):
        self.t = _f32(M, N, device=device)
        self._dev = device
        self._z = None
        self.zeroed = r.early_zero and L.will_split(M, N, K)
        if self.zeroed:
            if r.concurrent:
                cur = torch.cuda.current_stream(device)
                self._z = r.fork(2)
                self._z.wait_stream(cur)
                with torch.cuda.stream(self._z):
                    self.t.zero_()
            else:
                self.t.zero_()

    def ready(self) -> bool:
        if self._z is not None:
            torch.cuda.current_stream(self._dev).wait_stream(self._z)
            self._z = None
        return self.zeroed


def _dw_out(r: "Runtime", wp: "Planes", N: int, K: int, device) -> torch.Tensor:
    """Where the gradient of the weight(s) behind the planes ``wp`` is written: straight into its slot of the data-parallel
    exchange buffer when a GradientExchange plan registered one (autograd then adopts that view as ``.grad`` and the
    exchange has nothing to copy), else a fresh tensor.  Never when a parameter already holds a gradient (accumulation
    would then add the buffer to itself)."""
    if r.grad_sink:
        slot = r.grad_sink.get(wp.addr)
        if slot is not None:
            t, params = slot
            if t.shape[0] == N and t.shape[1] == K and all(q.grad is None for q in params):
                return t.view(N, K)                  # (a fresh view object: autograd only adopts unshared tensors)
    return _f32(N, K, device=device)


def _keep_for(side: torch.cuda.Stream, *objs):
    """Tensors consumed (or produced) by work left running on ``side``: the caching allocator must not hand their
    memory to a later allocation of the issuing stream before that work has run."""
    for o in objs:
        if o is None:
            continue
        t = o.keep if isinstance(o, Planes) else o
        t.record_stream(side)


def _ln_bwd_chain(r: Runtime, dz2, s, gamma, stats, ds, dsp, acc, M, C, drop_p, site, rng):
    """LayerNorm backward of a dense -> dropout -> (+residual) -> LN block, split in two: the part the backward chain
    waits for (ds and its dropped planes) is launched now; the returned closure issues the column reductions
    (dgamma, dbeta -> acc[0], acc[1]; bias gradient of the dense layer -> acc[2]) and is meant to run on the trailing
    stream of the caller, next to the weight gradient.  YVB200_LN_SPLIT=0 keeps the single fused kernel."""
    if not r.ln_split:
        L.layernorm_bwd(dz2, s, gamma, stats, ds, dsp, acc[0], acc[1], M, C, pre_drop_p=drop_p, pre_drop_site=site, rng=rng,
                        dbias=acc[2])
        return lambda: None
    L.layernorm_bwd_dx(dz2, s, gamma, stats, ds, dsp, M, C, pre_drop_p=drop_p, pre_drop_site=site, rng=rng)

    def cols():
        L.layernorm_bwd_cols(dz2, s, stats, acc[0], acc[1], M, C)
        L.colsum_planes(dsp, acc[2], accumulate=True)
    return cols


def _linear_bwd(r: Runtime, dp: Planes, xp: Optional[Planes], wp: Planes, M: int, N: int, K: int, device,
                need_dx: bool, need_dw: bool, dx_residual: Optional[torch.Tensor] = None,
                db: Optional[torch.Tensor] = None, defer: bool = False, dx_out: Optional["_EarlyOut"] = None,
                side_first=None, side_keep=()):
    """dx = dp . W ;  dW = dp^T . x ;  db = colsum(dp)   with dp [M,N], W [N,K], x [M,K].
    ``db`` may arrive precomputed (fused into the kernel that produced ``dp``).  With ``defer`` the weight / bias
    gradient is left running on the helper stream (joined at the end of the backward pass, see Runtime.defer_join)."""
    dx = dW = None
    have_db = db is not None
    dx_zeroed = False
    if need_dx:
        if dx_out is not None:
            dx, dx_zeroed = dx_out.t, dx_out.ready()
        else:
            dx = _f32(M, K, device=device)
    if need_dw:
        dW = _dw_out(r, wp, N, K, device)
        if not have_db:
            db = _f32(N, device=device)
    fork = r.concurrent and need_dx and need_dw
    if fork:                                    # wgrad + bias grad on the helper stream, dgrad on this one
        cur = torch.cuda.current_stream(device)
        trailing = defer and r.defer_wgrad
        side = r.helper() if trailing else r.fork()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if side_first is not None:
                side_first()
            L.gemm(N, K, M, L.op_of(dp, True), L.op_of(xp, True), passes=r.passes, out32=dW, ld_out=K)
            if not have_db:
                L.colsum_planes(dp, db)
        L.gemm(M, K, N, L.op_of(dp), L.op_of(wp, True), passes=r.passes, out32=dx, ld_out=K, residual=dx_residual,
               out32_zeroed=dx_zeroed)
        if trailing:
            _keep_for(side, dp, xp, dW, None if have_db else db, *side_keep)
            r.defer_join()
        else:
            cur.wait_stream(side)
        return dx, dW, db
    if side_first is not None:
        side_first()
    if need_dx:
        L.gemm(M, K, N, L.op_of(dp), L.op_of(wp, True), passes=r.passes, out32=dx, ld_out=K, residual=dx_residual,
               out32_zeroed=dx_zeroed)
    if need_dw:
        L.gemm(N, K, M, L.op_of(dp, True), L.op_of(xp, True), passes=r.passes, out32=dW, ld_out=K)
        if not have_db:
            L.colsum_planes(dp, db)
    return dx, dW, db


# ----------------------------------------------------------------------------------------------------
# DenseAct
# ----------------------------------------------------------------------------------------------------
class DenseActFn(Function):
    @staticmethod
    def forward(ctx, x, W, b, spec):
        r = rt(x.device)
        x2 = _c2d(x)
        M, K = x2.shape
        N = W.shape[0]
        xp = planes_of(x, x2)
        wp = r.arena.get((W,))
        act = spec.act
        need = any(ctx.needs_input_grad)
        # rows of wide, oddly sized outputs (30522 / 1601 logits) are padded to 8 floats so every store is a float4
        ldN = (N + 7) // 8 * 8 if getattr(spec, "pad_out", False) else N
        y = _f32(M, ldN, device=x.device)
        yp = Planes.empty(M, N, x.device) if spec.want_planes else None
        pre = _f32(M, N, device=x.device) if (act == L.ACT_GELU and need) else None
        L.gemm(M, N, K, L.op_of(xp), L.op_of(wp), passes=r.passes, bias=b, act=act, aux_out=pre, out32=y, ld_out=ldN,
               out_planes=yp.ptr() if yp else None, ld_pl=yp.ld if yp else 0,
               pl_plane_stride=yp.plane_stride if yp else 0)
        ctx.r, ctx.xp, ctx.wp, ctx.act, ctx.dims = r, xp, wp, act, (M, N, K)
        ctx.aux = pre if act == L.ACT_GELU else (y if act == L.ACT_RELU else None)
        ctx.xshape = x.shape
        spec.out_planes = yp
        if ldN != N:
            return y[:, :N].view(*x.shape[:-1], N)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        r = ctx.r
        M, N, K = ctx.dims
        dy2 = _c2d(dy)
        dp = Planes.empty(M, N, dy.device)
        db = _f32(N, device=dy.device) if ctx.needs_input_grad[2] else None
        L.act_bwd_split(dy2, ctx.aux, ctx.act, dp, db)
        dx, dW, db = _linear_bwd(r, dp, ctx.xp, ctx.wp, M, N, K, dy.device, ctx.needs_input_grad[0],
                                 ctx.needs_input_grad[1], db=db)
        return (dx.view(ctx.xshape) if dx is not None else None), dW, db, None


def dense_act(x, W, b, act: int, want_planes: bool = True, pad_out: bool = False):
    spec = types.SimpleNamespace(act=act, want_planes=want_planes, out_planes=None, pad_out=pad_out)
    y = DenseActFn.apply(x, W, b, spec)
    if spec.out_planes is not None:
        attach_planes(y, spec.out_planes)
    return y


# ----------------------------------------------------------------------------------------------------
# DenseResLN
# ----------------------------------------------------------------------------------------------------
class DenseResLNFn(Function):
    @staticmethod
    def forward(ctx, x, res, W, b, gamma, beta, spec):
        r = rt(x.device)
        ctx.rng = rng = r.rng
        x2 = _c2d(x)
        res2 = _c2d(res)
        M, K = x2.shape
        N = W.shape[0]
        xp = planes_of(x, x2)
        wp = r.arena.get((W,))
        s = _f32(M, N, device=x.device)
        L.gemm(M, N, K, L.op_of(xp), L.op_of(wp), passes=r.passes, bias=b, residual=res2, out32=s, ld_out=N,
               drop_p=spec.drop_p, drop_site=spec.site, rng=rng)
        z = _f32(M, N, device=x.device)
        zp = Planes.empty(M, N, x.device)
        stats = _f32(M, 2, device=x.device)
        L.layernorm_fwd(s, gamma, beta, LN_EPS, z, zp, stats, M, N)
        ctx.r, ctx.xp, ctx.wp, ctx.dims, ctx.spec = r, xp, wp, (M, N, K), spec
        ctx.save_for_backward(s, stats, gamma)
        ctx.xshape, ctx.rshape = x.shape, res.shape
        spec.out_planes = zp
        return z.view(res.shape)

    @staticmethod
    def backward(ctx, dz):
        r, spec = ctx.r, ctx.spec
        M, N, K = ctx.dims
        s, stats, gamma = ctx.saved_tensors
        dz2 = _c2d(dz)
        dev = dz.device
        dx_out = _EarlyOut(r, M, K, N, dev) if ctx.needs_input_grad[0] else None
        ds = _f32(M, N, device=dev)
        dsp = Planes.empty(M, N, dev)
        acc = r.zeros(3, N)                                               # dgamma, dbeta, dbias
        cols = _ln_bwd_chain(r, dz2, s, gamma, stats, ds, dsp, acc, M, N, spec.drop_p, spec.site, ctx.rng)
        dx, dW, db = _linear_bwd(r, dsp, ctx.xp, ctx.wp, M, N, K, dev, ctx.needs_input_grad[0], ctx.needs_input_grad[2],
                                 db=acc[2], defer=True, dx_out=dx_out, side_first=cols, side_keep=(dz2, s, stats, acc))
        return (dx.view(ctx.xshape) if dx is not None else None), ds.view(ctx.rshape), dW, db, acc[0], acc[1], None


def dense_res_ln(x, res, W, b, gamma, beta, drop_p: float, site: int):
    spec = types.SimpleNamespace(drop_p=float(drop_p), site=site, out_planes=None)
    z = DenseResLNFn.apply(x, res, W, b, gamma, beta, spec)
    attach_planes(z, spec.out_planes)
    return z


# ----------------------------------------------------------------------------------------------------
# DenseActLN (prediction-head transforms)
# ----------------------------------------------------------------------------------------------------
class DenseActLNFn(Function):
    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, spec):
        r = rt(x.device)
        x2 = _c2d(x)
        M, K = x2.shape
        N = W.shape[0]
        xp = planes_of(x, x2)
        wp = r.arena.get((W,))
        pre = _f32(M, N, device=x.device)
        g = _f32(M, N, device=x.device)
        L.gemm(M, N, K, L.op_of(xp), L.op_of(wp), passes=r.passes, bias=b, act=spec.act, aux_out=pre, out32=g, ld_out=N)
        z = _f32(M, N, device=x.device)
        zp = Planes.empty(M, N, x.device)
        stats = _f32(M, 2, device=x.device)
        L.layernorm_fwd(g, gamma, beta, LN_EPS, z, zp, stats, M, N)
        ctx.r, ctx.xp, ctx.wp, ctx.dims, ctx.act = r, xp, wp, (M, N, K), spec.act
        ctx.save_for_backward(pre, g, stats, gamma)
        ctx.xshape = x.shape
        spec.out_planes = zp
        return z.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dz):
        r = ctx.r
        M, N, K = ctx.dims
        pre, g, stats, gamma = ctx.saved_tensors
        dev = dz.device
        dg = _f32(M, N, device=dev)
        acc = r.zeros(2, N)
        dgamma, dbeta = acc[0], acc[1]
        L.layernorm_bwd(_c2d(dz), g, gamma, stats, dg, None, dgamma, dbeta, M, N)
        dp = Planes.empty(M, N, dev)
        aux = pre if ctx.act == L.ACT_GELU else g
        db = _f32(N, device=dev)
        L.act_bwd_split(dg, aux, ctx.act, dp, db)
        dx, dW, db = _linear_bwd(r, dp, ctx.xp, ctx.wp, M, N, K, dev, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                 db=db)
        return (dx.view(ctx.xshape) if dx is not None else None), dW, db, dgamma, dbeta, None


def dense_act_ln(x, W, b, gamma, beta, act: int):
    spec = types.SimpleNamespace(act=act, out_planes=None)
    z = DenseActLNFn.apply(x, W, b, gamma, beta, spec)
    attach_planes(z, spec.out_planes)
    return z


# ----------------------------------------------------------------------------------------------------
# attention core (shared by self- and bi-attention); q/k/v are "head views" into projection buffers
# ----------------------------------------------------------------------------------------------------
class HeadView:
    """Columns [off, off + heads*dh) of a plane pair [pairs*S, ld] seen as [pairs, heads, S, dh]."""
    __slots__ = ("p", "off", "S")

    def __init__(self, p: Planes, off: int, S: int):
        self.p, self.off, self.S = p, off, S

    def operand(self, pairs: int, heads: int, dh: int, mn_major: bool):
        return L.operand(self.p.ptr(self.off), dh, self.S, self.p.ld, self.p.plane_stride, mn_major, heads, dh, pairs,
                         self.S * self.p.ld)

    def view(self):
        return L.head_view(self.p, self.off, self.S)


def _score_operand(p: Planes, Tq: int, Tk: int, ldS: int, pairs: int, heads: int, mn_major: bool):
    return L.operand(p.ptr(), Tk, Tq, ldS, p.plane_stride, mn_major, heads, Tq * ldS, pairs, heads * Tq * ldS)


class _AttnSaved:
    """What an attention forward leaves for its backward: the row log-sum-exp (fused kernels, probabilities are
    recomputed) or the probabilities and their dropped bf16 planes (un-fused chain)."""
    __slots__ = ("fused", "lse", "mask", "ctxp", "P", "Pp")

    def __init__(self, fused, lse=None, mask=None, ctxp=None, P=None, Pp=None):
        self.fused, self.lse, self.mask, self.ctxp, self.P, self.Pp = fused, lse, mask, ctxp, P, Pp

    def keep_for(self, side):
        _keep_for(side, self.lse, self.P, self.Pp)


def _use_fused_attention(r: Runtime, dh: int) -> bool:
    return r.fused_attention and r.want_probs is False and L.attn_supported(dh, r.attn_passes)


def _attn_fwd(r: Runtime, q: HeadView, k: HeadView, v: HeadView, mask: torch.Tensor, pairs: int, heads: int, dh: int,
              drop_p: float, site: int, ctxp: Planes, ctx32: Optional[torch.Tensor], rng=None) -> _AttnSaved:
    """softmax(q k^T / sqrt(dh) + mask) v -> ``ctxp`` (merged heads): one fused launch, or GEMM -> softmax -> GEMM when
    the probabilities themselves are wanted (or the head size has no fused kernel)."""
    if _use_fused_attention(r, dh):
        lse = _f32(pairs * heads * q.S, device=mask.device)
        L.attn_fwd(q.view(), k.view(), v.view(), mask, pairs, heads, dh, 1.0 / math.sqrt(dh), ctxp, ctx32, lse,
                   passes=r.attn_passes, drop_p=drop_p, drop_site=site, rng=rng)
        return _AttnSaved(True, lse=lse, mask=mask, ctxp=ctxp)
    P, Pp = _attn_fwd_unfused(r, q, k, v, mask, pairs, heads, dh, drop_p, site, ctxp, ctx32, rng)
    return _AttnSaved(False, P=P, Pp=Pp)


def _attn_bwd(r: Runtime, dOp: Planes, q: HeadView, k: HeadView, v: HeadView, saved: _AttnSaved, pairs: int, heads: int,
              dh: int, drop_p: float, site: int, dq: HeadView, dk: HeadView, dv: HeadView, fork_ok: bool = True, rng=None):
    if saved.fused:
        dev = dOp.keep.device
        ws = torch.empty(L.attn_bwd_workspace_bytes(pairs, heads, dh, q.S, k.S), dtype=torch.uint8, device=dev)
        L.attn_bwd(q.view(), k.view(), v.view(), L.head_view(dOp, 0, q.S), L.head_view(saved.ctxp, 0, q.S), saved.mask,
                   saved.lse, pairs, heads, dh, 1.0 / math.sqrt(dh), dq.view(), dk.view(), dv.view(), ws,
                   r.attn_tickets(pairs * heads), passes=r.attn_passes, drop_p=drop_p, drop_site=site, rng=rng)
        return
    _attn_bwd_unfused(r, dOp, q, k, v, saved.P, saved.Pp, pairs, heads, dh, drop_p, site, dq, dk, dv, fork_ok, rng)


def _attn_fwd_unfused(r: Runtime, q: HeadView, k: HeadView, v: HeadView, mask: torch.Tensor, pairs: int, heads: int, dh: int,
                      drop_p: float, site: int, ctxp: Planes, ctx32: Optional[torch.Tensor], rng=None):
    Tq, Tk = q.S, k.S
    ldS = (Tk + 7) // 8 * 8
    dev = mask.device
    S = _f32(pairs, heads, Tq, ldS, device=dev)
    L.gemm(Tq, Tk, dh, q.operand(pairs, heads, dh, False), k.operand(pairs, heads, dh, False), passes=r.attn_passes,
           out32=S, ld_out=ldS, out_sb0=Tq * ldS, out_sb1=heads * Tq * ldS)
    rows = pairs * heads * Tq
    Pp = Planes.empty(rows, Tk, dev, ld=ldS)
    L.softmax_fwd(S, ldS, mask, rows, Tk, heads * Tq, 1.0 / math.sqrt(dh), Pp, drop_p, site, rng)
    Hout = heads * dh
    L.gemm(Tq, dh, Tk, _score_operand(Pp, Tq, Tk, ldS, pairs, heads, False), v.operand(pairs, heads, dh, True),
           passes=r.attn_passes, out32=ctx32, ld_out=Hout, out_sb0=dh, out_sb1=Tq * Hout, out_planes=ctxp.ptr(),
           ld_pl=ctxp.ld, pl_sb0=dh, pl_sb1=Tq * ctxp.ld, pl_plane_stride=ctxp.plane_stride)
    return S, Pp


def _attn_bwd_unfused(r: Runtime, dOp: Planes, q: HeadView, k: HeadView, v: HeadView, P: torch.Tensor, Pp: Planes,
                      pairs: int, heads: int, dh: int, drop_p: float, site: int, dq: HeadView, dk: HeadView, dv: HeadView,
                      fork_ok: bool = True, rng=None):
    """dP -> softmax' -> dQ on the issuing stream; dV (needs only P and dO) and dK (needs dS) on a forked stream when
    ``fork_ok`` (callers that already run this function on a forked stream pass False)."""
    Tq, Tk = q.S, k.S
    ldS = P.shape[-1]
    dev = P.device
    rows = pairs * heads * Tq
    dO = HeadView(dOp, 0, Tq)
    cur = torch.cuda.current_stream(dev)
    side = r.fork(1) if (r.concurrent and fork_ok) else cur

    def out_view(hv: HeadView):
        return dict(out_planes=hv.p.ptr(hv.off), ld_pl=hv.p.ld, pl_sb0=dh, pl_sb1=hv.S * hv.p.ld,
                    pl_plane_stride=hv.p.plane_stride)

    if side is not cur:
        side.wait_stream(cur)
    with torch.cuda.stream(side):
        # dV = Pd^T . dO     [Tk, dh], contraction over Tq
        L.gemm(Tk, dh, Tq, _score_operand(Pp, Tq, Tk, ldS, pairs, heads, True), dO.operand(pairs, heads, dh, True),
               passes=r.attn_passes, **out_view(dv))
    # dPd = dO . V^T
    dPd = _f32(pairs, heads, Tq, ldS, device=dev)
    L.gemm(Tq, Tk, dh, dO.operand(pairs, heads, dh, False), v.operand(pairs, heads, dh, False), passes=r.attn_passes,
           out32=dPd, ld_out=ldS, out_sb0=Tq * ldS, out_sb1=heads * Tq * ldS)
    dSp = Planes.empty(rows, Tk, dev, ld=ldS)
    L.softmax_bwd(P, dPd, ldS, rows, Tk, 1.0 / math.sqrt(dh), dSp, drop_p, site, rng)
    if side is not cur:
        side.wait_stream(cur)
    with torch.cuda.stream(side):
        # dK = dS^T . Q      [Tk, dh], contraction over Tq
        L.gemm(Tk, dh, Tq, _score_operand(dSp, Tq, Tk, ldS, pairs, heads, True), q.operand(pairs, heads, dh, True),
               passes=r.attn_passes, **out_view(dk))
    # dQ = dS . K        [Tq, dh], contraction over Tk
    L.gemm(Tq, dh, Tk, _score_operand(dSp, Tq, Tk, ldS, pairs, heads, False), k.operand(pairs, heads, dh, True),
           passes=r.attn_passes, **out_view(dq))
    if side is not cur:
        cur.wait_stream(side)


def _mask2d(mask: torch.Tensor, pairs: int, Tk: int) -> torch.Tensor:
    """The reference passes the additive mask as [N,1,1,S] (vilbert/vilbert.py:1268-1287)."""
    if mask.numel() != pairs * Tk:
        raise RuntimeError(f"yvb200: attention mask of shape {tuple(mask.shape)} is not [N,1,1,{Tk}]")
    m = mask.reshape(pairs, Tk)
    if m.dtype != torch.float32:
        m = m.float()
    return m.contiguous()


def _cat_bias(r: "Runtime", bs):
    return r.arena.bias_cat(tuple(bs))


class SelfAttentionFn(Function):
    @staticmethod
    def forward(ctx, x, mask, Wq, bq, Wk, bk, Wv, bv, spec):
        r = rt(x.device)
        ctx.rng = rng = r.rng
        pairs, S, K = x.shape
        H = Wq.shape[0]
        heads = spec.heads
        dh = H // heads
        x2 = _c2d(x)
        M = pairs * S
        xp = planes_of(x, x2)
        wp = r.arena.get((Wq, Wk, Wv))
        m2 = _mask2d(mask, pairs, S)
        qkv = Planes.empty(M, 3 * H, x.device)
        L.gemm(M, 3 * H, K, L.op_of(xp), L.op_of(wp), passes=r.passes, bias=_cat_bias(r, (bq, bk, bv)),
               out_planes=qkv.ptr(), ld_pl=qkv.ld, pl_plane_stride=qkv.plane_stride)
        c32 = _f32(M, H, device=x.device)
        cp = Planes.empty(M, H, x.device)
        q, k, v = HeadView(qkv, 0, S), HeadView(qkv, H, S), HeadView(qkv, 2 * H, S)
        att = _attn_fwd(r, q, k, v, m2, pairs, heads, dh, spec.drop_p, spec.site, cp, c32, rng)
        ctx.r, ctx.xp, ctx.wp, ctx.qkv, ctx.att, ctx.spec = r, xp, wp, qkv, att, spec
        ctx.dims = (pairs, S, K, H, heads, dh)
        spec.out_planes = cp
        spec.probs = att.P
        return c32.view(pairs, S, H)

    @staticmethod
    def backward(ctx, dc):
        r, spec = ctx.r, ctx.spec
        pairs, S, K, H, heads, dh = ctx.dims
        M = pairs * S
        dev = dc.device
        dOp = L.split_planes(_c2d(dc))
        dqkv = Planes.empty(M, 3 * H, dev)
        qkv = ctx.qkv
        q, k, v = HeadView(qkv, 0, S), HeadView(qkv, H, S), HeadView(qkv, 2 * H, S)
        _attn_bwd(r, dOp, q, k, v, ctx.att, pairs, heads, dh, spec.drop_p, spec.site,
                  HeadView(dqkv, 0, S), HeadView(dqkv, H, S), HeadView(dqkv, 2 * H, S), rng=ctx.rng)
        dx, dW, db = _linear_bwd(r, dqkv, ctx.xp, ctx.wp, M, 3 * H, K, dev, ctx.needs_input_grad[0], True, defer=True)
        return ((dx.view(pairs, S, K) if dx is not None else None), None,
                dW[:H], db[:H], dW[H:2 * H], db[H:2 * H], dW[2 * H:], db[2 * H:], None)


def self_attention(x, mask, Wq, bq, Wk, bk, Wv, bv, heads: int, drop_p: float, site: int):
    spec = types.SimpleNamespace(heads=heads, drop_p=float(drop_p), site=site, out_planes=None, probs=None)
    c = SelfAttentionFn.apply(x, mask, Wq, bq, Wk, bk, Wv, bv, spec)
    attach_planes(c, spec.out_planes)
    return c, spec.probs


class AttnBlockFn(Function):
    """Self-attention + output projection + residual + LayerNorm as ONE autograd node: the reference's
    ``Bert(Image)Attention`` = ``Bert(Image)SelfOutput(Bert(Image)SelfAttention(x), x)`` (vilbert/vilbert.py:328-337,
    :456-464).  Inside one node the fp32 copies of the context and of its gradient are never written, the out-proj
    dgrad emits the bf16 planes the attention backward consumes (no split pass), and the residual gradient is added in
    the epilogue of the QKV dgrad instead of by a separate autograd accumulation kernel."""

    @staticmethod
    def forward(ctx, x, mask, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, beta, spec):
        r = rt(x.device)
        ctx.rng = rng = r.rng
        pairs, S, K = x.shape
        H = Wq.shape[0]
        heads = spec.heads
        dh = H // heads
        x2 = _c2d(x)
        if x2.stride(0) != K:
            x2 = x2.contiguous()
        M = pairs * S
        dev = x.device
        xp = planes_of(x, x2)
        wqkv = r.arena.get((Wq, Wk, Wv))
        wo = r.arena.get((Wo,))
        m2 = _mask2d(mask, pairs, S)
        s_out = _EarlyOut(r, M, H, H, dev)
        qkv = Planes.empty(M, 3 * H, dev)
        L.gemm(M, 3 * H, K, L.op_of(xp), L.op_of(wqkv), passes=r.passes, bias=_cat_bias(r, (bq, bk, bv)),
               out_planes=qkv.ptr(), ld_pl=qkv.ld, pl_plane_stride=qkv.plane_stride)
        cp = Planes.empty(M, H, dev)
        q, k, v = HeadView(qkv, 0, S), HeadView(qkv, H, S), HeadView(qkv, 2 * H, S)
        att = _attn_fwd(r, q, k, v, m2, pairs, heads, dh, spec.drop_p, spec.site, cp, None, rng)
        s = s_out.t
        L.gemm(M, H, H, L.op_of(cp), L.op_of(wo), passes=r.passes, bias=bo, residual=x2, out32=s, ld_out=H,
               drop_p=spec.out_drop_p, drop_site=spec.out_site, rng=rng, out32_zeroed=s_out.ready())
        z = _f32(M, H, device=dev)
        zp = Planes.empty(M, H, dev)
        stats = _f32(M, 2, device=dev)
        L.layernorm_fwd(s, gamma, beta, LN_EPS, z, zp, stats, M, H)
        ctx.r, ctx.spec = r, spec
        ctx.keep = (xp, wqkv, wo, qkv, att, cp)
        ctx.dims = (pairs, S, K, H, heads, dh)
        ctx.save_for_backward(s, stats, gamma)
        spec.out_planes = zp
        spec.probs = att.P
        return z.view(pairs, S, H)

    @staticmethod
    def backward(ctx, dz):
        r, spec = ctx.r, ctx.spec
        pairs, S, K, H, heads, dh = ctx.dims
        xp, wqkv, wo, qkv, att, cp = ctx.keep
        s, stats, gamma = ctx.saved_tensors
        M = pairs * S
        dev = dz.device
        dx_out = _EarlyOut(r, M, K, 3 * H, dev) if ctx.needs_input_grad[0] else None
        ds = _f32(M, H, device=dev)
        dsp = Planes.empty(M, H, dev)
        acc = r.zeros(3, H)                                               # dgamma, dbeta, d(out bias)
        dz2 = _c2d(dz)
        cols = _ln_bwd_chain(r, dz2, s, gamma, stats, ds, dsp, acc, M, H, spec.out_drop_p, spec.out_site, ctx.rng)
        dWo = _dw_out(r, wo, H, H, dev)
        dOp = Planes.empty(M, H, dev)
        cur = torch.cuda.current_stream(dev)
        side = (r.helper() if r.defer_wgrad else r.fork()) if r.concurrent else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                     # LN column sums, wgrad of the output projection
            cols()
            L.gemm(H, H, M, L.op_of(dsp, True), L.op_of(cp, True), passes=r.passes, out32=dWo, ld_out=H)
        # dgrad of the output projection straight into the planes the attention backward reads
        L.gemm(M, H, H, L.op_of(dsp), L.op_of(wo, True), passes=r.passes, out_planes=dOp.ptr(), ld_pl=dOp.ld,
               pl_plane_stride=dOp.plane_stride)
        dqkv = Planes.empty(M, 3 * H, dev)
        q, k, v = HeadView(qkv, 0, S), HeadView(qkv, H, S), HeadView(qkv, 2 * H, S)
        _attn_bwd(r, dOp, q, k, v, att, pairs, heads, dh, spec.drop_p, spec.site,
                  HeadView(dqkv, 0, S), HeadView(dqkv, H, S), HeadView(dqkv, 2 * H, S), rng=ctx.rng)
        if side is not cur and r.defer_wgrad:
            _keep_for(side, dsp, cp, dWo, dz2, s, stats, acc)
            r.defer_join()
        else:
            cur.wait_stream(side)
        need_dx = ctx.needs_input_grad[0]
        dx, dW, db = _linear_bwd(r, dqkv, xp, wqkv, M, 3 * H, K, dev, need_dx, True, dx_residual=ds, defer=True,
                                 dx_out=dx_out)
        if not need_dx:
            dx = None
        return ((dx.view(pairs, S, K) if dx is not None else None), None,
                dW[:H], db[:H], dW[H:2 * H], db[H:2 * H], dW[2 * H:], db[2 * H:], dWo, acc[2], acc[0], acc[1], None)


def attention_block(x, mask, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, beta, heads: int, drop_p: float, site: int,
                    out_drop_p: float, out_site: int):
    """LN(dropout(out_proj(attention(x))) + x); returns (output, attention probabilities)."""
    spec = types.SimpleNamespace(heads=heads, drop_p=float(drop_p), site=site, out_drop_p=float(out_drop_p),
                                 out_site=out_site, out_planes=None, probs=None)
    z = AttnBlockFn.apply(x, mask, Wq, bq, Wk, bk, Wv, bv, Wo, bo, gamma, beta, spec)
    attach_planes(z, spec.out_planes)
    return z, spec.probs


class FFNFn(Function):
    """Feed-forward block as ONE autograd node: ``LN(dropout(W2 gelu(W1 x + b1) + b2) + x)`` -- the reference's
    ``Bert(Image)Output(Bert(Image)Intermediate(x), x)`` (vilbert/vilbert.py:340-368, :467-495, used at :380-381,
    :507-508, :674-677).  The hidden activation only exists as bf16 planes (plus its fp32 pre-activation for GELU'),
    the dgrad of W2 multiplies by GELU'(pre) in its epilogue and emits planes (no activation-backward pass), and the
    residual gradient rides in the epilogue of the W1 dgrad."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, gamma, beta, spec):
        r = rt(x.device)
        ctx.rng = rng = r.rng
        x2 = _c2d(x)
        M, H = x2.shape
        if x2.stride(0) != H:
            x2 = x2.contiguous()
        FF = W1.shape[0]
        dev = x.device
        xp = planes_of(x, x2)
        w1 = r.arena.get((W1,))
        w2 = r.arena.get((W2,))
        s_out = _EarlyOut(r, M, H, FF, dev)
        pre = _f32(M, FF, device=dev) if any(ctx.needs_input_grad) else None
        hp = Planes.empty(M, FF, dev)
        L.gemm(M, FF, H, L.op_of(xp), L.op_of(w1), passes=r.passes, bias=b1, act=L.ACT_GELU, aux_out=pre, ld_out=FF,
               out_planes=hp.ptr(), ld_pl=hp.ld, pl_plane_stride=hp.plane_stride)
        s = s_out.t
        L.gemm(M, H, FF, L.op_of(hp), L.op_of(w2), passes=r.passes, bias=b2, residual=x2, out32=s, ld_out=H,
               drop_p=spec.drop_p, drop_site=spec.site, rng=rng, out32_zeroed=s_out.ready())
        z = _f32(M, H, device=dev)
        zp = Planes.empty(M, H, dev)
        stats = _f32(M, 2, device=dev)
        L.layernorm_fwd(s, gamma, beta, LN_EPS, z, zp, stats, M, H)
        ctx.r, ctx.spec, ctx.dims = r, spec, (M, H, FF)
        ctx.keep = (xp, w1, w2, hp)
        ctx.save_for_backward(pre, s, stats, gamma)
        ctx.xshape = x.shape
        spec.out_planes = zp
        return z.view(x.shape)

    @staticmethod
    def backward(ctx, dz):
        r, spec = ctx.r, ctx.spec
        M, H, FF = ctx.dims
        xp, w1, w2, hp = ctx.keep
        pre, s, stats, gamma = ctx.saved_tensors
        dev = dz.device
        dx_out = _EarlyOut(r, M, H, FF, dev) if ctx.needs_input_grad[0] else None
        ds = _f32(M, H, device=dev)
        dsp = Planes.empty(M, H, dev)
        acc = r.zeros(3, H)                                               # dgamma, dbeta, db2
        dz2 = _c2d(dz)
        cols = _ln_bwd_chain(r, dz2, s, gamma, stats, ds, dsp, acc, M, H, spec.drop_p, spec.site, ctx.rng)
        dW2 = _dw_out(r, w2, H, FF, dev)
        dW1 = _dw_out(r, w1, FF, H, dev)
        db1 = _f32(FF, device=dev)
        dprep = Planes.empty(M, FF, dev)
        cur = torch.cuda.current_stream(dev)
        side = (r.helper() if r.defer_wgrad else r.fork()) if r.concurrent else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                     # LN column sums, dW2 = ds^T . h
            cols()
            L.gemm(H, FF, M, L.op_of(dsp, True), L.op_of(hp, True), passes=r.passes, out32=dW2, ld_out=FF)
        # d(pre) = (ds . W2) * gelu'(pre), written as planes only
        L.gemm(M, FF, H, L.op_of(dsp), L.op_of(w2, True), passes=r.passes, act=L.ACT_MUL_GELU_GRAD, aux_in=pre, ld_out=FF,
               out_planes=dprep.ptr(), ld_pl=dprep.ld, pl_plane_stride=dprep.plane_stride)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                                     # dW1 = d(pre)^T . x ; db1 = colsum(d(pre))
            L.gemm(FF, H, M, L.op_of(dprep, True), L.op_of(xp, True), passes=r.passes, out32=dW1, ld_out=H)
            L.colsum_planes(dprep, db1)
        dx = None
        if ctx.needs_input_grad[0]:                                       # dx = d(pre) . W1 + ds (residual branch)
            dx = dx_out.t
            L.gemm(M, H, FF, L.op_of(dprep), L.op_of(w1, True), passes=r.passes, out32=dx, ld_out=H, residual=ds,
                   out32_zeroed=dx_out.ready())
        if side is not cur and r.defer_wgrad:
            _keep_for(side, dsp, hp, dW2, dprep, xp, dW1, db1, dz2, s, stats, acc)
            r.defer_join()
        else:
            cur.wait_stream(side)
        return ((dx.view(ctx.xshape) if dx is not None else None), dW1, db1, dW2, acc[2], acc[0], acc[1], None)


def ffn(x, W1, b1, W2, b2, gamma, beta, drop_p: float, site: int):
    """LN(dropout(W2 gelu(W1 x + b1) + b2) + x) (GELU(erf) intermediate activation)."""
    spec = types.SimpleNamespace(drop_p=float(drop_p), site=site, out_planes=None)
    z = FFNFn.apply(x, W1, b1, W2, b2, gamma, beta, spec)
    attach_planes(z, spec.out_planes)
    return z


class BiAttentionFn(Function):
    """vision stream = "1", text stream = "2" as in the reference: ctx1 = softmax(q2 k1^T) v1 (text attends vision),
    ctx2 = softmax(q1 k2^T) v2 (vision attends text)."""

    @staticmethod
    def forward(ctx, xv, vmask, xt, tmask, Wq1, bq1, Wk1, bk1, Wv1, bv1, Wq2, bq2, Wk2, bk2, Wv2, bv2, spec):
        r = rt(xv.device)
        ctx.rng = rng = r.rng
        pairs, V, Kv = xv.shape
        _, T, Kt = xt.shape
        H = Wq1.shape[0]
        heads = spec.heads
        dh = H // heads
        dev = xv.device
        xv2, xt2 = _c2d(xv), _c2d(xt)
        Mv, Mt = pairs * V, pairs * T
        xvp, xtp = planes_of(xv, xv2), planes_of(xt, xt2)
        w1 = r.arena.get((Wq1, Wk1, Wv1))
        w2 = r.arena.get((Wq2, Wk2, Wv2))
        qkv1 = Planes.empty(Mv, 3 * H, dev)
        qkv2 = Planes.empty(Mt, 3 * H, dev)
        b1c, b2c = _cat_bias(r, (bq1, bk1, bv1)), _cat_bias(r, (bq2, bk2, bv2))
        vm, tm = _mask2d(vmask, pairs, V), _mask2d(tmask, pairs, T)
        c1 = _f32(Mt, H, device=dev)
        c1p = Planes.empty(Mt, H, dev)
        c2 = _f32(Mv, H, device=dev)
        c2p = Planes.empty(Mv, H, dev)
        q1, k1, v1 = HeadView(qkv1, 0, V), HeadView(qkv1, H, V), HeadView(qkv1, 2 * H, V)
        q2, k2, v2 = HeadView(qkv2, 0, T), HeadView(qkv2, H, T), HeadView(qkv2, 2 * H, T)
        cur = torch.cuda.current_stream(dev)
        side = r.fork() if r.concurrent else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            L.gemm(Mt, 3 * H, Kt, L.op_of(xtp), L.op_of(w2), passes=r.passes, bias=b2c,
                   out_planes=qkv2.ptr(), ld_pl=qkv2.ld, pl_plane_stride=qkv2.plane_stride)
        L.gemm(Mv, 3 * H, Kv, L.op_of(xvp), L.op_of(w1), passes=r.passes, bias=b1c,
               out_planes=qkv1.ptr(), ld_pl=qkv1.ld, pl_plane_stride=qkv1.plane_stride)
        cur.wait_stream(side)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            att1 = _attn_fwd(r, q2, k1, v1, vm, pairs, heads, dh, spec.drop_p1, spec.site1, c1p, c1, rng)
        att2 = _attn_fwd(r, q1, k2, v2, tm, pairs, heads, dh, spec.drop_p2, spec.site2, c2p, c2, rng)
        cur.wait_stream(side)
        att1.keep_for(cur)
        ctx.r, ctx.spec = r, spec
        ctx.keep = (xvp, xtp, w1, w2, qkv1, qkv2, att1, att2)
        ctx.dims = (pairs, V, T, Kv, Kt, H, heads, dh)
        spec.out_planes = (c1p, c2p)
        spec.probs = (att1.P, att2.P)
        return c1.view(pairs, T, H), c2.view(pairs, V, H)

    @staticmethod
    def backward(ctx, dc1, dc2):
        r, spec = ctx.r, ctx.spec
        pairs, V, T, Kv, Kt, H, heads, dh = ctx.dims
        xvp, xtp, w1, w2, qkv1, qkv2, att1, att2 = ctx.keep
        dev = dc1.device
        Mv, Mt = pairs * V, pairs * T
        dxv_out = _EarlyOut(r, Mv, Kv, 3 * H, dev) if ctx.needs_input_grad[0] else None
        dxt_out = _EarlyOut(r, Mt, Kt, 3 * H, dev) if ctx.needs_input_grad[2] else None
        dO1 = L.split_planes(_c2d(dc1))
        dO2 = L.split_planes(_c2d(dc2))
        d1 = Planes.empty(Mv, 3 * H, dev)
        d2 = Planes.empty(Mt, 3 * H, dev)
        q1, k1, v1 = HeadView(qkv1, 0, V), HeadView(qkv1, H, V), HeadView(qkv1, 2 * H, V)
        q2, k2, v2 = HeadView(qkv2, 0, T), HeadView(qkv2, H, T), HeadView(qkv2, 2 * H, T)
        dq1, dk1, dv1 = HeadView(d1, 0, V), HeadView(d1, H, V), HeadView(d1, 2 * H, V)
        dq2, dk2, dv2 = HeadView(d2, 0, T), HeadView(d2, H, T), HeadView(d2, 2 * H, T)
        cur = torch.cuda.current_stream(dev)
        side = r.fork() if r.concurrent else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            _attn_bwd(r, dO1, q2, k1, v1, att1, pairs, heads, dh, spec.drop_p1, spec.site1, dq2, dk1, dv1, rng=ctx.rng)
        _attn_bwd(r, dO2, q1, k2, v2, att2, pairs, heads, dh, spec.drop_p2, spec.site2, dq1, dk2, dv2, rng=ctx.rng)
        cur.wait_stream(side)
        dxv, dW1, db1 = _linear_bwd(r, d1, xvp, w1, Mv, 3 * H, Kv, dev, ctx.needs_input_grad[0], True, defer=True,
                                     dx_out=dxv_out)
        dxt, dW2, db2 = _linear_bwd(r, d2, xtp, w2, Mt, 3 * H, Kt, dev, ctx.needs_input_grad[2], True, defer=True,
                                     dx_out=dxt_out)
        return ((dxv.view(pairs, V, Kv) if dxv is not None else None), None,
                (dxt.view(pairs, T, Kt) if dxt is not None else None), None,
                dW1[:H], db1[:H], dW1[H:2 * H], db1[H:2 * H], dW1[2 * H:], db1[2 * H:],
                dW2[:H], db2[:H], dW2[H:2 * H], db2[H:2 * H], dW2[2 * H:], db2[2 * H:], None)


def bi_attention(xv, vmask, xt, tmask, params1, params2, heads: int, drop_p1: float, site1: int, drop_p2: float,
                 site2: int):
    spec = types.SimpleNamespace(heads=heads, drop_p1=float(drop_p1), site1=site1, drop_p2=float(drop_p2), site2=site2,
                                 out_planes=None, probs=None)
    c1, c2 = BiAttentionFn.apply(xv, vmask, xt, tmask, *params1, *params2, spec)
    attach_planes(c1, spec.out_planes[0])
    attach_planes(c2, spec.out_planes[1])
    return c1, c2, spec.probs


# ----------------------------------------------------------------------------------------------------
# embeddings
# ----------------------------------------------------------------------------------------------------
class TextEmbedFn(Function):
    @staticmethod
    def forward(ctx, tok, seg, word, pos, typ, gamma, beta, spec):
        r = rt(word.device)
        ctx.rng = rng = r.rng
        pairs, T = tok.shape
        H = word.shape[1]
        M = pairs * T
        dev = word.device
        tok_c, seg_c = tok.contiguous().long(), seg.contiguous().long()
        e = _f32(M, H, device=dev)
        L.embed_text_fwd(tok_c, seg_c, word, pos, typ, e, M, T, H)
        y = _f32(M, H, device=dev)
        yp = Planes.empty(M, H, dev)
        stats = _f32(M, 2, device=dev)
        L.layernorm_fwd(e, gamma, beta, LN_EPS, y, yp, stats, M, H, spec.drop_p, spec.site, rng)
        ctx.r, ctx.spec, ctx.dims = r, spec, (pairs, T, H)
        ctx.shapes = (word.shape, pos.shape, typ.shape)
        ctx.save_for_backward(tok_c, seg_c, e, stats, gamma)
        spec.out_planes = yp
        return y.view(pairs, T, H)

    @staticmethod
    def backward(ctx, dy):
        r, spec = ctx.r, ctx.spec
        pairs, T, H = ctx.dims
        tok, seg, e, stats, gamma = ctx.saved_tensors
        M = pairs * T
        dev = dy.device
        de = _f32(M, H, device=dev)
        acc = r.zeros(2, H)
        dgamma, dbeta = acc[0], acc[1]
        # these are the last kernels of the backward chain: the 94 MB zero fill of the word-embedding gradient runs on a
        # forked stream next to the LayerNorm backward instead of after it
        dword = torch.empty(ctx.shapes[0], dtype=torch.float32, device=dev)
        dpos = torch.empty(ctx.shapes[1], dtype=torch.float32, device=dev)
        dtyp = torch.empty(ctx.shapes[2], dtype=torch.float32, device=dev)
        cur = torch.cuda.current_stream(dev)
        side = r.fork(2) if r.concurrent else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            dword.zero_()
            dpos.zero_()
            dtyp.zero_()
        L.layernorm_bwd(_c2d(dy), e, gamma, stats, de, None, dgamma, dbeta, M, H, post_drop_p=spec.drop_p,
                        post_drop_site=spec.site, rng=ctx.rng)
        cur.wait_stream(side)
        L.embed_text_bwd(tok, seg, de, dword, dpos, dtyp, M, T, H, spec.padding_idx)
        return None, None, dword, dpos, dtyp, dgamma, dbeta, None


def text_embed(tok, seg, word, pos, typ, gamma, beta, drop_p: float, site: int, padding_idx: int = 0):
    spec = types.SimpleNamespace(drop_p=float(drop_p), site=site, padding_idx=padding_idx, out_planes=None)
    y = TextEmbedFn.apply(tok, seg, word, pos, typ, gamma, beta, spec)
    attach_planes(y, spec.out_planes)
    return y


class ImageEmbedFn(Function):
    @staticmethod
    def forward(ctx, feat, loc, Wi, bi, w5, b5, w4, b4, w2, b2, seq, gamma, beta, spec):
        r = rt(Wi.device)
        ctx.rng = rng = r.rng
        pairs, V, F = feat.shape
        H = Wi.shape[0]
        M = pairs * V
        dev = Wi.device
        f2 = _c2d(feat)
        fp = planes_of(feat, f2)
        loc2 = _c2d(loc)
        if loc2.shape[1] != 12:
            raise RuntimeError("yvb200: image_loc must have 12 columns (5 box + 4 orientation + 2 next + frame index)")
        le = _f32(M, H, device=dev)
        L.embed_loc_fwd(loc2, w5, b5, w4, b4, w2, b2, seq, le, M, H)
        wp = r.arena.get((Wi,))
        e = _f32(M, H, device=dev)
        L.gemm(M, H, F, L.op_of(fp), L.op_of(wp), passes=r.passes, bias=bi, residual=le, out32=e, ld_out=H)
        y = _f32(M, H, device=dev)
        yp = Planes.empty(M, H, dev)
        stats = _f32(M, 2, device=dev)
        L.layernorm_fwd(e, gamma, beta, LN_EPS, y, yp, stats, M, H, spec.drop_p, spec.site, rng)
        ctx.r, ctx.spec, ctx.dims, ctx.fp, ctx.wp = r, spec, (pairs, V, F, H), fp, wp
        ctx.save_for_backward(loc2, e, stats, gamma)
        spec.out_planes = yp
        return y.view(pairs, V, H)

    @staticmethod
    def backward(ctx, dy):
        r, spec = ctx.r, ctx.spec
        pairs, V, F, H = ctx.dims
        loc2, e, stats, gamma = ctx.saved_tensors
        M = pairs * V
        dev = dy.device
        de = _f32(M, H, device=dev)
        dep = Planes.empty(M, H, dev)
        acc = r.zeros(3, H)
        dgamma, dbeta = acc[0], acc[1]
        L.layernorm_bwd(_c2d(dy), e, gamma, stats, de, dep, dgamma, dbeta, M, H, post_drop_p=spec.drop_p,
                        post_drop_site=spec.site, rng=ctx.rng, dbias=acc[2])
        dfeat, dWi, dbi = _linear_bwd(r, dep, ctx.fp, ctx.wp, M, H, F, dev, ctx.needs_input_grad[0], True, db=acc[2])
        # the seven small accumulators of the location / frame embeddings come out of the pass's zero slab (no fill kernels
        # at the very end of the backward chain): H x (5 + 1 + 4 + 1 + 2 + 1) floats and the [32, H] frame table
        zb = r.zeros(46, H).view(-1)
        cuts, off = [], 0
        for n_, shape in ((5 * H, (H, 5)), (H, (H,)), (4 * H, (H, 4)), (H, (H,)), (2 * H, (H, 2)), (H, (H,)), (32 * H, (32, H))):
            cuts.append(zb[off:off + n_].view(shape))
            off += n_
        dw5, db5, dw4, db4, dw2, db2, dseq = cuts
        L.embed_loc_bwd(loc2, de, dw5, db5, dw4, db4, dw2, db2, dseq, M, H)
        return ((dfeat.view(pairs, V, F) if dfeat is not None else None), None, dWi, dbi, dw5, db5, dw4, db4, dw2, db2,
                dseq, dgamma, dbeta, None)


def image_embed(feat, loc, Wi, bi, w5, b5, w4, b4, w2, b2, seq, gamma, beta, drop_p: float, site: int):
    spec = types.SimpleNamespace(drop_p=float(drop_p), site=site, out_planes=None)
    y = ImageEmbedFn.apply(feat, loc, Wi, bi, w5, b5, w4, b4, w2, b2, seq, gamma, beta, spec)
    attach_planes(y, spec.out_planes)
    return y
