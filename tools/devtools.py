"""Developer instrumentation for bench.py and tools/ (NOT part of the product package): wraps the loaded C ABI so that

  * ``contraction_only()``  -- every launcher except the tensor-core contraction kernels (yv_gemm, yv_attn_fwd,
    yv_attn_bwd) returns without launching, and every contraction launch is recorded with its algorithmic FLOPs.  A step
    captured under it is a CUDA graph holding exactly the step's contraction launches (operands are then uninitialised
    memory: timing only, never results);
  * ``trace_gemms()``       -- brackets every yv_gemm launch with CUDA events on the launching stream.

Both work by temporarily replacing ``yvb200.lib._lib`` (the object ``lib.load()`` hands out) with a proxy."""
import contextlib
import ctypes as C

import torch

CONTRACTIONS = ("yv_gemm", "yv_attn_fwd", "yv_attn_bwd")
QUERIES = ("yv_last_error", "yv_version", "yv_launch_count", "yv_gemm_splits", "yv_gemm_set_variant", "yv_rng_advance",
           "yv_attn_supported", "yv_attn_bwd_workspace_bytes")


def _flops(name, arg):
    s = arg._obj
    if name == "yv_gemm":
        return 2.0 * s.M * s.N * s.K * s.a.nb0 * s.a.nb1, (s.M, s.N, s.K, int(s.a.nb0 * s.a.nb1), s.passes)
    pairs, heads, dh, Tq, Tk = s.pairs, s.heads, s.dh, s.q.rows, s.k.rows
    per = 2.0 * pairs * heads * Tq * Tk * dh                       # one [Tq x Tk x dh] product
    return (2 if name == "yv_attn_fwd" else 4) * per, (name, pairs, heads, Tq, Tk, dh, s.passes)


class _Proxy:
    def __init__(self, real, skip_others, trace, events):
        self._real = real
        for name in dir(real):
            pass
        self._skip, self._trace, self._events = skip_others, trace, events

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if name in CONTRACTIONS:
            def call(arg, stream, _fn=fn, _name=name):
                fl, shape = _flops(_name, arg)
                if self._events:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = _fn(arg, stream)
                    e1.record()
                    self._trace.append((_name, shape, fl, e0, e1))
                    return rc
                self._trace.append((_name, shape, fl, None, None))
                return _fn(arg, stream)
            return call
        if name in QUERIES or not self._skip:
            return fn
        return lambda *a: 0


@contextlib.contextmanager
def _patched(skip_others, events):
    from yvb200 import lib
    real = lib.load()
    trace = []
    lib._lib = _Proxy(real, skip_others, trace, events)
    try:
        yield trace
    finally:
        lib._lib = real


def contraction_only():
    return _patched(True, False)


def trace_gemms():
    return _patched(False, True)
