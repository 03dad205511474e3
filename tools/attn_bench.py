"""Developer benchmark (GPU box): fused attention kernels vs the un-fused GEMM -> softmax -> GEMM chain on the four
attention shapes of the cfg2 step (8 pairs), forward and backward, with and without dropout.  CUDA events, 30 launches
back to back after 5 warm-ups, 256 MB L2 flush omitted (operands are L2-resident in the real step as well).
    python tools/attn_bench.py [pairs]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
from yvb200 import lib as L, ops  # noqa: E402

SHAPES = [("text self 80x80 dh64", 12, 64, 80, 80), ("vision self 288x288 dh128", 8, 128, 288, 288),
          ("bi t<-v 80x288 dh128", 8, 128, 80, 288), ("bi v<-t 288x80 dh128", 8, 128, 288, 80)]


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    r = ops.rt("cuda")
    for name, heads, dh, Tq, Tk in SHAPES:
        H = heads * dh
        qp = L.split_planes(torch.randn(pairs * Tq, 3 * H, device="cuda"))
        kp = L.split_planes(torch.randn(pairs * Tk, 3 * H, device="cuda"))
        dOp = L.split_planes(torch.randn(pairs * Tq, H, device="cuda"))
        mask = torch.zeros(pairs, Tk, device="cuda")
        q, k, v = ops.HeadView(qp, 0, Tq), ops.HeadView(kp, H, Tk), ops.HeadView(kp, 2 * H, Tk)
        out = L.Planes.empty(pairs * Tq, H, "cuda")
        lse = torch.empty(pairs * heads * Tq, device="cuda")
        dq = L.Planes.empty(pairs * Tq, H, "cuda")
        dkv = L.Planes.empty(pairs * Tk, 2 * H, "cuda")
        ws = torch.empty(L.attn_bwd_workspace_bytes(pairs, heads, dh, Tq, Tk), dtype=torch.uint8, device="cuda")
        tk = torch.zeros(pairs * heads, dtype=torch.int32, device="cuda")
        scale = 1.0 / math.sqrt(dh)
        flop = 4.0 * pairs * heads * Tq * Tk * dh
        for p_drop in (0.0, 0.1):
            def f_fwd():
                L.attn_fwd(q.view(), k.view(), v.view(), mask, pairs, heads, dh, scale, out, None, lse, drop_p=p_drop,
                           drop_site=5, rng=r.rng)

            def f_bwd():
                L.attn_bwd(q.view(), k.view(), v.view(), L.head_view(dOp, 0, Tq), L.head_view(out, 0, Tq), mask, lse, pairs,
                           heads, dh, scale, L.head_view(dq, 0, Tq), L.head_view(dkv, 0, Tk), L.head_view(dkv, H, Tk), ws, tk,
                           drop_p=p_drop, drop_site=5, rng=r.rng)
            saved = {}

            def u_fwd():
                saved["P"], saved["Pp"] = ops._attn_fwd_unfused(r, q, k, v, mask, pairs, heads, dh, p_drop, 5, out, None, r.rng)

            def u_bwd():
                ops._attn_bwd_unfused(r, dOp, q, k, v, saved["P"], saved["Pp"], pairs, heads, dh, p_drop, 5,
                                      ops.HeadView(dq, 0, Tq), ops.HeadView(dkv, 0, Tk), ops.HeadView(dkv, H, Tk), rng=r.rng)
            tf, tb = timeit(f_fwd), timeit(f_bwd)
            tuf = timeit(u_fwd)
            tub = timeit(u_bwd)
            print(f"{name:28s} p={p_drop}: fused fwd {tf:6.1f} us ({flop / tf / 1e6:6.1f} TFLOP/s)  bwd {tb:6.1f} us "
                  f"({2.5 * flop / tb / 1e6:6.1f})   un-fused fwd {tuf:6.1f} us  bwd {tub:6.1f} us", flush=True)


if __name__ == "__main__":
    main()
