"""Developer tool: latency of dependent kernel chains inside a CUDA graph (what the step's critical path sees)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L

def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def chain_gemm(M, N, K, n=20, passes=3, full_epi=True):
    W = L.split_planes(torch.randn(N, K, device="cuda") * 0.05)
    bias = torch.randn(N, device="cuda")
    xs = [L.Planes.empty(M, K, "cuda") for _ in range(2)]
    L.split_planes(torch.randn(M, K, device="cuda"), xs[0])
    out = torch.empty(M, N, device="cuda")
    assert N == K
    def fn():
        for i in range(n):
            a, b = xs[i % 2], xs[(i + 1) % 2]
            if full_epi:
                L.gemm(M, N, K, L.op_of(a), L.op_of(W), passes=passes, bias=bias, out32=out, ld_out=N,
                       out_planes=b.ptr(), ld_pl=b.ld, pl_plane_stride=b.plane_stride)
            else:
                L.gemm(M, N, K, L.op_of(a), L.op_of(W), passes=passes, out_planes=b.ptr(), ld_pl=b.ld, pl_plane_stride=b.plane_stride)
    t = graph_time(fn)
    print(f"chain gemm M={M} N={N} K={K} p={passes} full_epi={full_epi}: {t/n:.1f} us per GEMM  ({2.0*M*N*K/(t/n)/1e6:.0f} algTF/s)")

def chain_ln(M, C, n=20):
    x = torch.randn(M, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    y = torch.empty(M, C, device="cuda"); yp = L.Planes.empty(M, C, "cuda"); st = torch.empty(M, 2, device="cuda")
    dx = torch.empty(M, C, device="cuda"); dxp = L.Planes.empty(M, C, "cuda"); dg = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    def f1():
        for i in range(n): L.layernorm_fwd(x, g, b, 1e-12, y, yp, st, M, C)
    def f2():
        for i in range(n): L.layernorm_bwd(y, x, g, st, dx, dxp, dg, db, M, C)
    def f3():
        for i in range(n): L.colsum_planes(yp, db)
    def f4():
        for i in range(n): L.split_planes(x, yp)
    print(f"M={M} C={C}: ln_fwd {graph_time(f1)/n:.1f} us, ln_bwd {graph_time(f2)/n:.1f} us, colsum_planes {graph_time(f3)/n:.1f} us, split {graph_time(f4)/n:.1f} us")

if "ln" not in sys.argv:
    chain_gemm(2304, 1024, 1024)
    chain_gemm(2304, 1024, 1024, full_epi=False)
    chain_gemm(2304, 1024, 1024, passes=1)
    chain_gemm(640, 768, 768)
    chain_gemm(640, 768, 768, passes=1)
    chain_gemm(128, 128, 128)
chain_ln(2304, 1024)
chain_ln(640, 768)
