import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L
def bench(M,N,K,passes=3,iters=50, interleave=False):
    A = torch.randn(M,K,device="cuda"); B = torch.randn(N,K,device="cuda")
    pa, pb = L.split_planes(A), L.split_planes(B)
    out = torch.empty(M,N,device="cuda")
    x = torch.randn(1024,1024,device="cuda")
    for _ in range(3): L.gemm(M,N,K,L.op_of(pa),L.op_of(pb),passes=passes,out32=out,ld_out=N)
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e8))
    e0,e1 = torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.gemm(M,N,K,L.op_of(pa),L.op_of(pb),passes=passes,out32=out,ld_out=N)
        if interleave: x.add_(1.0)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)/iters*1e3
    print(f"M={M} N={N} K={K} p={passes} interleave={interleave}: {t:.1f} us/iter  {2.0*M*N*K/t/1e6:.1f} algTF/s")
for il in (False, True):
    bench(8,1024,8,3,interleave=il)
    bench(128,128,64,3,interleave=il)
    bench(128,128,4096,3,interleave=il)
    bench(2304,1024,1024,3,interleave=il)
    bench(2304,1024,1024,1,interleave=il)
    bench(4096,4096,4096,1,iters=10,interleave=il)
    bench(4096,4096,4096,3,iters=10,interleave=il)
