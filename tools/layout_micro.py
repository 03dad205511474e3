import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L
def bench(M,N,K,a_mn,b_mn,passes=3,iters=30):
    A = torch.randn(K,M,device="cuda") if a_mn else torch.randn(M,K,device="cuda")
    B = torch.randn(K,N,device="cuda") if b_mn else torch.randn(N,K,device="cuda")
    pa, pb = L.split_planes(A), L.split_planes(B)
    out = torch.empty(M,N,device="cuda")
    for _ in range(3): L.gemm(M,N,K,L.op_of(pa,a_mn),L.op_of(pb,b_mn),passes=passes,out32=out,ld_out=N)
    torch.cuda.synchronize(); torch.cuda._sleep(int(1e8))
    e0,e1 = torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.gemm(M,N,K,L.op_of(pa,a_mn),L.op_of(pb,b_mn),passes=passes,out32=out,ld_out=N)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)/iters*1e3
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} p={passes}: {t:.1f} us  {2.0*M*N*K/t/1e6:.1f} algTF/s")
for (M,N,K) in [(768,3072,640),(1024,1024,2304),(2304,1024,1024)]:
    for a_mn in (False,True):
        for b_mn in (False,True):
            bench(M,N,K,a_mn,b_mn)
