import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import ops, lib as L
torch.manual_seed(0)
for (M,K,N) in [(96,96,128),(96,64,192),(256,256,256),(96,128,64)]:
    x = torch.randn(M,K,device="cuda"); W = torch.nn.Parameter(torch.randn(N,K,device="cuda")*0.1); b = torch.nn.Parameter(torch.randn(N,device="cuda"))
    ref = x.double()@W.double().t()+b.double()
    y = ops.dense_act(x, W, b, 0)
    print("arena", M,K,N, float((y.double()-ref).norm()/ref.norm()))
    wp = L.split_planes(W.detach()); xp = L.split_planes(x)
    out = torch.empty(M,N,device="cuda")
    L.gemm(M,N,K,L.op_of(xp),L.op_of(wp),passes=3,bias=b,out32=out,ld_out=N)
    print("direct", float((out.double()-ref).norm()/ref.norm()))
    r = ops.rt("cuda"); ap = r.arena.get((W,))
    t = ap.keep.view(2,-1)
    off = (ap.addr - ap.keep.data_ptr())//2
    hi = t[0, off:off+N*K].view(N,K).float(); lo = t[1, off:off+N*K].view(N,K).float()
    print("arena planes err", float((hi+lo-W.detach()).norm()/W.norm()), "plane_stride", ap.plane_stride, "off", off)
    out2 = torch.empty(M,N,device="cuda")
    L.gemm(M,N,K,L.op_of(xp),L.op_of(ap),passes=3,bias=b,out32=out2,ld_out=N)
    print("direct+arena W", float((out2.double()-ref).norm()/ref.norm()))
    L.gemm(M,N,K,L.op_of(xp),L.op_of(ap),passes=1,bias=b,out32=out2,ld_out=N)
    print("direct+arena W passes1", float((out2.double()-ref).norm()/ref.norm()))
