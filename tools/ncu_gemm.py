"""A handful of representative yv_gemm launches for `ncu --set full` (BertBiAttention projection shapes of cfg2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L
def run(M, N, K, passes, variant=0):
    L.set_gemm_variant(variant)
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") * 0.05
    pa, pb = L.split_planes(A), L.split_planes(B)
    bias = torch.randn(N, device="cuda")
    out = L.Planes.empty(M, N, "cuda")
    for _ in range(2):
        L.gemm(M, N, K, L.op_of(pa), L.op_of(pb), passes=passes, bias=bias, out_planes=out.ptr(), ld_pl=out.ld,
               pl_plane_stride=out.plane_stride)
    torch.cuda.synchronize()
    L.set_gemm_variant(0)
run(2304, 3072, 1024, 3)          # query1|key1|value1 projection of BertBiAttention (vision stream, 8 pairs x 288 regions)
run(640, 3072, 768, 3)            # query2|key2|value2 projection (text stream, 8 pairs x 80 tokens)
run(2304, 1024, 1024, 3)          # dense1 / vision FFN shape
run(2304, 3072, 1024, 3, 256)     # the vision projection on CTA pairs (cta_group::2), 256-wide pair tiles
run(2304, 1024, 1024, 3, 128)     # dense1 on CTA pairs, 128-wide pair tiles
run(2304, 3072, 1024, 1)          # vision projection, single-pass bf16
