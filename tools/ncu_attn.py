"""The fused attention launches of one cfg2 layer set for `ncu --set full` (8 pairs): text self-attention, vision
self-attention and the two directions of BertBiAttention, forward and backward, dropout on."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L, ops
r = ops.rt("cuda")
pairs = 8
for name, heads, dh, Tq, Tk in (("bi t<-v", 8, 128, 80, 288), ("bi v<-t", 8, 128, 288, 80), ("vision", 8, 128, 288, 288),
                                ("text", 12, 64, 80, 80)):
    H = heads * dh
    qp = L.split_planes(torch.randn(pairs * Tq, 3 * H, device="cuda"))
    kp = L.split_planes(torch.randn(pairs * Tk, 3 * H, device="cuda"))
    dOp = L.split_planes(torch.randn(pairs * Tq, H, device="cuda"))
    mask = torch.zeros(pairs, Tk, device="cuda")
    out = L.Planes.empty(pairs * Tq, H, "cuda")
    lse = torch.empty(pairs * heads * Tq, device="cuda")
    dq = L.Planes.empty(pairs * Tq, H, "cuda")
    dkv = L.Planes.empty(pairs * Tk, 2 * H, "cuda")
    ws = torch.empty(L.attn_bwd_workspace_bytes(pairs, heads, dh, Tq, Tk), dtype=torch.uint8, device="cuda")
    tk = torch.zeros(pairs * heads, dtype=torch.int32, device="cuda")
    scale = 1.0 / math.sqrt(dh)
    for _ in range(2):
        L.attn_fwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), mask, pairs, heads, dh, scale,
                   out, None, lse, drop_p=0.1, drop_site=5, rng=r.rng)
        L.attn_bwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), L.head_view(dOp, 0, Tq),
                   L.head_view(out, 0, Tq), mask, lse, pairs, heads, dh, scale, L.head_view(dq, 0, Tq), L.head_view(dkv, 0, Tk),
                   L.head_view(dkv, H, Tk), ws, tk, drop_p=0.1, drop_site=5, rng=r.rng)
    torch.cuda.synchronize()
