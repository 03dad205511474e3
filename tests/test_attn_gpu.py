"""GPU parity of the fused attention kernels (yv_attn_fwd / yv_attn_bwd, called through the C ABI) against
  (a) a plain torch fp64 restatement of vilbert/vilbert.py:294-306 / :423-435 / :577-616 fed with the very same operand
      values (hi + lo planes), and
  (b) the un-fused kernel chain (yv_gemm -> yv_softmax_fwd -> yv_gemm and its backward), which shares the dropout
      counter RNG, so train-mode results must agree as well.
Tolerance: 2e-5 relative (bf16x3 operands, fp32 accumulation); dropout-on comparisons 1e-4."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _planes(x2d):
    from yvb200 import lib as L
    return L.split_planes(x2d.contiguous())


def _setup(pairs, heads, dh, Tq, Tk, seed, grow=False):
    """Random Q / K / V projections laid out like the drop-in's fused Q|K|V buffers ([pairs*S, 3H] plane pairs)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    H = heads * dh
    qbuf = torch.randn(pairs * Tq, 3 * H, device="cuda", generator=g)
    kbuf = torch.randn(pairs * Tk, 3 * H, device="cuda", generator=g)
    if grow:
        # logits that keep growing along the key axis: every chunk raises the running row maximum by far more than the
        # lazy-rescale threshold, so the online-softmax rescaling of the TMEM accumulator is exercised
        ramp = torch.linspace(0.2, 3.0, Tk, device="cuda").repeat(pairs).unsqueeze(1)
        kbuf = kbuf.clone()
        kbuf[:, H:2 * H] = kbuf[:, H:2 * H].abs() * ramp
        qbuf = qbuf.clone()
        qbuf[:, :H] = qbuf[:, :H].abs()
    mask = torch.zeros(pairs, Tk, device="cuda")
    if Tk > 4:
        mask[1 % pairs, Tk - max(1, Tk // 8):] = -10000.0
    return qbuf, kbuf, mask


def _ref(qp, kp, mask, pairs, heads, dh, Tq, Tk, dO=None):
    """fp64 attention on the values the kernels see (hi + lo)."""
    H = heads * dh
    qf = qp.float().double().view(pairs, Tq, 3 * H)
    kf = kp.float().double().view(pairs, Tk, 3 * H)
    q = qf[..., :H].reshape(pairs, Tq, heads, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    k = kf[..., H:2 * H].reshape(pairs, Tk, heads, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    v = kf[..., 2 * H:].reshape(pairs, Tk, heads, dh).permute(0, 2, 1, 3).clone().requires_grad_(True)
    s = q @ k.transpose(-1, -2) / math.sqrt(dh) + mask.double()[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(pairs * Tq, H)
    lse = torch.logsumexp(s, dim=-1)
    if dO is None:
        return o, lse
    o.backward(dO.double())
    merge = lambda t, T: t.grad.permute(0, 2, 1, 3).reshape(pairs * T, H)   # noqa: E731
    return o.detach(), lse.detach(), merge(q, Tq), merge(k, Tk), merge(v, Tk)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


SHAPES = [
    # pairs, heads, dh, Tq, Tk
    (2, 12, 64, 80, 80),        # text self-attention (cfg2)
    (2, 8, 128, 288, 288),      # vision self-attention (cfg2)
    (2, 8, 128, 80, 288),       # bi-attention: text queries over vision keys
    (2, 8, 128, 288, 80),       # bi-attention: vision queries over text keys
    (2, 12, 64, 20, 20),        # cfg1 text (ragged: 20 tokens)
    (2, 8, 128, 36, 20),        # cfg1 bi-attention (1 frame x 36 regions)
    (1, 2, 128, 1152, 1152),    # 32-frame trajectory (cfg4)
    (1, 2, 128, 80, 1152),
    (3, 4, 64, 130, 70),        # odd sizes: query tile edge, key chunk edge
]


@pytest.mark.parametrize("pairs,heads,dh,Tq,Tk", SHAPES)
def test_fused_attention_forward_backward_match_fp64(pairs, heads, dh, Tq, Tk):
    _need_gpu()
    from yvb200 import lib as L
    H = heads * dh
    qbuf, kbuf, mask = _setup(pairs, heads, dh, Tq, Tk, seed=Tq * 7 + Tk)
    qp, kp = _planes(qbuf), _planes(kbuf)
    out = L.Planes.empty(pairs * Tq, H, "cuda")
    out32 = torch.empty(pairs * Tq, H, device="cuda")
    lse = torch.empty(pairs * heads * Tq, device="cuda")
    scale = 1.0 / math.sqrt(dh)
    n0 = L.launch_count()
    L.attn_fwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), mask, pairs, heads, dh, scale,
               out, out32, lse)
    assert L.launch_count() == n0 + 1
    dO = torch.randn(pairs * Tq, H, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    dOp = _planes(dO)
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = _ref(qp, kp, mask, pairs, heads, dh, Tq, Tk, dOp.float())
    assert _rel(out32, o_ref) < 2e-5
    assert _rel(out.float(), o_ref) < 2e-5
    assert _rel(lse.view(pairs, heads, Tq), lse_ref) < 1e-5
    dq = L.Planes.empty(pairs * Tq, H, "cuda")
    dkv = L.Planes.empty(pairs * Tk, 2 * H, "cuda")
    ws = torch.empty(L.attn_bwd_workspace_bytes(pairs, heads, dh, Tq, Tk), dtype=torch.uint8, device="cuda")
    tk = torch.zeros(pairs * heads, dtype=torch.int32, device="cuda")
    L.attn_bwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), L.head_view(dOp, 0, Tq),
               L.head_view(out, 0, Tq), mask, lse, pairs, heads, dh, scale, L.head_view(dq, 0, Tq),
               L.head_view(dkv, 0, Tk), L.head_view(dkv, H, Tk), ws, tk)
    torch.cuda.synchronize()
    assert _rel(dq.float(), dq_ref) < 5e-5
    assert _rel(dkv.float()[:, :H], dk_ref) < 5e-5
    assert _rel(dkv.float()[:, H:], dv_ref) < 5e-5


def test_fused_attention_rescales_accumulator_when_row_maximum_grows():
    _need_gpu()
    from yvb200 import lib as L
    pairs, heads, dh, Tq, Tk = 2, 8, 128, 288, 288
    H = heads * dh
    qbuf, kbuf, mask = _setup(pairs, heads, dh, Tq, Tk, seed=3, grow=True)
    qp, kp = _planes(qbuf), _planes(kbuf)
    out = L.Planes.empty(pairs * Tq, H, "cuda")
    lse = torch.empty(pairs * heads * Tq, device="cuda")
    L.attn_fwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), mask, pairs, heads, dh,
               1.0 / math.sqrt(dh), out, None, lse)
    o_ref, lse_ref = _ref(qp, kp, mask, pairs, heads, dh, Tq, Tk)
    # the reference's own row maxima must really differ between the first and the last key chunk by > 8
    assert _rel(out.float(), o_ref) < 2e-5
    assert _rel(lse.view(pairs, heads, Tq), lse_ref) < 1e-5


@pytest.mark.parametrize("pairs,heads,dh,Tq,Tk", [(2, 12, 64, 80, 80), (2, 8, 128, 288, 288), (2, 8, 128, 80, 288)])
def test_fused_attention_with_dropout_matches_unfused_kernels(pairs, heads, dh, Tq, Tk):
    """Same counter RNG, same element index: the fused kernels drop exactly the probabilities the un-fused chain drops."""
    _need_gpu()
    from yvb200 import lib as L, ops
    r = ops.rt("cuda")
    H = heads * dh
    qbuf, kbuf, mask = _setup(pairs, heads, dh, Tq, Tk, seed=11)
    qp, kp = _planes(qbuf), _planes(kbuf)
    q, k, v = ops.HeadView(qp, 0, Tq), ops.HeadView(kp, H, Tk), ops.HeadView(kp, 2 * H, Tk)
    site, p_drop = 77, 0.1
    dO = torch.randn(pairs * Tq, H, device="cuda", generator=torch.Generator(device="cuda").manual_seed(6))
    dOp = _planes(dO)
    # un-fused chain
    c_un = L.Planes.empty(pairs * Tq, H, "cuda")
    P, Pp = ops._attn_fwd_unfused(r, q, k, v, mask, pairs, heads, dh, p_drop, site, c_un, None, r.rng)
    d_un = L.Planes.empty(pairs * Tq, H, "cuda")
    dkv_un = L.Planes.empty(pairs * Tk, 2 * H, "cuda")
    ops._attn_bwd_unfused(r, dOp, q, k, v, P, Pp, pairs, heads, dh, p_drop, site, ops.HeadView(d_un, 0, Tq),
                          ops.HeadView(dkv_un, 0, Tk), ops.HeadView(dkv_un, H, Tk), rng=r.rng)
    # fused
    c_f = L.Planes.empty(pairs * Tq, H, "cuda")
    lse = torch.empty(pairs * heads * Tq, device="cuda")
    scale = 1.0 / math.sqrt(dh)
    L.attn_fwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), mask, pairs, heads, dh, scale,
               c_f, None, lse, drop_p=p_drop, drop_site=site, rng=r.rng)
    d_f = L.Planes.empty(pairs * Tq, H, "cuda")
    dkv_f = L.Planes.empty(pairs * Tk, 2 * H, "cuda")
    ws = torch.empty(L.attn_bwd_workspace_bytes(pairs, heads, dh, Tq, Tk), dtype=torch.uint8, device="cuda")
    tk = torch.zeros(pairs * heads, dtype=torch.int32, device="cuda")
    L.attn_bwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), L.head_view(dOp, 0, Tq),
               L.head_view(c_f, 0, Tq), mask, lse, pairs, heads, dh, scale, L.head_view(d_f, 0, Tq),
               L.head_view(dkv_f, 0, Tk), L.head_view(dkv_f, H, Tk), ws, tk, drop_p=p_drop, drop_site=site, rng=r.rng)
    torch.cuda.synchronize()
    assert _rel(c_f.float(), c_un.float()) < 1e-4
    assert _rel(d_f.float(), d_un.float()) < 1e-4
    assert _rel(dkv_f.float(), dkv_un.float()) < 1e-4
    # and dropout is really on: the eval-mode context differs
    c_e = L.Planes.empty(pairs * Tq, H, "cuda")
    L.attn_fwd(L.head_view(qp, 0, Tq), L.head_view(kp, H, Tk), L.head_view(kp, 2 * H, Tk), mask, pairs, heads, dh, scale,
               c_e, None, None)
    assert _rel(c_f.float(), c_e.float()) > 1e-2
