"""GPU parity tests of the C-ABI primitives against plain fp64/fp32 host-free torch math on the same
inputs.  Tolerances (written per test): bf16x3 GEMM 3e-5 relative to the fp64 product of the fp32 operands;
plain-bf16 GEMM 1e-5 against the fp64 product of the bf16-rounded operands; row ops 1e-5."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yvb200 import lib
    lib.load()
    return lib


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _mk(rows, cols, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g, device="cuda") * scale).contiguous()


def _run_gemm(L, A, B, a_mn, b_mn, passes, **kw):
    """A is [M,K] math-wise, B is [N,K] math-wise; the stored layout is transposed when *_mn is set."""
    M, K = A.shape
    N = B.shape[0]
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    # pad leading dims to a multiple of 8 elements
    def planes_of(x):
        r, c = x.shape
        ld = (c + 7) // 8 * 8
        buf = torch.zeros(r, ld, device="cuda")
        buf[:, :c] = x
        p = L.Planes.empty(r, c, "cuda", ld=ld)
        L.split_planes(buf[:, :c], p)
        return p
    pa, pb = planes_of(As), planes_of(Bs)
    out = torch.full((M, N), float("nan"), device="cuda")
    L.gemm(M, N, K, L.op_of(pa, a_mn), L.op_of(pb, b_mn), passes=passes, out32=out, ld_out=N, **kw)
    torch.cuda.synchronize()
    return out, pa, pb


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(256, 256, 192), (80, 1601, 104), (640, 768, 768), (8, 1024, 768), (300, 72, 40)])
def test_gemm_bf16x3_layouts(L, a_mn, b_mn, shape):
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("transposed operand needs ld % 8 == 0 for the test helper")
    A, B = _mk(M, K, 1), _mk(N, K, 2, 0.05)
    out, _, _ = _run_gemm(L, A, B, a_mn, b_mn, 3)
    ref = A.double() @ B.double().t()
    assert _rel(out, ref) < 3e-5


@pytest.fixture
def variant(request, L):
    """Force one yv_gemm kernel variant for the test, restore automatic selection afterwards."""
    L.set_gemm_variant(request.param)
    yield request.param
    L.set_gemm_variant(0)


VARIANTS = [32, 64, 128, 256]


@pytest.mark.parametrize("variant", VARIANTS, indirect=True)
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("shape", [(512, 384, 160), (2304, 1024, 96), (640, 1601, 104), (1024, 768, 2304), (264, 72, 40),
                                   (768, 30522, 64)])
def test_gemm_every_variant_matches_fp64(L, variant, a_mn, b_mn, shape):
    """Each kernel variant (single CTA k32 / k64, CTA pairs of width 128 / 256) on whole, ragged, split-K and
    wide problems in all four operand layouts: 3e-5 of the fp64 product."""
    M, N, K = shape
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("transposed operand needs ld % 8 == 0 for the test helper")
    A, B = _mk(M, K, 21), _mk(N, K, 22, 0.05)
    out, _, _ = _run_gemm(L, A, B, a_mn, b_mn, 3)
    ref = A.double() @ B.double().t()
    assert not torch.isnan(out).any()
    assert _rel(out, ref) < 3e-5


@pytest.mark.parametrize("variant", [128, 256], indirect=True)
def test_gemm_pair_variants_full_epilogue_bitwise_equal_single_cta(L, variant):
    """The CTA-pair kernels run the same epilogue on the same fp32 accumulators: bias, GELU, saved pre-activation,
    dropout, residual and plane outputs must agree with the single-CTA kernel to the last bit (same k order)."""
    M, N, K = 640, 1048, 328
    A, B = _mk(M, K, 31), _mk(N, K, 32, 0.1)
    bias = _mk(1, N, 33)[0].contiguous()
    res = _mk(M, N, 34)
    rng = torch.tensor([99, 3], dtype=torch.int64, device="cuda")

    def run():
        aux = torch.zeros(M, N, device="cuda")
        pl = L.Planes.empty(M, N, "cuda")
        pl.keep.zero_()
        out, _, _ = _run_gemm(L, A, B, False, False, 3, bias=bias, act=L.ACT_GELU, aux_out=aux, residual=res,
                              out_planes=pl.ptr(), ld_pl=pl.ld, pl_plane_stride=pl.plane_stride, alpha=0.5,
                              drop_p=0.1, drop_site=5, rng=rng)
        return out, aux, pl.keep.clone()

    got = run()
    L.set_gemm_variant(32)
    want = run()
    for g, w in zip(got, want):
        assert torch.equal(g, w)


@pytest.mark.parametrize("variant", [128, 256], indirect=True)
def test_gemm_pair_variants_plain_bf16(L, variant):
    A, B = _mk(520, 512, 3), _mk(264, 512, 4)
    out, _, _ = _run_gemm(L, A, B, False, False, 1)
    ref = A.bfloat16().double() @ B.bfloat16().double().t()
    assert _rel(out, ref) < 1e-5


def test_gemm_plain_bf16_matches_rounded_operands(L):
    A, B = _mk(384, 512, 3), _mk(256, 512, 4)
    out, _, _ = _run_gemm(L, A, B, False, False, 1)
    ref = A.bfloat16().double() @ B.bfloat16().double().t()
    assert _rel(out, ref) < 1e-5


def test_gemm_epilogue_bias_gelu_aux_residual_planes(L):
    M, N, K = 200, 328, 136
    A, B = _mk(M, K, 5), _mk(N, K, 6, 0.1)
    bias = _mk(1, N, 7)[0].contiguous()
    res = _mk(M, N, 8)
    aux = torch.zeros(M, N, device="cuda")
    pl = L.Planes.empty(M, N, "cuda")
    out, _, _ = _run_gemm(L, A, B, False, False, 3, bias=bias, act=L.ACT_GELU, aux_out=aux, residual=res,
                          out_planes=pl.ptr(), ld_pl=pl.ld, pl_plane_stride=pl.plane_stride, alpha=0.5)
    pre = 0.5 * (A.double() @ B.double().t()) + bias.double()
    ref = pre * 0.5 * (1 + torch.erf(pre / math.sqrt(2))) + res.double()
    assert _rel(aux, pre) < 3e-5
    assert _rel(out, ref) < 3e-5
    assert _rel(pl.float(), ref) < 3e-5           # hi + lo reproduces fp32 to ~2^-16


def test_gemm_epilogue_backward_multipliers(L):
    M, N, K = 128, 256, 64
    A, B = _mk(M, K, 9), _mk(N, K, 10, 0.1)
    pre = _mk(M, N, 11)
    out, _, _ = _run_gemm(L, A, B, False, False, 3, act=L.ACT_MUL_GELU_GRAD, aux_in=pre)
    x = pre.double()
    gp = 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    assert _rel(out, (A.double() @ B.double().t()) * gp) < 3e-5
    out, _, _ = _run_gemm(L, A, B, False, False, 3, act=L.ACT_MUL_RELU_MASK, aux_in=pre)
    assert _rel(out, (A.double() @ B.double().t()) * (pre > 0)) < 3e-5


def test_gemm_dropout_is_deterministic_and_unbiased(L):
    M, N, K = 512, 512, 64
    A, B = _mk(M, K, 12), _mk(N, K, 13)
    rng = torch.tensor([1234, 0], dtype=torch.int64, device="cuda")
    o1, _, _ = _run_gemm(L, A, B, False, False, 1, drop_p=0.1, drop_site=7, rng=rng)
    o2, _, _ = _run_gemm(L, A, B, False, False, 1, drop_p=0.1, drop_site=7, rng=rng)
    o3, _, _ = _run_gemm(L, A, B, False, False, 1, drop_p=0.1, drop_site=8, rng=rng)
    base, _, _ = _run_gemm(L, A, B, False, False, 1)
    assert torch.equal(o1, o2)
    keep = (o1 != 0).float().mean().item()
    assert abs(keep - 0.9) < 0.01
    kept = o1 != 0
    assert _rel(o1[kept], base[kept] / 0.9) < 1e-6
    assert ((o1 != 0) != (o3 != 0)).float().mean().item() > 0.1      # another site draws another mask
    L.rng_advance(rng)
    o4, _, _ = _run_gemm(L, A, B, False, False, 1, drop_p=0.1, drop_site=7, rng=rng)
    assert ((o1 != 0) != (o4 != 0)).float().mean().item() > 0.1      # another step draws another mask


def test_gemm_batched_head_views(L):
    """Q.K^T per (pair, head) straight out of a fused [M, 3H] projection buffer, and P.V with MN-major V."""
    pairs, heads, T, dh = 3, 4, 80, 64
    H = heads * dh
    qkv = _mk(pairs * T, 3 * H, 14)
    pq = L.split_planes(qkv)
    S = torch.full((pairs, heads, T, T), float("nan"), device="cuda")
    a = L.operand(pq.ptr(0), dh, T, 3 * H, pq.plane_stride, False, heads, dh, pairs, T * 3 * H)
    b = L.operand(pq.ptr(H), dh, T, 3 * H, pq.plane_stride, False, heads, dh, pairs, T * 3 * H)
    L.gemm(T, T, dh, a, b, passes=3, out32=S, ld_out=T, out_sb0=T * T, out_sb1=heads * T * T)
    q = qkv[:, :H].view(pairs, T, heads, dh).permute(0, 2, 1, 3).double()
    k = qkv[:, H:2 * H].view(pairs, T, heads, dh).permute(0, 2, 1, 3).double()
    v = qkv[:, 2 * H:].view(pairs, T, heads, dh).permute(0, 2, 1, 3).double()
    torch.cuda.synchronize()
    assert _rel(S, q @ k.transpose(-1, -2)) < 3e-5
    P = torch.softmax(S, -1).contiguous()
    pp = L.split_planes(P.view(-1, T))
    ctx = L.Planes.empty(pairs * T, H, "cuda")
    a2 = L.operand(pp.ptr(), T, T, pp.ld, pp.plane_stride, False, heads, T * pp.ld, pairs, heads * T * pp.ld)
    b2 = L.operand(pq.ptr(2 * H), dh, T, 3 * H, pq.plane_stride, True, heads, dh, pairs, T * 3 * H)
    L.gemm(T, dh, T, a2, b2, passes=3, out_planes=ctx.ptr(), ld_pl=H, pl_sb0=dh, pl_sb1=T * H,
           pl_plane_stride=ctx.plane_stride)
    torch.cuda.synchronize()
    ref = (P.double() @ v).permute(0, 2, 1, 3).reshape(pairs * T, H)
    assert _rel(ctx.float(), ref) < 3e-5


@pytest.mark.parametrize("M,Cd", [(300, 768), (2304, 1024), (37, 64), (9, 32), (130, 200), (641, 512)])
def test_layernorm_fwd_bwd(L, M, Cd):
    x = _mk(M, Cd, 20).requires_grad_(True)
    g = (1 + 0.1 * _mk(1, Cd, 21)[0]).contiguous().requires_grad_(True)
    b = (0.1 * _mk(1, Cd, 22)[0]).contiguous().requires_grad_(True)
    y = torch.empty(M, Cd, device="cuda")
    yp = L.Planes.empty(M, Cd, "cuda")
    st = torch.empty(M, 2, device="cuda")
    L.layernorm_fwd(x.detach(), g.detach(), b.detach(), 1e-12, y, yp, st, M, Cd)
    xd = x.double()
    u = xd.mean(-1, keepdim=True)
    s = (xd - u).pow(2).mean(-1, keepdim=True)
    ref = g.double() * ((xd - u) / torch.sqrt(s + 1e-12)) + b.double()
    assert _rel(y, ref) < 1e-6 and _rel(yp.float(), ref) < 2e-5
    dy = _mk(M, Cd, 23)
    ref.backward(dy.double())
    dx = torch.empty(M, Cd, device="cuda")
    dxp = L.Planes.empty(M, Cd, "cuda")
    dg = torch.zeros(Cd, device="cuda")
    db = torch.zeros(Cd, device="cuda")
    dbias = torch.zeros(Cd, device="cuda")
    L.layernorm_bwd(dy, x.detach(), g.detach(), st, dx, dxp, dg, db, M, Cd, dbias=dbias)
    torch.cuda.synchronize()
    assert _rel(dx, x.grad) < 1e-5 and _rel(dxp.float(), x.grad) < 2e-5
    assert _rel(dg, g.grad) < 1e-5 and _rel(db, b.grad) < 1e-5
    assert _rel(dbias, x.grad.sum(0)) < 1e-4
    # the same backward as two launches: dx (+ planes) on the chain, the column sums trailing
    dx2 = torch.empty(M, Cd, device="cuda")
    dxp2 = L.Planes.empty(M, Cd, "cuda")
    acc = torch.zeros(3, Cd, device="cuda")
    L.layernorm_bwd_dx(dy, x.detach(), g.detach(), st, dx2, dxp2, M, Cd)
    L.layernorm_bwd_cols(dy, x.detach(), st, acc[0], acc[1], M, Cd)
    L.colsum_planes(dxp2, acc[2], accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(dx2, dx) and torch.equal(dxp2.float(), dxp.float())
    assert _rel(acc[0], g.grad) < 1e-5 and _rel(acc[1], b.grad) < 1e-5 and _rel(acc[2], x.grad.sum(0)) < 1e-4


@pytest.mark.parametrize("Tq,Tk", [(20, 288), (80, 80), (7, 36), (5, 12), (3, 1152), (2, 2048), (9, 50), (4, 130)])
def test_softmax_fwd_bwd(L, Tq, Tk):
    """Vectorised kernels (Tk % 4 == 0) and the scalar fall-back (Tk = 50, 130), rows padded to 8 floats as in ops.py."""
    pairs, heads = 2, 3
    rows = pairs * heads * Tq
    ld = (Tk + 7) // 8 * 8
    S = _mk(rows, Tk, 30, 3.0)
    mask = torch.zeros(pairs, Tk, device="cuda")
    mask[1, -(Tk // 8 + 1):] = -10000.0
    Sd = S.double().clone().requires_grad_(True)
    ref = torch.softmax(Sd * 0.125 + mask.double().repeat_interleave(heads * Tq, 0), -1)
    work = torch.full((rows, ld), float("nan"), device="cuda")
    work[:, :Tk] = S
    pp = L.Planes.empty(rows, Tk, "cuda", ld=ld)
    L.softmax_fwd(work, ld, mask, rows, Tk, heads * Tq, 0.125, pp)
    torch.cuda.synchronize()
    assert _rel(work[:, :Tk], ref) < 1e-6 and _rel(pp.float(), ref) < 2e-5
    dP = torch.zeros(rows, ld, device="cuda")
    dP[:, :Tk] = _mk(rows, Tk, 31)
    ref.backward(dP[:, :Tk].double())
    dsp = L.Planes.empty(rows, Tk, "cuda", ld=ld)
    L.softmax_bwd(work, dP, ld, rows, Tk, 0.125, dsp)
    torch.cuda.synchronize()
    assert _rel(dsp.float(), Sd.grad) < 2e-5
    # dropout: forward and backward draw the same mask (keys row * Tk + column), keep-rate ~ 1 - p
    rng = torch.tensor([77, 5], dtype=torch.int64, device="cuda")
    w2 = work.clone()
    w2[:, :Tk] = S
    pd = L.Planes.empty(rows, Tk, "cuda", ld=ld)
    L.softmax_fwd(w2, ld, mask, rows, Tk, heads * Tq, 0.125, pd, 0.25, 9, rng)
    kept = pd.float() != 0
    live = ref > 1e-30
    if int(live.sum()) > 2000:
        assert abs(float(kept[live].float().mean()) - 0.75) < 0.05
    assert _rel(pd.float()[kept], (ref / 0.75)[kept]) < 2e-5
    ones = torch.zeros(rows, ld, device="cuda")
    ones[:, :Tk] = 1.0
    dsd = L.Planes.empty(rows, Tk, "cuda", ld=ld)
    L.softmax_bwd(w2, ones, ld, rows, Tk, 1.0, dsd, 0.25, 9, rng)
    torch.cuda.synchronize()
    P = w2[:, :Tk].double()
    dPm = kept.double() / 0.75                      # what backward must have used as the masked upstream gradient
    want = P * (dPm - (dPm * P).sum(-1, keepdim=True))
    assert _rel(dsd.float(), want) < 2e-5


def test_embeddings_and_colsum(L):
    M, T, H, V = 40, 20, 64, 100
    tok = torch.randint(0, V, (M,), device="cuda")
    seg = torch.randint(0, 2, (M,), device="cuda")
    word, pos, typ = _mk(V, H, 40), _mk(64, H, 41), _mk(2, H, 42)
    out = torch.empty(M, H, device="cuda")
    L.embed_text_fwd(tok, seg, word, pos, typ, out, M, T, H)
    ref = word[tok] + pos[torch.arange(M, device="cuda") % T] + typ[seg]
    assert _rel(out, ref) < 1e-7
    dout = _mk(M, H, 43)
    dw, dp, dt = torch.zeros_like(word), torch.zeros_like(pos), torch.zeros_like(typ)
    L.embed_text_bwd(tok, seg, dout, dw, dp, dt, M, T, H, 0)
    rw = torch.zeros_like(word).index_add_(0, tok, dout * (tok != 0)[:, None])
    rp = torch.zeros_like(pos).index_add_(0, torch.arange(M, device="cuda") % T, dout)
    assert _rel(dw, rw) < 1e-6 and _rel(dp, rp) < 1e-6
    cs = torch.empty(H, device="cuda")
    L.colsum(dout, H, M, H, cs)
    assert _rel(cs, dout.sum(0)) < 1e-6
    # image location embedding
    Hv = 128
    loc = torch.rand(M, 12, device="cuda")
    loc[:, 11] = torch.randint(0, 8, (M,), device="cuda").float()
    w5, b5, w4, b4, w2, b2, sq = (_mk(Hv, 5, 44), _mk(1, Hv, 45)[0].contiguous(), _mk(Hv, 4, 46),
                                  _mk(1, Hv, 47)[0].contiguous(), _mk(Hv, 2, 48), _mk(1, Hv, 49)[0].contiguous(),
                                  _mk(32, Hv, 50))
    o = torch.empty(M, Hv, device="cuda")
    L.embed_loc_fwd(loc, w5, b5, w4, b4, w2, b2, sq, o, M, Hv)
    ref = loc[:, :5] @ w5.t() + b5 + loc[:, 5:9] @ w4.t() + b4 + loc[:, 9:11] @ w2.t() + b2 + sq[loc[:, 11].long()]
    assert _rel(o, ref) < 1e-6
    do = _mk(M, Hv, 51)
    grads = [torch.zeros_like(t) for t in (w5, b5, w4, b4, w2, b2, sq)]
    L.embed_loc_bwd(loc, do, *grads, M, Hv)
    assert _rel(grads[0], do.t() @ loc[:, :5]) < 1e-5
    assert _rel(grads[2], do.t() @ loc[:, 5:9]) < 1e-5
    assert _rel(grads[4], do.t() @ loc[:, 9:11]) < 1e-5
    assert _rel(grads[1], do.sum(0)) < 1e-5
    assert _rel(grads[6], torch.zeros_like(sq).index_add_(0, loc[:, 11].long(), do)) < 1e-5


def test_losses(L):
    rows, V = 64, 30522
    ld = 30528
    logits = torch.zeros(rows, ld, device="cuda")
    logits[:, :V] = _mk(rows, V, 60, 2.0)
    tgt = torch.randint(0, V, (rows,), device="cuda")
    tgt[::3] = -1
    lg = logits[:, :V].double().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lg, tgt, ignore_index=-1)
    ref.backward()
    ls, cnt = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    L.ce_loss(logits, ld, tgt, rows, V, ls, cnt)
    assert abs(float(ls / cnt) - float(ref)) < 1e-5 * abs(float(ref))
    dl = torch.zeros(rows, ld, device="cuda")
    dlp = L.Planes.empty(rows, V, "cuda", ld=ld)
    L.ce_grad(logits, ld, tgt, rows, V, cnt, None, dl, dlp)
    torch.cuda.synchronize()
    assert _rel(dl[:, :V], lg.grad) < 1e-5 and _rel(dlp.float(), lg.grad) < 2e-5
    # KL
    rows, Cc, ldc = 50, 1601, 1608
    logits = torch.zeros(rows, ldc, device="cuda")
    logits[:, :Cc] = _mk(rows, Cc, 61, 2.0)
    target = torch.softmax(_mk(rows, Cc, 62), -1).contiguous()
    mask = (torch.rand(rows, device="cuda") < 0.5).long()
    lg = logits[:, :Cc].double().clone().requires_grad_(True)
    kl = torch.nn.functional.kl_div(torch.log_softmax(lg, -1), target.double(), reduction="none")
    ref = (kl * mask[:, None]).sum() / max(1, int(mask.sum()))
    ref.backward()
    ls, cnt = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    L.kl_loss(logits, ldc, target, Cc, mask, rows, Cc, ls, cnt)
    assert abs(float(ls / cnt.clamp_min(1)) - float(ref)) < 1e-5 * abs(float(ref))
    dl = torch.zeros(rows, ldc, device="cuda")
    L.kl_grad(logits, ldc, target, Cc, mask, rows, Cc, cnt, None, dl, None)
    torch.cuda.synchronize()
    assert _rel(dl[:, :Cc], lg.grad) < 1e-5


@pytest.mark.parametrize("shape", [(640, 768, 3072), (640, 768, 30522 // 8 * 8), (1024, 1024, 2304), (256, 128, 512)])
def test_gemm_split_k_linear_epilogue(L, shape):
    """Few-tile / long-K problems take the split-K path (vector reductions into a zeroed f32 output)."""
    M, N, K = shape
    A, B = _mk(M, K, 70), _mk(N, K, 71, 0.05)
    bias = _mk(1, N, 72)[0].contiguous()
    res = _mk(M, N, 73)
    out, _, _ = _run_gemm(L, A, B, False, False, 3, bias=bias, residual=res)
    ref = A.double() @ B.double().t() + bias.double() + res.double()
    assert _rel(out, ref) < 3e-5
    # MN-major operands (wgrad form) through the same path
    out, _, _ = _run_gemm(L, A, B, True, True, 3)
    assert _rel(out, A.double() @ B.double().t()) < 3e-5


@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (8, 3), (5, 4)])
def test_mean_chunks_matches_rank_ordered_sum(L, world, rank):
    """Chunk mean of the copy-engine gradient exchange: bit-identical to summing the ranks' copies in rank order."""
    from yvb200.step import GradientExchange as GE
    n = 4 * 1237
    torch.manual_seed(world * 10 + rank)
    stage = torch.randn(world - 1, n + 64, device="cuda")
    own = torch.randn(n, device="cuda")
    want = torch.cat([own.clone(), torch.zeros(0, device="cuda")])
    sl = torch.zeros(world * n, device="cuda")
    sl[rank * n:(rank + 1) * n] = own
    GE.ce_reduce(rank, world, sl, n, stage[:, :n])
    want = sl[rank * n:(rank + 1) * n]
    L.mean_chunks(own, stage, world, rank)
    torch.cuda.synchronize()
    assert torch.equal(own, want)
