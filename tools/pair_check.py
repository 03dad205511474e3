"""Developer tool: correctness + timing of the yv_gemm variants on the step's main GEMM shapes (one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youtube-vln_b200"))
import torch
from yvb200 import lib as L

torch.manual_seed(0)
FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def planes(x):
    return L.split_planes(x.contiguous())


def case(name, M, N, K, a_mn=False, b_mn=False, want_planes=False, want32=True, bias=True, act=0, variants=(32, 64, 128, 256, 0)):
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") * 0.05
    pa = planes(A.t() if a_mn else A)
    pb = planes(B.t() if b_mn else B)
    bv = torch.randn(N, device="cuda") if bias else None
    ref = A.double() @ B.double().t() + (bv.double() if bias else 0)
    res = []
    for v in variants:
        L.set_gemm_variant(v)
        ldN = (N + 7) // 8 * 8
        out = torch.full((M, ldN), float("nan"), device="cuda") if want32 else None
        pl = L.Planes.empty(M, N, "cuda") if want_planes else None
        def run():
            L.gemm(M, N, K, L.op_of(pa, a_mn), L.op_of(pb, b_mn), passes=3, bias=bv, act=act, out32=out, ld_out=ldN,
                   out_planes=pl.ptr() if pl else None, ld_pl=pl.ld if pl else 0,
                   pl_plane_stride=pl.plane_stride if pl else 0)
        try:
            run()
            torch.cuda.synchronize()
        except Exception as e:
            res.append(f"v{v}: ERROR {str(e)[:120]}")
            continue
        got = out[:, :N] if want32 else pl.float()
        err = float((got.double() - ref).norm() / ref.norm())
        # warm back-to-back
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 20 * 1e3
        # cold: L2 flushed before each launch
        tc = 0.0
        for _ in range(5):
            FLUSH.fill_(1)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            tc += e0.elapsed_time(e1)
        cold = tc / 5 * 1e3
        tf = 2.0 * M * N * K / (warm * 1e-6) / 1e12
        res.append(f"v{v}: err {err:.1e} warm {warm:6.1f}us ({tf:5.0f} algTF/s) cold {cold:6.1f}us")
    L.set_gemm_variant(0)
    print(f"{name} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} planes={int(want_planes)} f32={int(want32)}")
    for r in res:
        print("    " + r)
    sys.stdout.flush()


if __name__ == "__main__":
    case("vision dense fwd", 2304, 1024, 1024, want_planes=True)
    case("vision qkv fwd", 2304, 3072, 1024, want_planes=True, want32=False)
    case("vision dgrad", 2304, 1024, 1024, b_mn=True)
    case("vision qkv dgrad", 2304, 1024, 3072, b_mn=True)
    case("vision wgrad", 1024, 1024, 2304, a_mn=True, b_mn=True, bias=False)
    case("vision qkv wgrad", 3072, 1024, 2304, a_mn=True, b_mn=True, bias=False)
    case("img embed", 2304, 1024, 2048)
    case("text ffn1 fwd", 640, 3072, 768, want_planes=True, act=1)
    case("text ffn2 fwd", 640, 768, 3072)
    case("text qkv fwd", 640, 2304, 768, want_planes=True, want32=False)
    case("text ffn1 wgrad", 3072, 768, 640, a_mn=True, b_mn=True, bias=False)
    case("text ffn2 wgrad", 768, 3072, 640, a_mn=True, b_mn=True, bias=False)
    case("lm head fwd", 640, 30522, 768)
    case("lm head dgrad", 640, 768, 30522, b_mn=True, bias=False)
    case("lm head wgrad", 30522, 768, 640, a_mn=True, b_mn=True, bias=False)
    case("img head fwd", 2304, 1601, 1024)
