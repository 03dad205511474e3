"""ctypes binding of libyvb200.so (the C ABI declared in include/yvb200.h).

PyTorch is used for device memory and streams only: every function here passes raw device pointers and
the current CUDA stream to the library.  A missing / unloadable library is a hard error on the CUDA path
(no fallback): ``load()`` raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyvb200.so")

ACT_NONE, ACT_GELU, ACT_RELU, ACT_MUL_GELU_GRAD, ACT_MUL_RELU_MASK = 0, 1, 2, 3, 4

#: every symbol include/yvb200.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "yv_last_error", "yv_version", "yv_launch_count", "yv_gemm", "yv_gemm_splits", "yv_gemm_set_variant", "yv_split_planes", "yv_split_multi",
    "yv_rng_advance", "yv_layernorm_fwd", "yv_layernorm_bwd", "yv_layernorm_bwd_dx", "yv_layernorm_bwd_cols", "yv_softmax_fwd", "yv_softmax_bwd",
    "yv_embed_text_fwd", "yv_embed_text_bwd", "yv_embed_loc_fwd", "yv_embed_loc_bwd", "yv_colsum",
    "yv_colsum_planes", "yv_act_bwd_split", "yv_adamw_multi", "yv_ce_loss", "yv_ce_grad", "yv_kl_loss", "yv_kl_grad",
    "yv_mask_tokens", "yv_mask_regions", "yv_attn_supported", "yv_attn_fwd", "yv_attn_bwd", "yv_attn_bwd_workspace_bytes", "yv_mean_chunks",
]


class YvOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("inner", C.c_int64), ("rows", C.c_int64), ("ld", C.c_int64),
                ("nb0", C.c_int64), ("sb0", C.c_int64), ("nb1", C.c_int64), ("sb1", C.c_int64),
                ("plane_stride", C.c_int64), ("mn_major", C.c_int32), ("_pad", C.c_int32)]


class YvGemm(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("passes", C.c_int32),
                ("a", YvOperand), ("b", YvOperand),
                ("alpha", C.c_float), ("act", C.c_int32),
                ("bias", C.c_void_p), ("aux_out", C.c_void_p), ("aux_in", C.c_void_p),
                ("residual", C.c_void_p), ("out32", C.c_void_p),
                ("ld_out", C.c_int64), ("out_sb0", C.c_int64), ("out_sb1", C.c_int64),
                ("out_planes", C.c_void_p),
                ("ld_pl", C.c_int64), ("pl_sb0", C.c_int64), ("pl_sb1", C.c_int64), ("pl_plane_stride", C.c_int64),
                ("drop_p", C.c_float), ("drop_site", C.c_uint32), ("rng", C.c_void_p),
                ("out32_zeroed", C.c_int32), ("_pad2", C.c_int32)]


class YvHeadView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("plane_stride", C.c_int64), ("pair_stride", C.c_int64),
                ("rows", C.c_int32), ("_pad", C.c_int32)]


class YvAttnFwd(C.Structure):
    _fields_ = [("pairs", C.c_int32), ("heads", C.c_int32), ("dh", C.c_int32), ("passes", C.c_int32),
                ("q", YvHeadView), ("k", YvHeadView), ("v", YvHeadView),
                ("mask", C.c_void_p), ("scale", C.c_float), ("drop_p", C.c_float), ("drop_site", C.c_uint32),
                ("_pad", C.c_uint32), ("rng", C.c_void_p),
                ("out_planes", C.c_void_p), ("ld_out", C.c_int64), ("out_plane_stride", C.c_int64),
                ("out32", C.c_void_p), ("ld_out32", C.c_int64), ("lse", C.c_void_p)]


class YvAttnBwd(C.Structure):
    _fields_ = [("pairs", C.c_int32), ("heads", C.c_int32), ("dh", C.c_int32), ("passes", C.c_int32),
                ("q", YvHeadView), ("k", YvHeadView), ("v", YvHeadView), ("dout", YvHeadView), ("out", YvHeadView),
                ("mask", C.c_void_p), ("scale", C.c_float), ("drop_p", C.c_float), ("drop_site", C.c_uint32),
                ("_pad", C.c_uint32), ("rng", C.c_void_p), ("lse", C.c_void_p),
                ("dq", YvHeadView), ("dk", YvHeadView), ("dv", YvHeadView),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("tickets", C.c_void_p)]


class YvSplitSeg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_off", C.c_int64), ("numel", C.c_int64), ("first_blk", C.c_int64)]


_lib = None


def load():
    """Load libyvb200.so once; raise loudly if it is missing (the CUDA path has no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"yvb200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C youtube-vln_b200/csrc`). The CUDA path has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.yv_last_error.restype = C.c_char_p
    lib.yv_version.restype = C.c_int
    lib.yv_launch_count.restype = C.c_uint64
    if hasattr(lib, "yv_attn_bwd_workspace_bytes"):
        lib.yv_attn_bwd_workspace_bytes.restype = C.c_size_t
    for name in SYMBOLS:
        if not hasattr(lib, name):
            raise RuntimeError(f"yvb200: {LIB_PATH} does not export {name}")
    _lib = lib
    return _lib


def available() -> bool:
    return os.path.exists(LIB_PATH)


def launch_count() -> int:
    return int(load().yv_launch_count())


def set_gemm_variant(variant: int = 0):
    """Tuning / test knob: 0 automatic, 32 / 64 single-CTA kernels, 2 CTA pairs, 128 / 256 CTA pairs of that width."""
    _check(load().yv_gemm_set_variant(C.c_int(variant)), "gemm_set_variant")
    _GEMM_VARIANT[0] = variant


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"yvb200 {what} failed: {load().yv_last_error().decode()}")


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Planes:
    """A bf16 hi/lo plane pair holding a 2-D fp32 matrix ``[rows, cols]`` split as hi + lo.

    ``addr`` is the device address of element (0, 0) of the hi plane, the lo plane starts
    ``plane_stride`` elements later; ``keep`` pins the owning torch storage."""
    __slots__ = ("keep", "addr", "rows", "cols", "ld", "plane_stride")

    def __init__(self, keep: torch.Tensor, addr: int, rows: int, cols: int, ld: int, plane_stride: int):
        self.keep, self.addr, self.rows, self.cols, self.ld, self.plane_stride = keep, addr, rows, cols, ld, plane_stride

    @staticmethod
    def empty(rows: int, cols: int, device, ld: Optional[int] = None) -> "Planes":
        ld = ld if ld is not None else (cols + 7) // 8 * 8
        t = torch.empty((2, rows, ld), dtype=torch.bfloat16, device=device)
        return Planes(t, t.data_ptr(), rows, cols, ld, rows * ld)

    def ptr(self, elem_off: int = 0) -> int:
        return self.addr + 2 * elem_off

    def float(self) -> torch.Tensor:
        """hi + lo as fp32 [rows, cols] (debug / tests; only for planes created by ``empty``)."""
        t = self.keep.view(2, self.rows, self.ld)
        return (t[0].float() + t[1].float())[:, : self.cols]


def operand(ptr: int, inner: int, rows: int, ld: int, plane_stride: int, mn_major: bool = False,
            nb0: int = 1, sb0: int = 0, nb1: int = 1, sb1: int = 0) -> YvOperand:
    return YvOperand(ptr, inner, rows, ld, nb0, sb0, nb1, sb1, plane_stride, 1 if mn_major else 0, 0)


def op_of(p: Planes, mn_major: bool = False) -> YvOperand:
    """Whole plane pair as an operand: K-major (rows x cols, contraction over cols) or its transpose."""
    return operand(p.ptr(), p.cols, p.rows, p.ld, p.plane_stride, mn_major)


def gemm(M: int, N: int, K: int, a: YvOperand, b: YvOperand, *, passes: int = 3, alpha: float = 1.0,
         act: int = ACT_NONE, bias=None, aux_out=None, aux_in=None, residual=None, out32=None, ld_out: int = 0,
         out_sb0: int = 0, out_sb1: int = 0, out_planes: Optional[int] = None, ld_pl: int = 0, pl_sb0: int = 0,
         pl_sb1: int = 0, pl_plane_stride: int = 0, drop_p: float = 0.0, drop_site: int = 0, rng=None,
         out32_zeroed: bool = False):
    """Raw yv_gemm call.  out32/aux/residual/bias are tensors (or None); out_planes is a device address."""
    g = YvGemm()
    g.M, g.N, g.K, g.passes = M, N, K, passes
    g.a, g.b = a, b
    g.alpha, g.act = alpha, act
    g.bias, g.aux_out, g.aux_in, g.residual, g.out32 = _p(bias), _p(aux_out), _p(aux_in), _p(residual), _p(out32)
    g.ld_out, g.out_sb0, g.out_sb1 = ld_out, out_sb0, out_sb1
    g.out_planes = out_planes
    g.ld_pl, g.pl_sb0, g.pl_sb1, g.pl_plane_stride = ld_pl, pl_sb0, pl_sb1, pl_plane_stride
    g.drop_p, g.drop_site, g.rng = drop_p, drop_site, _p(rng)
    g.out32_zeroed = 1 if out32_zeroed else 0
    _check(load().yv_gemm(C.byref(g), _stream()), "gemm")


_SPLIT_CACHE = {}
_GEMM_VARIANT = [0]


def will_split(M: int, N: int, K: int) -> bool:
    """Would an un-batched yv_gemm with a linear epilogue into a contiguous, aligned fp32 [M, N] output split K?
    (Such launches reduce into a zero-filled output; callers use this to zero-fill early, off the dependency chain.)"""
    key = (M, N, K, _GEMM_VARIANT[0])
    r = _SPLIT_CACHE.get(key)
    if r is None:
        g = YvGemm()
        g.M, g.N, g.K, g.passes = M, N, K, 3
        g.a = operand(0x1000, K, M, K, M * K)
        g.b = operand(0x1000, K, N, K, N * K)
        g.alpha, g.act = 1.0, ACT_NONE
        g.out32, g.ld_out = 0x1000, N
        r = _SPLIT_CACHE[key] = int(load().yv_gemm_splits(C.byref(g))) > 1
    return r


def split_planes(src: torch.Tensor, dst: Optional[Planes] = None) -> Planes:
    """fp32 [rows, cols] (row stride arbitrary) -> Planes."""
    assert src.dim() == 2 and src.dtype == torch.float32 and src.stride(1) == 1
    rows, cols = src.shape
    if dst is None:
        dst = Planes.empty(rows, cols, src.device)
    _check(load().yv_split_planes(C.c_void_p(src.data_ptr()), C.c_int64(src.stride(0)), C.c_void_p(dst.ptr()),
                                  C.c_int64(dst.ld), C.c_int64(dst.plane_stride), C.c_int64(rows), C.c_int64(cols),
                                  _stream()), "split_planes")
    return dst


def split_multi(segs_dev: torch.Tensor, nseg: int, total_blocks: int, planes: torch.Tensor, plane_stride: int):
    _check(load().yv_split_multi(C.c_void_p(segs_dev.data_ptr()), C.c_int32(nseg), C.c_int64(total_blocks),
                                 C.c_void_p(planes.data_ptr()), C.c_int64(plane_stride), _stream()), "split_multi")


def adamw_multi(segs_dev: torch.Tensor, nseg: int, total_blocks: int, hyper_dev: torch.Tensor):
    _check(load().yv_adamw_multi(C.c_void_p(segs_dev.data_ptr()), C.c_int32(nseg), C.c_int64(total_blocks),
                                 C.c_void_p(hyper_dev.data_ptr()), _stream()), "adamw_multi")


def rng_advance(rng: torch.Tensor):
    _check(load().yv_rng_advance(C.c_void_p(rng.data_ptr()), _stream()), "rng_advance")


def layernorm_fwd(x, gamma, beta, eps, y32, y_planes: Optional[Planes], stats, M, Cdim, drop_p=0.0, drop_site=0,
                  rng=None):
    _check(load().yv_layernorm_fwd(C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()), C.c_void_p(beta.data_ptr()),
                                   C.c_float(eps), C.c_void_p(_p(y32)),
                                   C.c_void_p(y_planes.ptr() if y_planes is not None else None),
                                   C.c_int64(y_planes.plane_stride if y_planes is not None else 0),
                                   C.c_void_p(_p(stats)), C.c_int64(M), C.c_int32(Cdim), C.c_float(drop_p),
                                   C.c_uint32(drop_site), C.c_void_p(_p(rng)), _stream()), "layernorm_fwd")


def layernorm_bwd(dy, x, gamma, stats, dx32, dx_planes: Optional[Planes], dgamma, dbeta, M, Cdim, *, post_drop_p=0.0,
                  post_drop_site=0, dx_add=None, pre_drop_p=0.0, pre_drop_site=0, rng=None, dbias=None):
    _check(load().yv_layernorm_bwd(C.c_void_p(dy.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                   C.c_void_p(stats.data_ptr()), C.c_float(post_drop_p), C.c_uint32(post_drop_site),
                                   C.c_void_p(_p(dx_add)), C.c_void_p(_p(dx32)),
                                   C.c_void_p(dx_planes.ptr() if dx_planes is not None else None),
                                   C.c_int64(dx_planes.plane_stride if dx_planes is not None else 0),
                                   C.c_float(pre_drop_p), C.c_uint32(pre_drop_site), C.c_void_p(_p(rng)),
                                   C.c_void_p(_p(dgamma)), C.c_void_p(_p(dbeta)), C.c_void_p(_p(dbias)), C.c_int64(M),
                                   C.c_int32(Cdim), _stream()), "layernorm_bwd")


def layernorm_bwd_dx(dy, x, gamma, stats, dx32, dx_planes: Optional[Planes], M, Cdim, *, dx_add=None, pre_drop_p=0.0,
                     pre_drop_site=0, rng=None):
    _check(load().yv_layernorm_bwd_dx(C.c_void_p(dy.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(gamma.data_ptr()),
                                      C.c_void_p(stats.data_ptr()), C.c_void_p(_p(dx_add)), C.c_void_p(_p(dx32)),
                                      C.c_void_p(dx_planes.ptr() if dx_planes is not None else None),
                                      C.c_int64(dx_planes.plane_stride if dx_planes is not None else 0),
                                      C.c_float(pre_drop_p), C.c_uint32(pre_drop_site), C.c_void_p(_p(rng)), C.c_int64(M),
                                      C.c_int32(Cdim), _stream()), "layernorm_bwd_dx")


def layernorm_bwd_cols(dy, x, stats, dgamma, dbeta, M, Cdim):
    _check(load().yv_layernorm_bwd_cols(C.c_void_p(dy.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(stats.data_ptr()),
                                        C.c_void_p(dgamma.data_ptr()), C.c_void_p(dbeta.data_ptr()), C.c_int64(M),
                                        C.c_int32(Cdim), _stream()), "layernorm_bwd_cols")


def softmax_fwd(s, ld_s, mask, rows, cols, rows_per_pair, scale, p_planes: Planes, drop_p=0.0, drop_site=0, rng=None):
    _check(load().yv_softmax_fwd(C.c_void_p(s.data_ptr()), C.c_int64(ld_s), C.c_void_p(_p(mask)), C.c_int64(rows),
                                 C.c_int32(cols), C.c_int64(rows_per_pair), C.c_float(scale), C.c_void_p(p_planes.ptr()),
                                 C.c_int64(p_planes.ld), C.c_int64(p_planes.plane_stride), C.c_float(drop_p),
                                 C.c_uint32(drop_site), C.c_void_p(_p(rng)), _stream()), "softmax_fwd")


def softmax_bwd(p, dpd, ld_s, rows, cols, scale, ds_planes: Planes, drop_p=0.0, drop_site=0, rng=None):
    _check(load().yv_softmax_bwd(C.c_void_p(p.data_ptr()), C.c_void_p(dpd.data_ptr()), C.c_int64(ld_s), C.c_int64(rows),
                                 C.c_int32(cols), C.c_float(scale), C.c_void_p(ds_planes.ptr()), C.c_int64(ds_planes.ld),
                                 C.c_int64(ds_planes.plane_stride), C.c_float(drop_p), C.c_uint32(drop_site),
                                 C.c_void_p(_p(rng)), _stream()), "softmax_bwd")


def embed_text_fwd(tok, seg, word, pos, typ, out, M, T, H):
    _check(load().yv_embed_text_fwd(C.c_void_p(tok.data_ptr()), C.c_void_p(seg.data_ptr()), C.c_void_p(word.data_ptr()),
                                    C.c_void_p(pos.data_ptr()), C.c_void_p(typ.data_ptr()), C.c_void_p(out.data_ptr()),
                                    C.c_int64(M), C.c_int32(T), C.c_int32(H), _stream()), "embed_text_fwd")


def embed_text_bwd(tok, seg, dout, dword, dpos, dtyp, M, T, H, padding_idx=0):
    _check(load().yv_embed_text_bwd(C.c_void_p(tok.data_ptr()), C.c_void_p(seg.data_ptr()), C.c_void_p(dout.data_ptr()),
                                    C.c_void_p(_p(dword)), C.c_void_p(_p(dpos)), C.c_void_p(_p(dtyp)), C.c_int64(M),
                                    C.c_int32(T), C.c_int32(H), C.c_int32(padding_idx), _stream()), "embed_text_bwd")


def embed_loc_fwd(loc, w5, b5, w4, b4, w2, b2, seq, out, M, H):
    _check(load().yv_embed_loc_fwd(*[C.c_void_p(t.data_ptr()) for t in (loc, w5, b5, w4, b4, w2, b2, seq, out)],
                                   C.c_int64(M), C.c_int32(H), _stream()), "embed_loc_fwd")


def embed_loc_bwd(loc, dout, dw5, db5, dw4, db4, dw2, db2, dseq, M, H):
    _check(load().yv_embed_loc_bwd(*[C.c_void_p(t.data_ptr()) for t in (loc, dout, dw5, db5, dw4, db4, dw2, db2, dseq)],
                                   C.c_int64(M), C.c_int32(H), _stream()), "embed_loc_bwd")


def colsum(x, ld, rows, cols, out, accumulate=False):
    _check(load().yv_colsum(C.c_void_p(x.data_ptr()), C.c_int64(ld), C.c_int64(rows), C.c_int32(cols),
                            C.c_void_p(out.data_ptr()), C.c_int32(1 if accumulate else 0), _stream()), "colsum")


def mean_chunks(own: torch.Tensor, stage: Optional[torch.Tensor], world: int, rank: int):
    """own <- mean over the ranks (summed in rank order) of own and the ``world - 1`` staged peer copies (gradient
    exchange over peer-to-peer copies: row step-1 of ``stage`` came from rank (rank - step) mod world)."""
    if not (own.is_cuda and own.is_contiguous() and own.dtype == torch.float32):
        raise RuntimeError("mean_chunks: needs a contiguous CUDA float32 chunk")
    if stage is not None and not (stage.is_cuda and stage.dtype == torch.float32 and stage.stride(1) == 1
                                  and stage.shape[1] >= own.numel() and stage.shape[0] >= world - 1):
        raise RuntimeError("mean_chunks: staging rows do not cover the chunk")
    _check(load().yv_mean_chunks(C.c_void_p(own.data_ptr()), C.c_void_p(_p(stage)),
                                 C.c_int64(stage.stride(0) if stage is not None else 0), C.c_int32(world), C.c_int32(rank),
                                 C.c_int64(own.numel()), _stream()), "mean_chunks")


def colsum_planes(p: Planes, out, accumulate=False):
    _check(load().yv_colsum_planes(C.c_void_p(p.ptr()), C.c_int64(p.ld), C.c_int64(p.plane_stride), C.c_int64(p.rows),
                                   C.c_int32(p.cols), C.c_void_p(out.data_ptr()), C.c_int32(1 if accumulate else 0),
                                   _stream()), "colsum_planes")


def act_bwd_split(dy: torch.Tensor, aux: Optional[torch.Tensor], act: int, dst: Planes, dbias=None):
    """planes = dy * act'(aux); dy / aux are 2-D f32 with unit column stride; dbias (optional) = column sums."""
    rows, cols = dy.shape
    _check(load().yv_act_bwd_split(C.c_void_p(dy.data_ptr()), C.c_int64(dy.stride(0)), C.c_void_p(_p(aux)),
                                   C.c_int64(aux.stride(0) if aux is not None else 0), C.c_int32(act), C.c_void_p(dst.ptr()),
                                   C.c_int64(dst.ld), C.c_int64(dst.plane_stride), C.c_int64(rows), C.c_int64(cols),
                                   C.c_void_p(_p(dbias)), _stream()), "act_bwd_split")


def ce_loss(logits, ld, target, rows, cols, loss_sum, count):
    _check(load().yv_ce_loss(C.c_void_p(logits.data_ptr()), C.c_int64(ld), C.c_void_p(target.data_ptr()), C.c_int64(rows),
                             C.c_int32(cols), C.c_void_p(loss_sum.data_ptr()), C.c_void_p(count.data_ptr()), _stream()),
           "ce_loss")


def ce_grad(logits, ld, target, rows, cols, count, gscale, dl32, dl_planes: Optional[Planes]):
    _check(load().yv_ce_grad(C.c_void_p(logits.data_ptr()), C.c_int64(ld), C.c_void_p(target.data_ptr()), C.c_int64(rows),
                             C.c_int32(cols), C.c_void_p(count.data_ptr()), C.c_void_p(_p(gscale)), C.c_void_p(_p(dl32)),
                             C.c_void_p(dl_planes.ptr() if dl_planes is not None else None),
                             C.c_int64(dl_planes.ld if dl_planes is not None else 0),
                             C.c_int64(dl_planes.plane_stride if dl_planes is not None else 0), _stream()), "ce_grad")


def kl_loss(logits, ld, target, ld_t, mask, rows, cols, loss_sum, count):
    _check(load().yv_kl_loss(C.c_void_p(logits.data_ptr()), C.c_int64(ld), C.c_void_p(target.data_ptr()), C.c_int64(ld_t),
                             C.c_void_p(mask.data_ptr()), C.c_int64(rows), C.c_int32(cols), C.c_void_p(loss_sum.data_ptr()),
                             C.c_void_p(count.data_ptr()), _stream()), "kl_loss")


def kl_grad(logits, ld, target, ld_t, mask, rows, cols, count, gscale, dl32, dl_planes: Optional[Planes]):
    _check(load().yv_kl_grad(C.c_void_p(logits.data_ptr()), C.c_int64(ld), C.c_void_p(target.data_ptr()), C.c_int64(ld_t),
                             C.c_void_p(mask.data_ptr()), C.c_int64(rows), C.c_int32(cols), C.c_void_p(count.data_ptr()),
                             C.c_void_p(_p(gscale)), C.c_void_p(_p(dl32)),
                             C.c_void_p(dl_planes.ptr() if dl_planes is not None else None),
                             C.c_int64(dl_planes.ld if dl_planes is not None else 0),
                             C.c_int64(dl_planes.plane_stride if dl_planes is not None else 0), _stream()), "kl_grad")


def mask_tokens(tokens, mask_u8, p, random_tokens, forced_u8, mask_id: int, targets):
    _check(load().yv_mask_tokens(C.c_void_p(tokens.data_ptr()), C.c_void_p(mask_u8.data_ptr()), C.c_void_p(p.data_ptr()),
                                 C.c_void_p(random_tokens.data_ptr()), C.c_void_p(_p(forced_u8)), C.c_int64(mask_id),
                                 C.c_void_p(targets.data_ptr()), C.c_int64(tokens.numel()), _stream()), "mask_tokens")


def mask_regions(features, probs, mask, p, targets, targets_mask, rows: int, F: int, Cc: int):
    _check(load().yv_mask_regions(C.c_void_p(features.data_ptr()), C.c_void_p(probs.data_ptr()), C.c_void_p(mask.data_ptr()),
                                  C.c_void_p(p.data_ptr()), C.c_void_p(targets.data_ptr()),
                                  C.c_void_p(targets_mask.data_ptr()), C.c_int64(rows), C.c_int32(F), C.c_int32(Cc),
                                  _stream()), "mask_regions")


def head_view(p: Planes, col_off: int, rows_per_pair: int) -> YvHeadView:
    """Columns [col_off, col_off + heads*dh) of a plane pair [pairs*rows_per_pair, ld] as [pairs, rows, heads*dh]."""
    return YvHeadView(p.ptr(col_off), p.ld, p.plane_stride, rows_per_pair * p.ld, rows_per_pair, 0)


def attn_supported(dh: int, passes: int) -> bool:
    return bool(load().yv_attn_supported(C.c_int32(dh), C.c_int32(passes)))


def attn_bwd_workspace_bytes(pairs: int, heads: int, dh: int, Tq: int, Tk: int) -> int:
    return int(load().yv_attn_bwd_workspace_bytes(C.c_int32(pairs), C.c_int32(heads), C.c_int32(dh), C.c_int32(Tq),
                                                  C.c_int32(Tk)))


def attn_fwd(q: YvHeadView, k: YvHeadView, v: YvHeadView, mask, pairs: int, heads: int, dh: int, scale: float,
             out: Planes, out32=None, lse=None, passes: int = 3, drop_p: float = 0.0, drop_site: int = 0, rng=None):
    """Fused softmax(Q K^T * scale + mask) V (yv_attn_fwd); ``out`` receives the merged-head context as planes."""
    a = YvAttnFwd()
    a.pairs, a.heads, a.dh, a.passes = pairs, heads, dh, passes
    a.q, a.k, a.v = q, k, v
    a.mask, a.scale, a.drop_p, a.drop_site, a.rng = _p(mask), scale, drop_p, drop_site, _p(rng)
    a.out_planes, a.ld_out, a.out_plane_stride = out.ptr(), out.ld, out.plane_stride
    a.out32, a.ld_out32 = _p(out32), (out32.stride(0) if out32 is not None else 0)
    a.lse = _p(lse)
    _check(load().yv_attn_fwd(C.byref(a), _stream()), "attn_fwd")


def attn_bwd(q: YvHeadView, k: YvHeadView, v: YvHeadView, dout: YvHeadView, out: YvHeadView, mask, lse, pairs: int,
             heads: int, dh: int, scale: float, dq: YvHeadView, dk: YvHeadView, dv: YvHeadView, workspace: torch.Tensor,
             tickets: torch.Tensor, passes: int = 3, drop_p: float = 0.0, drop_site: int = 0, rng=None):
    """Fused attention backward (yv_attn_bwd); ``workspace`` is a byte tensor of ``attn_bwd_workspace_bytes`` bytes
    (contents irrelevant), ``tickets`` an int32 tensor of pairs*heads zeros (left zero by the kernel)."""
    a = YvAttnBwd()
    a.pairs, a.heads, a.dh, a.passes = pairs, heads, dh, passes
    a.q, a.k, a.v, a.dout, a.out = q, k, v, dout, out
    a.mask, a.scale, a.drop_p, a.drop_site, a.rng = _p(mask), scale, drop_p, drop_site, _p(rng)
    a.lse = _p(lse)
    a.dq, a.dk, a.dv = dq, dk, dv
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    assert tickets.numel() >= pairs * heads and tickets.element_size() == 4
    a.tickets = tickets.data_ptr()
    _check(load().yv_attn_bwd(C.byref(a), _stream()), "attn_bwd")
