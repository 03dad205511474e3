// Pieces shared by the tcgen05 GEMM variants (yv_gemm.cu: cta_group::1; yv_gemm_pair.cu: cta_group::2 CTA pairs):
// kernel parameters, PTX wrappers, shared-memory descriptors, the fused epilogue of one 32x32 accumulator chunk
// and the host-side tensor-map builder.  sm_100a only.
#pragma once
#ifndef YV_DBG_EPI
#define YV_DBG_EPI 0
#endif
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

namespace {

// developer timing (tools/gemm_timing.cu): clock64 stamps of CTA 0
#ifdef YV_GEMM_TIMING
__device__ long long yv_dbg[32];
#define YV_T(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) yv_dbg[i] = clock64(); } while (0)
#define YV_T64(i) do { if (blockIdx.x == 0 && threadIdx.x == 64) yv_dbg[i] = clock64(); } while (0)
#else
#define YV_T(i)
#define YV_T64(i)
#endif

struct KParams {
    int M, N, K;
    int nb0;
    int splits, kb_per_split;   // split-K (only for un-batched launches with a linear, f32-only epilogue)
    int total_tiles;            // tiles_m * tiles_n * batch * splits, walked persistently
    int a_mn, b_mn;
    float alpha;
    int act;
    const float* bias;
    float* aux_out;
    const float* aux_in;
    const float* residual;
    float* out32;
    long long ld_out, out_sb0, out_sb1;
    __nv_bfloat16* out_planes;
    long long ld_pl, pl_sb0, pl_sb1, pl_plane_stride;
    float drop_p;
    unsigned drop_site;
    const unsigned long long* rng;
    int pair_n, stages;         // CTA-pair variant only: N extent of the pair tile (128 / 256), depth of the operand ring
};

// ------------------------------------------------------------------------------------------- PTX
YV_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

YV_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
YV_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
YV_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s: a protocol bug must not hang the GPU
    }
}
YV_DEVINL void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                           int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
        "%7}], [%2];" ::"r"(dst),
        "l"((unsigned long long)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
YV_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
YV_DEVINL void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
YV_DEVINL void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (sm_100 "version 1")
//   K-major  (64B swizzle) : rows of 64 B, 8-row groups 512 B apart (SBO); LBO unused
//   MN-major (128B swizzle): 64-element chunks along M/N are LBO bytes apart, 8-k-row groups 1024 B apart (SBO)
YV_DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (Blackwell)
    d |= (uint64_t)layout << 61;   // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}

// one output element through the whole epilogue (ragged tile edges and unaligned leading dimensions only)
__device__ __noinline__ void epilogue_scalar(const KParams& p, const YvDrop& drop, float acc, int n, int z, int row,
                                             long long obase, long long pbase) {
    if (n >= p.N) return;
    float x = p.alpha * acc;
    if (p.bias) x += __ldg(p.bias + n);
    if (p.aux_out) p.aux_out[obase + n] = x;
    if (p.act == YV_ACT_GELU) x = yv_gelu(x);
    else if (p.act == YV_ACT_RELU) x = fmaxf(x, 0.f);
    if (drop.thresh) x *= yv_drop_mul(drop, (uint32_t)(((long long)z * p.M + row) * p.N + n));
    if (p.act == YV_ACT_MUL_GELU_GRAD) x *= yv_gelu_grad(p.aux_in[obase + n]);
    else if (p.act == YV_ACT_MUL_RELU_MASK) x = p.aux_in[obase + n] > 0.f ? x : 0.f;
    if (p.residual) x += p.residual[obase + n];
    if (p.out32) p.out32[obase + n] = x;
    if (p.out_planes) {
        __nv_bfloat16 h, l;
        yv_split(x, h, l);
        p.out_planes[pbase + n] = h;
        p.out_planes[pbase + n + p.pl_plane_stride] = l;
    }
}

// The epilogue's one global READ per element -- the residual, or aux_in for the activations that multiply by a saved
// tensor -- and the bias are fetched for a whole 32x32 chunk (8 float4 per lane) before the accumulator is even complete.  Loaded inside
// the row loop, each load sat behind the previous row's stores (they may alias), i.e. eight L2 round trips in series
// per chunk.
YV_DEVINL bool epilogue_pre_is_aux(int act) { return act == YV_ACT_MUL_GELU_GRAD || act == YV_ACT_MUL_RELU_MASK; }

YV_DEVINL void epilogue_prefetch(const KParams& p, int lane, int row0, int nc, long long obatch, int split, bool vec_ok,
                                 float4 (&pre)[8], float4& bias4) {
    const int n = nc + 4 * (lane & 7);
    const int r0 = lane >> 3;
    const float* src = epilogue_pre_is_aux(p.act) ? p.aux_in : (split != 0 ? nullptr : p.residual);
    bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(vec_ok && n + 3 < p.N)) src = nullptr;         // (the ragged-edge path loads per element)
    else if (p.bias && split == 0) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const int rows_left = p.M - (row0 + r0);
    const long long ob0 = obatch + (long long)(row0 + r0) * p.ld_out + n;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr && 4 * i < rows_left) pre[i] = *reinterpret_cast<const float4*>(src + ob0 + (long long)(4 * i) * p.ld_out);
    }
}

// Vector path of the epilogue for the 8 rows x 4 columns this lane owns in a staged 32x32 chunk.  ACT / DROP /
// SPLITK are compile-time so that a launch only issues the instructions of the features it uses: with run-time
// tests the compiler predicates the GELU / GELU' / dropout-hash code instead of branching around it, and the
// epilogue became issue-bound (~300 instructions per float4, ~10k cycles per 128x128 tile; see profiles/README.md).
// ACT == -1 selects the run-time generic version (rare feature combinations).
template <int ACT, bool DROP, bool SPLITK>
YV_DEVINL void epilogue_rows(const KParams& p, const YvDrop& drop, uint32_t stg, int lane, int row0, int n, int z,
                             long long obatch, long long pbatch, int split, float4 bias4, const float4 (&pre)[8]) {
    const int cg = lane & 7;
    const int act = ACT >= 0 ? ACT : p.act;
    const int r0 = lane >> 3;                            // this lane owns rows r0 + 4*i of the chunk
    // all eight staged float4 first: eight independent dependency chains for the scheduler to interleave
    // (the epilogue has only two warps per scheduler, so instruction-level parallelism is what hides latency)
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = r0 + 4 * i;
        const uint32_t addr = stg + (uint32_t)r * 128u + (uint32_t)((cg ^ (r & 7)) * 16);
        uint32_t x0, x1, x2, x3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
        v[i] = make_float4(__uint_as_float(x0), __uint_as_float(x1), __uint_as_float(x2), __uint_as_float(x3));
#if YV_DBG_EPI == 2
        v[i] = make_float4(1.f * r, 2.f, 3.f, 4.f);
#endif
    }
#if YV_DBG_EPI == 1
    {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
        if (acc == 12345.f) p.out32[0] = acc;
        return;
    }
#endif
    const long long ob0 = obatch + (long long)(row0 + r0) * p.ld_out + n;
    const long long pb0 = pbatch + (long long)(row0 + r0) * p.ld_pl + n;
    const long long ostep = 4 * p.ld_out, pstep = 4 * p.ld_pl;
    float* const out32 = p.out32;
    float* const aux_out = p.aux_out;
    // `pre` (epilogue_prefetch) holds aux_in when the activation reads it, else the residual; the other one (both in
    // one launch: not a combination the step uses) is loaded in the loop
    const bool pre_aux = epilogue_pre_is_aux(act);
    const float* const aux_in = p.aux_in;
    const float* const residual = (SPLITK && split != 0) ? nullptr : p.residual;
    __nv_bfloat16* const planes = p.out_planes;
    const long long plane_stride = p.pl_plane_stride;
    const float alpha = p.alpha;
    const int rows_left = p.M - (row0 + r0);             // row r0 + 4*i exists iff 4*i < rows_left
    const uint32_t drop_base = DROP ? (uint32_t)(((long long)z * p.M + row0 + r0) * p.N + n) : 0u;
    const uint32_t drop_step = DROP ? (uint32_t)(4 * p.N) : 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool ok = 4 * i < rows_left;
        const long long ob = ob0 + i * ostep;
        float4 x = v[i];
        x.x = alpha * x.x + bias4.x; x.y = alpha * x.y + bias4.y;
        x.z = alpha * x.z + bias4.z; x.w = alpha * x.w + bias4.w;
        if (!SPLITK) {
            if (aux_out && ok) *reinterpret_cast<float4*>(aux_out + ob) = x;
            if (act == YV_ACT_GELU) {
                x.x = yv_gelu(x.x); x.y = yv_gelu(x.y); x.z = yv_gelu(x.z); x.w = yv_gelu(x.w);
            } else if (act == YV_ACT_RELU) {
                x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
            }
        }
        if (DROP) {
            // (split-K: the epilogue is linear, dropout scales every partial sum)
            const uint32_t i0 = drop_base + (uint32_t)i * drop_step;
            x.x *= yv_drop_mul(drop, i0); x.y *= yv_drop_mul(drop, i0 + 1);
            x.z *= yv_drop_mul(drop, i0 + 2); x.w *= yv_drop_mul(drop, i0 + 3);
        }
        if (!SPLITK) {
            if (act == YV_ACT_MUL_GELU_GRAD) {
                const float4 t = pre[i];
                x.x *= yv_gelu_grad(t.x); x.y *= yv_gelu_grad(t.y); x.z *= yv_gelu_grad(t.z); x.w *= yv_gelu_grad(t.w);
            } else if (act == YV_ACT_MUL_RELU_MASK) {
                const float4 t = pre[i];
                x.x = t.x > 0.f ? x.x : 0.f; x.y = t.y > 0.f ? x.y : 0.f;
                x.z = t.z > 0.f ? x.z : 0.f; x.w = t.w > 0.f ? x.w : 0.f;
            }
        }
        if (!pre_aux) {                                  // (zero when there is no residual / the row does not exist)
            x.x += pre[i].x; x.y += pre[i].y; x.z += pre[i].z; x.w += pre[i].w;
        } else if (residual && ok) {
            const float4 t = *reinterpret_cast<const float4*>(residual + ob);
            x.x += t.x; x.y += t.y; x.z += t.z; x.w += t.w;
        }
        if (SPLITK) {
            // split-K: partial sums meet in a zero-initialised f32 output through vector reductions
            // (bias and residual come from split 0 only)
            if (ok)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out32 + ob), "f"(x.x), "f"(x.y),
                             "f"(x.z), "f"(x.w)
                             : "memory");
            continue;
        }
#ifdef YV_DBG_NOSTORE
        if (x.x != 12345.f) continue;
#endif
        if (out32 && ok) *reinterpret_cast<float4*>(out32 + ob) = x;
        if (planes) {
            __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
            yv_split(x.x, h0, l0); yv_split(x.y, h1, l1);
            yv_split(x.z, h2, l2); yv_split(x.w, h3, l3);
            uint2 hv, lv;
            hv.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            hv.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            lv.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            lv.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
            if (ok) {
                const long long pb = pb0 + i * pstep;
                *reinterpret_cast<uint2*>(planes + pb) = hv;
                *reinterpret_cast<uint2*>(planes + pb + plane_stride) = lv;
            }
        }
    }
}

// Fused epilogue of one 32-row x 32-column chunk of the accumulator.  `raw` holds the chunk as read from TMEM
// (thread = row); it goes through the warp's XOR-swizzled 4 KB staging buffer `stg` so that 8 lanes cover one
// 128-byte row segment and every global access is coalesced.  row0 = first row of the chunk, nc = first column.
YV_DEVINL void epilogue_chunk(const KParams& p, const YvDrop& drop, uint32_t stg, int lane, const uint32_t* raw, int row0,
                              int nc, int z, long long obatch, long long pbatch, int split, bool vec_ok,
                              const float4 (&pre)[8], const float4 bias4) {
    const int cg = lane & 7;                             // float4 column group of this lane inside the chunk
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) * 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(raw[4 * g]), "r"(raw[4 * g + 1]),
                     "r"(raw[4 * g + 2]), "r"(raw[4 * g + 3])
                     : "memory");
    }
    __syncwarp();
    YV_T64(9);
    const int n = nc + 4 * cg;
    if (n < p.N) {
        if (vec_ok && (n + 3 < p.N)) {
            const bool drop_on = drop.thresh != 0;
#define YV_EPI(A, D, S) epilogue_rows<A, D, S>(p, drop, stg, lane, row0, n, z, obatch, pbatch, split, bias4, pre)
            if (p.splits > 1) {
                if (drop_on) YV_EPI(YV_ACT_NONE, true, true); else YV_EPI(YV_ACT_NONE, false, true);
            } else if (p.act == YV_ACT_NONE) {
                if (drop_on) YV_EPI(YV_ACT_NONE, true, false); else YV_EPI(YV_ACT_NONE, false, false);
            } else if (drop_on) {
                YV_EPI(-1, true, false);
            } else if (p.act == YV_ACT_GELU) {
                YV_EPI(YV_ACT_GELU, false, false);
            } else if (p.act == YV_ACT_MUL_GELU_GRAD) {
                YV_EPI(YV_ACT_MUL_GELU_GRAD, false, false);
            } else if (p.act == YV_ACT_RELU) {
                YV_EPI(YV_ACT_RELU, false, false);
            } else {
                YV_EPI(YV_ACT_MUL_RELU_MASK, false, false);
            }
#undef YV_EPI
        } else {                                         // ragged right edge / unaligned leading dimension
            for (int i = 0; i < 8; ++i) {
                const int r = (lane >> 3) + 4 * i;
                const int row = row0 + r;
                if (row >= p.M) break;
                const uint32_t addr = stg + (uint32_t)r * 128u + (uint32_t)((cg ^ (r & 7)) * 16);
                uint32_t x0, x1, x2, x3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
                const long long ob = obatch + (long long)row * p.ld_out;
                const long long pb = pbatch + (long long)row * p.ld_pl;
                epilogue_scalar(p, drop, __uint_as_float(x0), n, z, row, ob, pb);
                epilogue_scalar(p, drop, __uint_as_float(x1), n + 1, z, row, ob, pb);
                epilogue_scalar(p, drop, __uint_as_float(x2), n + 2, z, row, ob, pb);
                epilogue_scalar(p, drop, __uint_as_float(x3), n + 3, z, row, ob, pb);
            }
        }
    }
    YV_T64(10);
    __syncwarp();                                    // staging buffer is reused by the next chunk
}

// ------------------------------------------------------------------------------------------- host
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
const bool g_split_k = []() { const char* e = getenv("YVB200_SPLIT_K"); return !(e && e[0] == '0'); }();

// Split-K plan shared by the variants: split when the tile count leaves most SMs idle and the epilogue is linear with
// a plain f32 output whose rows are 16-byte aligned (partial sums meet through vector reductions in a zero-filled
// output).  `ctas` = CTAs the un-split problem would launch.  Returns the number of splits (1 = no split).
inline int plan_split_k(const YvGemm* g, int ctas, int total_kb, int* kb_per_split) {
    *kb_per_split = total_kb;
    const bool linear_epi = g->act == YV_ACT_NONE && !g->aux_out && !g->out_planes && g->out32 &&
                            (g->ld_out % 4 == 0) && (g->N % 4 == 0) && (((uintptr_t)g->out32) & 15) == 0 &&
                            (!g->residual || g->residual != g->out32) && (!g->bias || (((uintptr_t)g->bias) & 15) == 0) &&
                            (!g->residual || (((uintptr_t)g->residual) & 15) == 0);
    if (g->a.nb0 * g->a.nb1 != 1 || !linear_epi || ctas * 2 > 148 || total_kb < 8 || !g_split_k) return 1;
    int s = 148 / ctas;
    if (s > total_kb / 4) s = total_kb / 4;
    if (s > 16) s = 16;
    if (s < 2) return 1;
    *kb_per_split = (total_kb + s - 1) / s;
    return (total_kb + *kb_per_split - 1) / *kb_per_split;
}

int get_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
        yv_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return 1;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return 0;
}

// K-major operands: box = block_k k-elements x box_rows rows (64B swizzle for block_k 32, 128B for 64);
// MN-major operands: box = 64 m/n-elements (128 B, 128B swizzle) x block_k k-rows
int make_map(CUtensorMap* map, const YvOperand& o, int passes, const char* which, int block_k, int box_rows) {
    YV_CHECK(o.ptr != nullptr, "yv_gemm: operand %s is NULL", which);
    YV_CHECK(((uintptr_t)o.ptr & 15) == 0, "yv_gemm: operand %s not 16-byte aligned", which);
    YV_CHECK(o.inner > 0 && o.rows > 0 && o.nb0 > 0 && o.nb1 > 0, "yv_gemm: operand %s has empty extent", which);
    YV_CHECK((o.ld & 7) == 0 && o.ld >= o.inner, "yv_gemm: operand %s ld=%lld must be a multiple of 8 and >= inner=%lld",
             which, (long long)o.ld, (long long)o.inner);
    YV_CHECK((o.nb0 == 1 || (o.sb0 & 7) == 0) && (o.nb1 == 1 || (o.sb1 & 7) == 0),
             "yv_gemm: operand %s batch strides must be multiples of 8 elements", which);
    YV_CHECK(passes == 1 || ((o.plane_stride & 7) == 0 && o.plane_stride > 0),
             "yv_gemm: operand %s plane_stride must be a positive multiple of 8", which);
    const int nplanes = passes == 3 ? 2 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)o.inner, (cuuint64_t)o.rows, (cuuint64_t)o.nb0, (cuuint64_t)o.nb1,
                          (cuuint64_t)nplanes};
    // strides of dims 1..4 in bytes (dim 0 is contiguous); unused dims get a harmless valid stride
    const cuuint64_t row_b = (cuuint64_t)o.ld * 2;
    cuuint64_t strides[4] = {row_b, o.nb0 > 1 ? (cuuint64_t)o.sb0 * 2 : row_b, o.nb1 > 1 ? (cuuint64_t)o.sb1 * 2 : row_b,
                             nplanes > 1 ? (cuuint64_t)o.plane_stride * 2 : row_b};
    cuuint32_t box[5] = {(cuuint32_t)(o.mn_major ? 64 : block_k), (cuuint32_t)(o.mn_major ? block_k : box_rows), 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(o.ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, (o.mn_major || block_k == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YV_CHECK(r == CUDA_SUCCESS, "yv_gemm: cuTensorMapEncodeTiled(%s) failed with %d (inner=%lld rows=%lld ld=%lld)", which,
             (int)r, (long long)o.inner, (long long)o.rows, (long long)o.ld);
    return 0;
}

}  // namespace
