// yv_gemm: batched, hi/lo-split (bf16x3) tcgen05 GEMM with a fused row-owning epilogue.  sm_100a only.
//
//   D[z] = epilogue(alpha * sum_p A_p[z] . B_p[z]^T)        p in {(hi,hi)} or {(hi,hi),(hi,lo),(lo,hi)}
//
// One CTA per 128x128 output tile, 192 threads:
//   warp 0   : TMA producer (one lane) -- cp.async.bulk.tensor.5d (swizzled) into a 3/6-stage, 96 KB ring
//              (128x128x32 tiles keep the CTA under half an SM's shared memory: two CTAs co-reside, so one
//               CTA's prologue / epilogue overlaps the other's MMA main loop)
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one lane), commits to mbarriers
//   warps 2-9: epilogue -- tcgen05.ld 32x32 chunks, transposed through swizzled smem for coalesced I/O
// Operands may be K-major or MN-major (transposed views): dgrad / wgrad / attention products need no
// explicit transposes.  Out-of-bounds rows/cols/k are zero-filled by TMA (each batch dim is its own
// tensor-map dim), so ragged shapes (T=80, vocab 30522, 1601 classes) need no padding copies.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 128;
// This file is compiled twice (see Makefile):
//   YV_BLOCK_K=64 -> yv_gemm_k64: persistent, one CTA per SM, 192 KB operand ring + dedicated epilogue staging.
//                    Best when a launch has several tiles per SM (epilogue of tile i overlaps main loop of i+1).
//   YV_BLOCK_K=32 -> yv_gemm_k32: one tile per CTA, 96 KB ring (epilogue staging reuses it), two CTAs per SM.
//                    Best for the single-wave problems of the 8-pair step: CTAs of concurrent launches share SMs.
#ifndef YV_BLOCK_K
#define YV_BLOCK_K 64
#endif
#if YV_BLOCK_K == 32
#define YV_GEMM_ENTRY yv_gemm_k32
#else
#define YV_GEMM_ENTRY yv_gemm_k64
#endif
constexpr int BLOCK_K = YV_BLOCK_K;                           // 32: 64 B rows / 64B swizzle / 2 CTAs per SM; 64: 128B swizzle
constexpr int KMAJ_LAYOUT = BLOCK_K == 32 ? 4 : 2;            // UMMA layout type of K-major tiles (SWIZZLE_64B / _128B)
constexpr int KMAJ_SBO = BLOCK_K * 2 * 8;                     // 8 rows of BLOCK_K bf16
constexpr int UMMA_K = 16;
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 2;             // 8 KB (A and B tiles have the same size)
constexpr int NUM_THREADS = 320;                               // TMA warp, MMA warp, 8 epilogue warps
constexpr int NUM_EPI_WARPS = 8;
constexpr bool PERSISTENT = BLOCK_K == 64;
constexpr int TMEM_COLS = PERSISTENT ? 256 : 128;             // two 128-column f32 accumulators when persistent
constexpr int EPI_STAGING_BYTES = PERSISTENT ? NUM_EPI_WARPS * 4096 : 0;   // k32: staging reuses the drained ring

template <int PASSES>
struct Cfg {
    static constexpr int TILES_PER_STAGE = PASSES == 3 ? 4 : 2;   // A_hi, B_hi, (A_lo, B_lo)
    static constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;
    static constexpr int STAGES = PASSES == 3 ? 3 : 6;            // 96 KB (BLOCK_K=32) / 192 KB (64) operand ring
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
    static_assert(PERSISTENT || STAGES * STAGE_BYTES >= NUM_EPI_WARPS * 4096, "k32 staging must fit in the ring");
};

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

#ifdef YV_GEMM_TIMING
__device__ long long yv_dbg[32];
#define YV_T(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) yv_dbg[i] = clock64(); } while (0)
#else
#define YV_T(i)
#endif

// ------------------------------------------------------------------------------------------- kernel
// Persistent: one CTA per SM walks the tile list (tile = blockIdx.x + i * gridDim.x).  Two TMEM accumulators
// (2 x 128 columns) let the epilogue of tile i overlap the TMA/MMA main loop of tile i+1.
template <int PASSES>
__global__ void __launch_bounds__(NUM_THREADS, PERSISTENT ? 1 : 2)
yv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const KParams p) {
    using C = Cfg<PASSES>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    // epilogue staging, 8 x 4 KB: after the ring (persistent) or on top of it (one tile per CTA: ring is drained)
    const uint32_t stage_smem = PERSISTENT ? smem_base + C::STAGES * C::STAGE_BYTES : smem_base;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + C::STAGES * C::STAGE_BYTES + EPI_STAGING_BYTES);
    // bars: [0,S) full, [S,2S) empty, 2S..2S+1 tmem_full, 2S+2..2S+3 tmem_empty; then the TMEM base address word
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
    if (threadIdx.x == 0) YV_T(0);
    yv_pdl_trigger();      // the next kernel may start its own prologue while this one runs

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), NUM_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_b) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    yv_pdl_wait();         // barriers, TMEM and descriptors are ready: now wait for the producers of our operands
    if (threadIdx.x == 0) YV_T(1);

    // tile -> (split, n block, m block, batch); m varies fastest so concurrently running CTAs share B tiles in L2
    auto decode = [&](int t, int& m0, int& n0, int& z, int& split) {
        split = t % p.splits;
        t /= p.splits;
        m0 = (t % tiles_m) * BLOCK_M;
        t /= tiles_m;
        n0 = (t % tiles_n) * BLOCK_N;
        z = t / tiles_n;
    };

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int m0, n0, z, split;
                decode(tile, m0, n0, z, split);
                const int b0 = z % p.nb0, b1 = z / p.nb0;
                const int kb_lo = split * p.kb_per_split;
                const int num_kb = min(p.kb_per_split, total_kb - kb_lo);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sbase = smem_base + stage * C::STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                    const int k0 = (kb_lo + kb) * BLOCK_K;
#pragma unroll
                    for (int pl = 0; pl < (PASSES == 3 ? 2 : 1); ++pl) {
                        const uint32_t sa = sbase + (pl * 2 + 0) * TILE_BYTES;
                        const uint32_t sb = sbase + (pl * 2 + 1) * TILE_BYTES;
                        if (!p.a_mn) {
                            tma_load_5d(sa, &map_a, full_bar(stage), k0, m0, b0, b1, pl);
                        } else {
                            tma_load_5d(sa, &map_a, full_bar(stage), m0, k0, b0, b1, pl);
                            tma_load_5d(sa + TILE_BYTES / 2, &map_a, full_bar(stage), m0 + 64, k0, b0, b1, pl);
                        }
                        if (!p.b_mn) {
                            tma_load_5d(sb, &map_b, full_bar(stage), k0, n0, b0, b1, pl);
                        } else {
                            tma_load_5d(sb, &map_b, full_bar(stage), n0, k0, b0, b1, pl);
                            tma_load_5d(sb + TILE_BYTES / 2, &map_b, full_bar(stage), n0 + 64, k0, b0, b1, pl);
                        }
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ========================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                                   ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            // per-UMMA_K advance of the descriptor start address (in 16-byte units)
            const uint32_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            const uint32_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int split = tile % p.splits;
                const int kb_lo = split * p.kb_per_split;
                const int num_kb = min(p.kb_per_split, total_kb - kb_lo);
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                uint32_t accum = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    if (kb == 0) YV_T(2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sbase = smem_base + stage * C::STAGE_BYTES;
                    uint64_t da[2], db[2];
#pragma unroll
                    for (int pl = 0; pl < (PASSES == 3 ? 2 : 1); ++pl) {
                        const uint32_t sa = sbase + (pl * 2 + 0) * TILE_BYTES;
                        const uint32_t sb = sbase + (pl * 2 + 1) * TILE_BYTES;
                        da[pl] = p.a_mn ? make_desc(sa, TILE_BYTES / 2, 1024, 2) : make_desc(sa, 16, KMAJ_SBO, KMAJ_LAYOUT);
                        db[pl] = p.b_mn ? make_desc(sb, TILE_BYTES / 2, 1024, 2) : make_desc(sb, 16, KMAJ_SBO, KMAJ_LAYOUT);
                    }
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        if (PASSES == 3) {
                            // small cross terms first, the dominant hi*hi term last
                            umma_bf16(tmem_d, da[1] + (uint64_t)(a_step * k), db[0] + (uint64_t)(b_step * k), idesc, accum);
                            accum = 1;
                            umma_bf16(tmem_d, da[0] + (uint64_t)(a_step * k), db[1] + (uint64_t)(b_step * k), idesc, 1);
                        }
                        umma_bf16(tmem_d, da[0] + (uint64_t)(a_step * k), db[0] + (uint64_t)(b_step * k), idesc, accum);
                        accum = 1;
                    }
                    umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs retire
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                YV_T(3);
                umma_commit(tmem_full_bar(acc));     // accumulator complete -> epilogue
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===================================== epilogue ==========================================
        // 8 warps: two per TMEM lane quarter, each owning 64 of the tile's 128 columns as two 32x32 chunks.
        // A chunk goes TMEM -> registers (thread = row) -> XOR-swizzled smem staging -> registers (8 lanes = one
        // 128-byte row segment), so every global access below is coalesced.
        const int ew = warp - 2;
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = ew >> 2;
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t stg = stage_smem + (uint32_t)ew * 4096u;
        const bool vec_ok = ((p.ld_out & 3) == 0) && ((p.out_sb0 & 3) == 0) && ((p.out_sb1 & 3) == 0) &&
                            ((p.ld_pl & 3) == 0) && ((p.pl_sb0 & 3) == 0) && ((p.pl_sb1 & 3) == 0) &&
                            ((p.pl_plane_stride & 3) == 0) &&
                            (((uintptr_t)p.out32 | (uintptr_t)p.aux_out | (uintptr_t)p.aux_in | (uintptr_t)p.residual |
                              (uintptr_t)p.bias) & 15) == 0 && (((uintptr_t)p.out_planes) & 7) == 0;
        const int cg = lane & 7;                             // float4 column group of this lane inside a chunk
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int m0, n0, z, split;
            decode(tile, m0, n0, z, split);
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const long long obatch = (long long)b0 * p.out_sb0 + (long long)b1 * p.out_sb1;
            const long long pbatch = (long long)b0 * p.pl_sb0 + (long long)b1 * p.pl_sb1;
            mbar_wait(tmem_full_bar(acc), acc_phase);
            if (threadIdx.x == 64) YV_T(4);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = half * 2 + cc;
                const int nc = n0 + c * 32;
                if (nc >= p.N) break;                            // warp-uniform
                uint32_t raw[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * 32), raw);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const uint32_t addr = stg + (uint32_t)lane * 128u + (uint32_t)((g ^ (lane & 7)) * 16);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(raw[4 * g]), "r"(raw[4 * g + 1]),
                                 "r"(raw[4 * g + 2]), "r"(raw[4 * g + 3])
                                 : "memory");
                }
                __syncwarp();
                const int n = nc + 4 * cg;
                const bool quad_ok = vec_ok && (n + 3 < p.N);
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias && quad_ok && split == 0) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll 2
                for (int i = 0; i < 8; ++i) {
                    const int r = (lane >> 3) + 4 * i;
                    const int row = m0 + q * 32 + r;
                    float4 v;
                    {
                        const uint32_t addr = stg + (uint32_t)r * 128u + (uint32_t)((cg ^ (r & 7)) * 16);
                        uint32_t x0, x1, x2, x3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
                        v = make_float4(__uint_as_float(x0), __uint_as_float(x1), __uint_as_float(x2), __uint_as_float(x3));
                    }
                    if (row >= p.M || n >= p.N) continue;
                    const long long ob = obatch + (long long)row * p.ld_out + n;
                    const long long pb = pbatch + (long long)row * p.ld_pl + n;
                    if (!quad_ok) {                              // ragged edge / unaligned leading dimension
                        epilogue_scalar(p, drop, v.x, n, z, row, ob - n, pb - n);
                        epilogue_scalar(p, drop, v.y, n + 1, z, row, ob - n, pb - n);
                        epilogue_scalar(p, drop, v.z, n + 2, z, row, ob - n, pb - n);
                        epilogue_scalar(p, drop, v.w, n + 3, z, row, ob - n, pb - n);
                        continue;
                    }
                    v.x = p.alpha * v.x + bias4.x; v.y = p.alpha * v.y + bias4.y;
                    v.z = p.alpha * v.z + bias4.z; v.w = p.alpha * v.w + bias4.w;
                    if (p.splits > 1) {
                        // split-K: partial sums meet in a zero-initialised f32 output through vector reductions; the
                        // epilogue is linear here (bias and residual come from split 0, dropout scales every partial)
                        if (drop.thresh) {
                            const uint32_t i0 = (uint32_t)(((long long)z * p.M + row) * p.N + n);
                            v.x *= yv_drop_mul(drop, i0); v.y *= yv_drop_mul(drop, i0 + 1);
                            v.z *= yv_drop_mul(drop, i0 + 2); v.w *= yv_drop_mul(drop, i0 + 3);
                        }
                        if (p.residual && split == 0) {
                            const float4 t = *reinterpret_cast<const float4*>(p.residual + ob);
                            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                        }
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out32 + ob), "f"(v.x), "f"(v.y),
                                     "f"(v.z), "f"(v.w)
                                     : "memory");
                        continue;
                    }
                    if (p.aux_out) *reinterpret_cast<float4*>(p.aux_out + ob) = v;
                    if (p.act == YV_ACT_GELU) {
                        v.x = yv_gelu(v.x); v.y = yv_gelu(v.y); v.z = yv_gelu(v.z); v.w = yv_gelu(v.w);
                    } else if (p.act == YV_ACT_RELU) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                    if (drop.thresh) {
                        const uint32_t i0 = (uint32_t)(((long long)z * p.M + row) * p.N + n);
                        v.x *= yv_drop_mul(drop, i0); v.y *= yv_drop_mul(drop, i0 + 1);
                        v.z *= yv_drop_mul(drop, i0 + 2); v.w *= yv_drop_mul(drop, i0 + 3);
                    }
                    if (p.act == YV_ACT_MUL_GELU_GRAD) {
                        const float4 t = *reinterpret_cast<const float4*>(p.aux_in + ob);
                        v.x *= yv_gelu_grad(t.x); v.y *= yv_gelu_grad(t.y); v.z *= yv_gelu_grad(t.z); v.w *= yv_gelu_grad(t.w);
                    } else if (p.act == YV_ACT_MUL_RELU_MASK) {
                        const float4 t = *reinterpret_cast<const float4*>(p.aux_in + ob);
                        v.x = t.x > 0.f ? v.x : 0.f; v.y = t.y > 0.f ? v.y : 0.f;
                        v.z = t.z > 0.f ? v.z : 0.f; v.w = t.w > 0.f ? v.w : 0.f;
                    }
                    if (p.residual) {
                        const float4 t = *reinterpret_cast<const float4*>(p.residual + ob);
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                    }
                    if (p.out32) *reinterpret_cast<float4*>(p.out32 + ob) = v;
                    if (p.out_planes) {
                        __align__(8) __nv_bfloat16 h4[4], l4[4];
                        yv_split(v.x, h4[0], l4[0]); yv_split(v.y, h4[1], l4[1]);
                        yv_split(v.z, h4[2], l4[2]); yv_split(v.w, h4[3], l4[3]);
                        *reinterpret_cast<uint2*>(p.out_planes + pb) = *reinterpret_cast<uint2*>(h4);
                        *reinterpret_cast<uint2*>(p.out_planes + pb + p.pl_plane_stride) = *reinterpret_cast<uint2*>(l4);
                    }
                }
                __syncwarp();                                    // staging buffer is reused by the next chunk
            }
            // this warp has finished reading its TMEM lanes of accumulator `acc`: hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(acc)) : "memory");
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    if (threadIdx.x == 64) YV_T(5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) YV_T(6);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------- host
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
const bool g_split_k = []() { const char* e = getenv("YVB200_SPLIT_K"); return !(e && e[0] == '0'); }();

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

int make_map(CUtensorMap* map, const YvOperand& o, int passes, const char* which) {
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
    // K-major: 32 k-elements (64 B, 64B swizzle) x 128 rows; MN-major: 64 m/n-elements (128 B, 128B swizzle) x 32 k-rows
    cuuint32_t box[5] = {(cuuint32_t)(o.mn_major ? 64 : BLOCK_K), (cuuint32_t)(o.mn_major ? BLOCK_K : BLOCK_M), 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(o.ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, (o.mn_major || BLOCK_K == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    YV_CHECK(r == CUDA_SUCCESS, "yv_gemm: cuTensorMapEncodeTiled(%s) failed with %d (inner=%lld rows=%lld ld=%lld)", which,
             (int)r, (long long)o.inner, (long long)o.rows, (long long)o.ld);
    return 0;
}

}  // namespace

void yv_count_launch();

extern "C" int YV_GEMM_ENTRY(const YvGemm* g, yv_stream_t stream) {
    YV_CHECK(g != nullptr, "yv_gemm: NULL args");
    YV_CHECK(g->passes == 1 || g->passes == 3, "yv_gemm: passes must be 1 or 3 (got %d)", g->passes);
    YV_CHECK(g->M > 0 && g->N > 0 && g->K > 0, "yv_gemm: empty problem M=%d N=%d K=%d", g->M, g->N, g->K);
    YV_CHECK(g->out32 || g->out_planes, "yv_gemm: no output requested");
    if (get_encode()) return 1;
    // operand extents must agree with the problem
    const YvOperand &a = g->a, &b = g->b;
    YV_CHECK((a.mn_major ? a.inner : a.rows) == g->M && (a.mn_major ? a.rows : a.inner) == g->K,
             "yv_gemm: A extents (%lld x %lld, mn_major=%d) do not match M=%d K=%d", (long long)a.rows, (long long)a.inner,
             a.mn_major, g->M, g->K);
    YV_CHECK((b.mn_major ? b.inner : b.rows) == g->N && (b.mn_major ? b.rows : b.inner) == g->K,
             "yv_gemm: B extents (%lld x %lld, mn_major=%d) do not match N=%d K=%d", (long long)b.rows, (long long)b.inner,
             b.mn_major, g->N, g->K);
    YV_CHECK(a.nb0 == b.nb0 && a.nb1 == b.nb1, "yv_gemm: batch counts differ");
    CUtensorMap ma, mb;
    if (make_map(&ma, a, g->passes, "A")) return 1;
    if (make_map(&mb, b, g->passes, "B")) return 1;

    KParams p;
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.nb0 = (int)a.nb0;
    p.a_mn = a.mn_major ? 1 : 0;
    p.b_mn = b.mn_major ? 1 : 0;
    p.alpha = g->alpha;
    p.act = g->act;
    p.bias = g->bias;
    p.aux_out = g->aux_out;
    p.aux_in = g->aux_in;
    p.residual = g->residual;
    p.out32 = g->out32;
    p.ld_out = g->ld_out; p.out_sb0 = g->out_sb0; p.out_sb1 = g->out_sb1;
    p.out_planes = reinterpret_cast<__nv_bfloat16*>(g->out_planes);
    p.ld_pl = g->ld_pl; p.pl_sb0 = g->pl_sb0; p.pl_sb1 = g->pl_sb1; p.pl_plane_stride = g->pl_plane_stride;
    p.drop_p = g->drop_p; p.drop_site = g->drop_site;
    p.rng = reinterpret_cast<const unsigned long long*>(g->rng);
    YV_CHECK((g->act != YV_ACT_MUL_GELU_GRAD && g->act != YV_ACT_MUL_RELU_MASK) || g->aux_in,
             "yv_gemm: act %d needs aux_in", g->act);

    const int tiles = ((g->N + BLOCK_N - 1) / BLOCK_N) * ((g->M + BLOCK_M - 1) / BLOCK_M);
    const int total_kb = (g->K + BLOCK_K - 1) / BLOCK_K;
    p.splits = 1;
    p.kb_per_split = total_kb;
    // split-K when the tile count leaves most SMs idle and the epilogue is linear with a plain f32 output whose
    // rows are 16-byte aligned (the output must then be zero-filled: done here with a memset node)
    const bool linear_epi = g->act == YV_ACT_NONE && !g->aux_out && !g->out_planes && g->out32 &&
                            (g->ld_out % 4 == 0) && (g->N % 4 == 0) && (((uintptr_t)g->out32) & 15) == 0 &&
                            (!g->residual || g->residual != g->out32) && (!g->bias || (((uintptr_t)g->bias) & 15) == 0) &&
                            (!g->residual || (((uintptr_t)g->residual) & 15) == 0);
    if (a.nb0 * a.nb1 == 1 && linear_epi && tiles * 2 <= 148 && total_kb >= 8 && g_split_k) {
        int s = 148 / tiles;
        if (s > total_kb / 4) s = total_kb / 4;
        if (s > 16) s = 16;
        if (s >= 2) {
            p.kb_per_split = (total_kb + s - 1) / s;
            p.splits = (total_kb + p.kb_per_split - 1) / p.kb_per_split;
        }
    }
    const long long total_tiles = (long long)tiles * a.nb0 * a.nb1 * p.splits;
    YV_CHECK(total_tiles < 2147483647LL, "yv_gemm: too many tiles");
    p.total_tiles = (int)total_tiles;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        YV_CUDA(cudaGetDevice(&dev));
        YV_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    dim3 grid((unsigned)((!PERSISTENT || total_tiles < num_sms) ? total_tiles : num_sms), 1, 1);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p.splits > 1)
        YV_CUDA(cudaMemset2DAsync(g->out32, sizeof(float) * g->ld_out, 0, sizeof(float) * g->N, g->M, st));
    static bool attr_set = false;
    if (!attr_set) {
        YV_CUDA(cudaFuncSetAttribute(yv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
        YV_CUDA(cudaFuncSetAttribute(yv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<3>::SMEM_BYTES));
        attr_set = true;
    }
    if (g->passes == 3)
        YV_CUDA(yv_launch(yv_gemm_kernel<3>, grid, dim3(NUM_THREADS), Cfg<3>::SMEM_BYTES, st, ma, mb, p));
    else
        YV_CUDA(yv_launch(yv_gemm_kernel<1>, grid, dim3(NUM_THREADS), Cfg<1>::SMEM_BYTES, st, ma, mb, p));
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}
