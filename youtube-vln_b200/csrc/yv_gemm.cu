// yv_gemm: batched, hi/lo-split (bf16x3) tcgen05 GEMM with a fused row-owning epilogue.  sm_100a only.
//
//   D[z] = epilogue(alpha * sum_p A_p[z] . B_p[z]^T)        p in {(hi,hi)} or {(hi,hi),(hi,lo),(lo,hi)}
//
// One CTA per 128x128 output tile, 320 threads:
//   warp 0   : TMA producer (one lane) -- cp.async.bulk.tensor.5d (swizzled) into a 3-stage (bf16x3) / 6-stage (bf16)
//              mbarrier ring: 192 KB in the persistent k64 build (default), 96 KB in the k32 build (two CTAs per SM)
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one lane), commits to mbarriers; two 128-column accumulators in the
//              persistent build so that the epilogue of tile i overlaps the main loop of tile i+1
//   warps 2-9: epilogue -- tcgen05.ld 32x32 chunks, transposed through swizzled smem for coalesced I/O
//              (yv_gemm_common.cuh: compile-time specialised per activation / dropout / split-K)
// Operands may be K-major or MN-major (transposed views): dgrad / wgrad / attention products need no
// explicit transposes.  Out-of-bounds rows/cols/k are zero-filled by TMA (each batch dim is its own
// tensor-map dim), so ragged shapes (T=80, vocab 30522, 1601 classes) need no padding copies.
#include "yv_gemm_common.cuh"

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


// ------------------------------------------------------------------------------------------- kernel
// Persistent: one CTA per SM walks the tile list (tile = blockIdx.x + i * gridDim.x).  Two TMEM accumulators
// (2 x 128 columns) let the epilogue of tile i overlap the TMA/MMA main loop of tile i+1.
template <int PASSES>
__global__ void __launch_bounds__(NUM_THREADS, PERSISTENT ? 1 : 2)
yv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ KParams p) {
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
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            int m0, n0, z, split;
            decode(tile, m0, n0, z, split);
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const long long obatch = (long long)b0 * p.out_sb0 + (long long)b1 * p.out_sb1;
            const long long pbatch = (long long)b0 * p.pl_sb0 + (long long)b1 * p.pl_sb1;
            // residual / aux_in of the first chunk: in flight while the main loop is still running
            float4 pre[8];
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + half * 64 < p.N) epilogue_prefetch(p, lane, m0 + q * 32, n0 + half * 64, obatch, split, vec_ok, pre, bias4);
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
                YV_T64(8);
                epilogue_chunk(p, drop, stg, lane, raw, m0 + q * 32, nc, z, obatch, pbatch, split, vec_ok, pre, bias4);
                if (cc == 0 && nc + 32 < p.N) epilogue_prefetch(p, lane, m0 + q * 32, nc + 32, obatch, split, vec_ok, pre, bias4);
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

struct DevCache {
    int num_sms = 0;
    bool attr_set = false;
};
DevCache g_dev_cache[64];
std::mutex g_dev_mutex;

}  // namespace

void yv_count_launch();

#if YV_BLOCK_K == 32
#define YV_GEMM_SPLITS yv_gemm_k32_splits
#else
#define YV_GEMM_SPLITS yv_gemm_k64_splits
#endif
// number of K splits YV_GEMM_ENTRY would use for this problem (the caller may pre-zero the output, see out32_zeroed)
extern "C" int YV_GEMM_SPLITS(const YvGemm* g) {
    const int tiles = ((g->N + BLOCK_N - 1) / BLOCK_N) * ((g->M + BLOCK_M - 1) / BLOCK_M);
    int kbps;
    return plan_split_k(g, tiles, (g->K + BLOCK_K - 1) / BLOCK_K, &kbps);
}

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
    if (make_map(&ma, a, g->passes, "A", BLOCK_K, BLOCK_M)) return 1;
    if (make_map(&mb, b, g->passes, "B", BLOCK_K, BLOCK_N)) return 1;

    KParams p;
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.nb0 = (int)a.nb0;
    p.pair_n = 0; p.stages = 0;
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
    p.splits = plan_split_k(g, tiles, total_kb, &p.kb_per_split);
    const long long total_tiles = (long long)tiles * a.nb0 * a.nb1 * p.splits;
    YV_CHECK(total_tiles < 2147483647LL, "yv_gemm: too many tiles");
    p.total_tiles = (int)total_tiles;
    // per-device cache behind a mutex: the reference's nn.DataParallel fallback (utils/distributed.py:100-102) calls
    // forward from one thread per device, and cudaFuncSetAttribute is a per-device setting
    int num_sms = 0;
    {
        int dev = 0;
        YV_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        DevCache& dc = g_dev_cache[dev & 63];
        if (dc.num_sms == 0) YV_CUDA(cudaDeviceGetAttribute(&dc.num_sms, cudaDevAttrMultiProcessorCount, dev));
        if (!dc.attr_set) {
            YV_CUDA(cudaFuncSetAttribute(yv_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
            YV_CUDA(cudaFuncSetAttribute(yv_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<3>::SMEM_BYTES));
            dc.attr_set = true;
        }
        num_sms = dc.num_sms;
    }
    dim3 grid((unsigned)((!PERSISTENT || total_tiles < num_sms) ? total_tiles : num_sms), 1, 1);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p.splits > 1 && !g->out32_zeroed)
        YV_CUDA(cudaMemset2DAsync(g->out32, sizeof(float) * g->ld_out, 0, sizeof(float) * g->N, g->M, st));
    if (g->passes == 3)
        YV_CUDA(yv_launch(yv_gemm_kernel<3>, grid, dim3(NUM_THREADS), Cfg<3>::SMEM_BYTES, st, ma, mb, p));
    else
        YV_CUDA(yv_launch(yv_gemm_kernel<1>, grid, dim3(NUM_THREADS), Cfg<1>::SMEM_BYTES, st, ma, mb, p));
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}
