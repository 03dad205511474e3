// yv_gemm_pair: the CTA-pair (tcgen05 cta_group::2) variant of yv_gemm.  sm_100a only.
//
// Why: with one CTA per 128x128 tile every SM pulls a full A tile and a full B tile (hi and lo planes) through the
// L2 -> SM crossbar for 3 x 128x128x32 MMAs; ncu shows the 2304x3072x1024 projection moving 453 MB at ~5500 B/clk
// (the chip-wide L2 -> SM limit) with the tensor pipe 53 % busy.  Here two CTAs on the two SMs of a TPC share one
// 256 x PAIR_N tile: each CTA stages its own 128 rows of A and only HALF of the B tile, and the leader's
// tcgen05.mma.cta_group::2 (UMMA_M = 256) reads both halves from both shared memories.  PAIR_N = 128 keeps the CTA
// count of the 128x128 tiling with 25 % fewer operand bytes per FLOP, PAIR_N = 256 halves the bytes per FLOP.
//
// One cluster of 2 CTAs per pair tile, 320 threads per CTA:
//   warp 0   : TMA producer (one lane, both CTAs) -- own A rows + own half of B into a ring of 48 KB (PAIR_N 128)
//              or 64 KB (256) stages filling shared memory; every load signals the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1   : TMEM allocation (cta_group::2, both CTAs); the leader's lane 0 issues every MMA and commits with
//              multicast arrives to the empty / accumulator-full barriers of both CTAs
//   warps 2-9: epilogue of this CTA's 128 rows x PAIR_N columns (same fused epilogue as yv_gemm.cu)
#include "yv_gemm_common.cuh"

namespace {

#ifndef YV_PAIR_BLOCK_K
#define YV_PAIR_BLOCK_K 64
#endif
constexpr int BLOCK_M = 128;                                  // rows per CTA; the pair tile covers 256
constexpr int BLOCK_K = YV_PAIR_BLOCK_K;                      // 64: 128 B rows / 128B swizzle (TMA moves ~1 row per 1.5 clk,
                                                              // so wide rows matter); 32: 64 B rows / 64B swizzle
constexpr int KMAJ_LAYOUT = BLOCK_K == 32 ? 4 : 2;            // UMMA layout type of K-major tiles (SWIZZLE_64B / _128B)
constexpr int KMAJ_SBO = BLOCK_K * 2 * 8;                     // 8 rows of BLOCK_K bf16
constexpr int MN_CHUNK = BLOCK_K * 128;                       // MN-major: 64 m/n-elements (128 B) x BLOCK_K k-rows
constexpr int UMMA_K = 16;
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 2;             // one plane of 128 rows x BLOCK_K k
constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_LIMIT = 232448;                            // 227 KB of dynamic shared memory per CTA

template <int PASSES>
struct PCfg {
    static constexpr int PLANES = PASSES == 3 ? 2 : 1;
    // one stage = per plane: this CTA's 128 rows of A, then its half (pair_n / 2 rows) of B
    static constexpr int stage_bytes(int pair_n) { return PLANES * (TILE_BYTES + (pair_n / 2) * BLOCK_K * 2); }
    static constexpr int smem_bytes(int stages, int pair_n) {
        return stages * stage_bytes(pair_n) + 1024 /*align slack*/ + 256 /*barriers*/;
    }
};

YV_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
YV_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) inside the CTA of rank `rank`
YV_DEVINL uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
YV_DEVINL void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5, %6, %7}], [%2];" ::"r"(dst),
        "l"((unsigned long long)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
YV_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
YV_DEVINL void umma2_commit(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((unsigned short)3)
        : "memory");
}

template <int PASSES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, BLOCK_K == 32 ? 2 : 1)
yv_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ KParams p) {
    using C = PCfg<PASSES>;
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: the MMA descriptors and the multicast commits address the peer by offset
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const int pair_n = p.pair_n;
    const int nb_half = pair_n >> 1;                             // B rows staged by each CTA
    const uint32_t a_bytes = TILE_BYTES, b_bytes = (uint32_t)(nb_half * BLOCK_K * 2);
    const uint32_t plane_bytes = a_bytes + b_bytes, stage_bytes = C::PLANES * plane_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + p.stages * stage_bytes);
    // bars: [0,8) full (used in the leader only), [8,16) empty, 16 accumulator full; then the TMEM base address word
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * MAX_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
    const int tiles_n = (p.N + pair_n - 1) / pair_n;
    const int tiles_mp = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    if (threadIdx.x == 0) YV_T(0);
    yv_pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_b) : "memory");
    }
    if (warp == 1) {     // the same warp of both CTAs allocates the pair's accumulator columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)pair_n)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();      // barriers of both CTAs are initialised before any remote arrive / complete_tx
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    yv_pdl_wait();
    if (threadIdx.x == 0) YV_T(1);

    // pair tile -> (split, n block, m pair, batch); m varies fastest so concurrently running pairs share B in L2
    int t = (int)(blockIdx.x >> 1);
    const int split = t % p.splits;
    t /= p.splits;
    const int m0 = (t % tiles_mp) * (2 * BLOCK_M) + (int)rank * BLOCK_M;   // first row of THIS CTA
    t /= tiles_mp;
    const int n0 = (t % tiles_n) * pair_n;
    const int z = t / tiles_n;
    const int b0 = z % p.nb0, b1 = z / p.nb0;
    const int kb_lo = split * p.kb_per_split;
    const int num_kb = min(p.kb_per_split, total_kb - kb_lo);

    if (warp == 0) {
        // ===================================== TMA producer (both CTAs) ==========================
        if (lane == 0) {
            const uint32_t tx_bytes = stage_bytes;
            const int nb0_row = n0 + (int)rank * nb_half;                  // first B row staged by this CTA
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sbase = smem_base + stage * stage_bytes;
                const uint32_t lbar = mapa_rank(full_bar(stage), 0);       // the leader's barrier collects both CTAs' bytes
                if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * tx_bytes);
                const int k0 = (kb_lo + kb) * BLOCK_K;
#pragma unroll
                for (int pl = 0; pl < C::PLANES; ++pl) {
                    const uint32_t sa = sbase + pl * plane_bytes;
                    const uint32_t sb = sa + a_bytes;
                    if (!p.a_mn) {
                        tma_load_5d_pair(sa, &map_a, lbar, k0, m0, b0, b1, pl);
                    } else {
                        tma_load_5d_pair(sa, &map_a, lbar, m0, k0, b0, b1, pl);
                        tma_load_5d_pair(sa + MN_CHUNK, &map_a, lbar, m0 + 64, k0, b0, b1, pl);
                    }
                    if (!p.b_mn) {
                        tma_load_5d_pair(sb, &map_b, lbar, k0, nb0_row, b0, b1, pl);
                    } else {
                        tma_load_5d_pair(sb, &map_b, lbar, nb0_row, k0, b0, b1, pl);
                        if (nb_half > 64) tma_load_5d_pair(sb + MN_CHUNK, &map_b, lbar, nb0_row + 64, k0, b0, b1, pl);
                    }
                }
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer (leader CTA only) ======================
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                                   ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(pair_n >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
            const uint32_t a_step = p.a_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            const uint32_t b_step = p.b_mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accum = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                if (kb == 0) YV_T(2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sbase = smem_base + stage * stage_bytes;
                uint64_t da[2], db[2];
#pragma unroll
                for (int pl = 0; pl < C::PLANES; ++pl) {
                    const uint32_t sa = sbase + pl * plane_bytes;
                    const uint32_t sb = sa + a_bytes;
                    da[pl] = p.a_mn ? make_desc(sa, MN_CHUNK, 1024, 2) : make_desc(sa, 16, KMAJ_SBO, KMAJ_LAYOUT);
                    db[pl] = p.b_mn ? make_desc(sb, MN_CHUNK, 1024, 2) : make_desc(sb, 16, KMAJ_SBO, KMAJ_LAYOUT);
                }
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    if (PASSES == 3) {
                        umma2_bf16(tmem_base, da[1] + (uint64_t)(a_step * k), db[0] + (uint64_t)(b_step * k), idesc, accum);
                        accum = 1;
                        umma2_bf16(tmem_base, da[0] + (uint64_t)(a_step * k), db[1] + (uint64_t)(b_step * k), idesc, 1);
                    }
                    umma2_bf16(tmem_base, da[0] + (uint64_t)(a_step * k), db[0] + (uint64_t)(b_step * k), idesc, accum);
                    accum = 1;
                }
                umma2_commit(empty_bar(stage));      // frees this ring slot in both CTAs once the MMAs retire
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            YV_T(3);
            umma2_commit(tmem_full_bar);             // accumulators of both CTAs complete -> both epilogues
        }
    } else {
        // ===================================== epilogue (both CTAs, own 128 rows) ================
        const int ew = warp - 2;
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = ew >> 2;
        const int cpw = pair_n >> 6;                         // 32-column chunks per warp: 2 (PAIR_N 128) or 4 (256)
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t stg = smem_base + (uint32_t)ew * 4096u;   // staging reuses the drained operand ring
        const bool vec_ok = ((p.ld_out & 3) == 0) && ((p.out_sb0 & 3) == 0) && ((p.out_sb1 & 3) == 0) &&
                            ((p.ld_pl & 3) == 0) && ((p.pl_sb0 & 3) == 0) && ((p.pl_sb1 & 3) == 0) &&
                            ((p.pl_plane_stride & 3) == 0) &&
                            (((uintptr_t)p.out32 | (uintptr_t)p.aux_out | (uintptr_t)p.aux_in | (uintptr_t)p.residual |
                              (uintptr_t)p.bias) & 15) == 0 && (((uintptr_t)p.out_planes) & 7) == 0;
        const long long obatch = (long long)b0 * p.out_sb0 + (long long)b1 * p.out_sb1;
        const long long pbatch = (long long)b0 * p.pl_sb0 + (long long)b1 * p.pl_sb1;
        // residual / aux_in of the first chunk: in flight while the main loop is still running
        float4 pre[8];
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + q * 32 < p.M && n0 + half * cpw * 32 < p.N)
            epilogue_prefetch(p, lane, m0 + q * 32, n0 + half * cpw * 32, obatch, split, vec_ok, pre, bias4);
        mbar_wait(tmem_full_bar, 0);
        if (threadIdx.x == 64) YV_T(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (m0 + q * 32 < p.M) {                             // warp-uniform: rows of this lane quarter exist
#pragma unroll 1
            for (int cc = 0; cc < cpw; ++cc) {
                const int c = half * cpw + cc;
                const int nc = n0 + c * 32;
                if (nc >= p.N) break;                        // warp-uniform
                uint32_t raw[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), raw);
                epilogue_chunk(p, drop, stg, lane, raw, m0 + q * 32, nc, z, obatch, pbatch, split, vec_ok, pre, bias4);
                if (cc + 1 < cpw && nc + 32 < p.N)
                    epilogue_prefetch(p, lane, m0 + q * 32, nc + 32, obatch, split, vec_ok, pre, bias4);
            }
        }
    }

    if (threadIdx.x == 64) YV_T(5);
    // both CTAs: nobody may leave (or free TMEM) while the peer's tensor core can still read this CTA's operands
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (threadIdx.x == 0) YV_T(6);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)pair_n)
                     : "memory");
    }
}

}  // namespace

void yv_count_launch();

static int pair_width(const YvGemm* g, int pair_n) {
    if (pair_n != 0) return pair_n;
    // 256-wide pair tiles halve the operand bytes per FLOP but also halve the CTA count: take them only when
    // the 128-wide tiling would need more than one resident wave
    const long long batch = g->a.nb0 * g->a.nb1;
    const long long ctas128 = 2LL * ((g->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((g->N + 127) / 128) * batch;
    return ctas128 > 2 * 148 ? 256 : 128;
}

extern "C" int yv_gemm_pair_splits(const YvGemm* g, int pair_n) {
    pair_n = pair_width(g, pair_n);
    const int pairs = ((g->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((g->N + pair_n - 1) / pair_n);
    int kbps;
    return plan_split_k(g, 2 * pairs, (g->K + BLOCK_K - 1) / BLOCK_K, &kbps);
}

// pair_n: 128 or 256 (0 = choose).  Un-batched and batched problems alike; M is tiled in 256-row pair tiles.
extern "C" int yv_gemm_pair(const YvGemm* g, int pair_n, yv_stream_t stream) {
    YV_CHECK(g != nullptr, "yv_gemm: NULL args");
    YV_CHECK(g->passes == 1 || g->passes == 3, "yv_gemm: passes must be 1 or 3 (got %d)", g->passes);
    YV_CHECK(g->M > 0 && g->N > 0 && g->K > 0, "yv_gemm: empty problem M=%d N=%d K=%d", g->M, g->N, g->K);
    YV_CHECK(g->out32 || g->out_planes, "yv_gemm: no output requested");
    YV_CHECK(pair_n == 0 || pair_n == 128 || pair_n == 256, "yv_gemm_pair: pair_n must be 0, 128 or 256");
    if (get_encode()) return 1;
    const YvOperand &a = g->a, &b = g->b;
    YV_CHECK((a.mn_major ? a.inner : a.rows) == g->M && (a.mn_major ? a.rows : a.inner) == g->K,
             "yv_gemm: A extents (%lld x %lld, mn_major=%d) do not match M=%d K=%d", (long long)a.rows, (long long)a.inner,
             a.mn_major, g->M, g->K);
    YV_CHECK((b.mn_major ? b.inner : b.rows) == g->N && (b.mn_major ? b.rows : b.inner) == g->K,
             "yv_gemm: B extents (%lld x %lld, mn_major=%d) do not match N=%d K=%d", (long long)b.rows, (long long)b.inner,
             b.mn_major, g->N, g->K);
    YV_CHECK(a.nb0 == b.nb0 && a.nb1 == b.nb1, "yv_gemm: batch counts differ");
    const long long batch = a.nb0 * a.nb1;
    const int tiles_mp = (g->M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    pair_n = pair_width(g, pair_n);
    CUtensorMap ma, mb;
    if (make_map(&ma, a, g->passes, "A", BLOCK_K, BLOCK_M)) return 1;
    if (make_map(&mb, b, g->passes, "B", BLOCK_K, pair_n / 2)) return 1;

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
    p.pair_n = pair_n;
    YV_CHECK((g->act != YV_ACT_MUL_GELU_GRAD && g->act != YV_ACT_MUL_RELU_MASK) || g->aux_in,
             "yv_gemm: act %d needs aux_in", g->act);

    const int pairs = tiles_mp * ((g->N + pair_n - 1) / pair_n);
    const int total_kb = (g->K + BLOCK_K - 1) / BLOCK_K;
    p.splits = plan_split_k(g, 2 * pairs, total_kb, &p.kb_per_split);
    const long long total_pairs = (long long)pairs * batch * p.splits;
    YV_CHECK(2 * total_pairs < 2147483647LL, "yv_gemm: too many tiles");
    p.total_tiles = (int)total_pairs;
    // as many ring stages as shared memory holds (one CTA per SM); the 32-deep build keeps 3 stages on multi-wave
    // launches so that two CTAs share an SM
    const int stage_b = g->passes == 3 ? PCfg<3>::stage_bytes(pair_n) : PCfg<1>::stage_bytes(pair_n);
    int stages = (SMEM_LIMIT - 1024 - 256) / stage_b;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (BLOCK_K == 32 && 2 * total_pairs > 148 && stages > 3) stages = 3;
    p.stages = stages;
    dim3 grid((unsigned)(2 * total_pairs), 1, 1);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (p.splits > 1 && !g->out32_zeroed)
        YV_CUDA(cudaMemset2DAsync(g->out32, sizeof(float) * g->ld_out, 0, sizeof(float) * g->N, g->M, st));
    {   // per device, behind a mutex (one host thread per device under nn.DataParallel)
        static bool attr_set[64] = {};
        static std::mutex mu;
        int dev = 0;
        YV_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(mu);
        if (!attr_set[dev & 63]) {
            YV_CUDA(cudaFuncSetAttribute(yv_gemm_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
            YV_CUDA(cudaFuncSetAttribute(yv_gemm_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
            attr_set[dev & 63] = true;
        }
    }
    if (g->passes == 3)
        YV_CUDA(yv_launch(yv_gemm_pair_kernel<3>, grid, dim3(NUM_THREADS), PCfg<3>::smem_bytes(p.stages, pair_n), st, ma, mb, p));
    else
        YV_CUDA(yv_launch(yv_gemm_pair_kernel<1>, grid, dim3(NUM_THREADS), PCfg<1>::smem_bytes(p.stages, pair_n), st, ma, mb, p));
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}
