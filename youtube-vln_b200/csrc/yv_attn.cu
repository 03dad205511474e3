// yv_attn: fused scaled-dot-product attention, forward and backward, on tcgen05 tensor cores.  sm_100a only.
//
//   forward :  O = dropout(softmax(Q K^T * scale + mask)) V            (one launch; the probabilities never leave the SM)
//   backward:  dQ, dK, dV from dO with the probabilities recomputed from Q, K and the saved row log-sum-exp
//
// replaces, in the reference, vilbert/vilbert.py:294-306 (BertSelfAttention), :423-435 (BertImageSelfAttention) and
// :577-616 (both directions of BertBiAttention) and their autograd backward -- on the un-fused path that is
// GEMM -> fp32 scores in HBM -> softmax kernel -> bf16 planes in HBM -> GEMM (3 launches forward, 6 backward).
//
// Work decomposition: one CTA per (128-query tile, head, pair); keys are streamed in chunks of 64.
//   warp 8 (one lane): TMA producer and tcgen05.mma issuer.  Q (and dO) tiles stay resident in shared memory, K / V
//                      chunks stream through single buffers (the next K chunk is loaded as soon as S = Q K^T of the
//                      current one has retired, the next V chunk as soon as P V has).
//   warps 0-7        : softmax.  Warp w owns TMEM lanes 32*(w&3).. (query rows) and key columns 32*(w>>2).. of the
//                      chunk: S comes out of TMEM with one tcgen05.ld.32x32b.x32, exp / dropout / hi-lo split run in
//                      registers, P goes back to shared memory in the 128B-swizzled K-major operand layout and is the
//                      A operand of the P V product (accumulated in TMEM across chunks, online-softmax rescaling of
//                      the accumulator only when a row maximum grows by more than 8).
// Operands are bf16 hi/lo plane pairs; products are contracted as lo*hi + hi*lo + hi*hi (PASSES = 3, parity mode) or
// hi*hi only (PASSES = 1).  A [rows x 64] 128B-swizzled tile is addressed as a K-major operand (contraction along the
// 64-element direction) or as an MN-major operand (contraction along rows) just by the descriptor, so the backward
// needs no transposed copies: dV^T = dO^T Pd, dK^T = Q^T dS and dQ = dS K all read the tiles the forward products use.
// dK / dV of one (pair, head) receive contributions from every query tile: they are reduced in an fp32 scratch with
// red.global.add and converted to planes by the last CTA of that (pair, head) (atomic ticket).
// Dropout masks use the same counter-based hash and element index (row * Tk + key) as yv_softmax_fwd / _bwd.
#include "yv_gemm_common.cuh"

namespace {

constexpr int QT = 128;            // query rows per CTA (UMMA M)
constexpr int KC = 64;             // keys per chunk (UMMA N of the score product, one 128-byte swizzled row)
constexpr int ATT_THREADS = 288;   // 8 softmax warps + 1 control warp
constexpr int SM_THREADS = 256;
constexpr uint32_t Q_BLK = QT * 128;    // bytes of a [128 rows x 64 bf16] block
constexpr uint32_t KV_BLK = KC * 128;   // bytes of a [64 rows x 64 bf16] block
constexpr float RESCALE_THRESHOLD = 8.f;

struct PlaneView {
    __nv_bfloat16* ptr;            // hi plane, element (row 0, head 0, d 0)
    long long ld, plane_stride;
};

struct AttnParams {
    int Tq, Tk, heads, pairs;
    float scale;
    const float* mask;             // [pairs, Tk] additive mask or nullptr
    float drop_p;
    unsigned drop_site;
    const unsigned long long* rng;
    // forward outputs
    PlaneView o;                   // context planes, row = pair * Tq + q, column = head * dh + d
    float* o32;
    long long o32_ld;
    float* lse;                    // [pairs * heads * Tq]
    // backward inputs / outputs
    PlaneView d_o, fwd_o;          // dO and O as plain global views (row dot products)
    PlaneView dq, dk, dv;
    float* dkv32;                  // fp32 scratch [pairs * Tk, dkv_ld]: dK at column dk_col, dV at dv_col (+ head * dh + d)
    long long dkv_ld;
    int dk_col, dv_col;
    unsigned* tickets;             // [pairs * heads], zero-initialised
};

YV_DEVINL void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
YV_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
YV_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
YV_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
YV_DEVINL void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
YV_DEVINL void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

YV_DEVINL uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// K-major descriptor of a [rows x 64] swizzled block (k-step s: +32 bytes); MN-major descriptor of the same bytes
// (contraction along rows; k-step s: +16 rows = 2048 bytes; `lbo` = distance to the next 64-element chunk along M / N)
YV_DEVINL uint64_t desc_k(uint32_t addr, int s) { return make_desc(addr, 16, 1024, 2) + (uint64_t)(2 * s); }
YV_DEVINL uint64_t desc_mn(uint32_t addr, uint32_t lbo, int s) { return make_desc(addr, lbo, 1024, 2) + (uint64_t)(128 * s); }
YV_DEVINL uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// one k-step of a (possibly three-term) product: lo*hi, hi*lo, then the dominant hi*hi
template <int PASSES>
YV_DEVINL void mma_terms(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                         uint32_t& accum) {
    if (PASSES == 3) {
        umma_bf16(tmem_d, a_lo, b_hi, idesc, accum);
        accum = 1;
        umma_bf16(tmem_d, a_hi, b_lo, idesc, 1);
    }
    umma_bf16(tmem_d, a_hi, b_hi, idesc, accum);
    accum = 1;
}

// D[128 x 64] = A[128 x DH] . B[64 x DH]^T, both operands K-major blocks [rows x 64] indexed [plane][DH / 64]
template <int DH, int PASSES>
YV_DEVINL void mma_rows_x_rows(uint32_t tmem_d, uint32_t a_base, uint32_t a_blk, uint32_t b_base, uint32_t b_blk) {
    constexpr int DB = DH / 64;
    const uint32_t idesc = make_idesc(QT, KC, 0, 0);
    uint32_t accum = 0;
#pragma unroll
    for (int b = 0; b < DB; ++b)
#pragma unroll
        for (int s = 0; s < 4; ++s)
            mma_terms<PASSES>(tmem_d, desc_k(a_base + b * a_blk, s), desc_k(a_base + (DB + b) * a_blk, s),
                              desc_k(b_base + b * b_blk, s), desc_k(b_base + (DB + b) * b_blk, s), idesc, accum);
}

// the 32 keys [key0, key0 + 32) of query row `row` as an operand tile row: element (row, key) of a [128 x 64] block at
// row * 128 + ((key / 8) ^ (row % 8)) * 16 + (key % 8) * 2 (the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B)
template <int PASSES>
YV_DEVINL void store_tile_row(uint32_t tile, int row, int half, const float* x) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            yv_split(x[8 * g + 2 * e], h0, l0);
            yv_split(x[8 * g + 2 * e + 1], h1, l1);
            hi[e] = pack_bf16(h0, h1);
            lo[e] = pack_bf16(l0, l1);
        }
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((4 * half + g) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                     "r"(hi[3])
                     : "memory");
        if (PASSES == 3)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + Q_BLK + off), "r"(lo[0]), "r"(lo[1]),
                         "r"(lo[2]), "r"(lo[3])
                         : "memory");
    }
}

// 32 consecutive fp32 values -> hi / lo planes at `dst` (16-byte aligned), optional fp32 copy
template <int PASSES>
YV_DEVINL void store_row32(__nv_bfloat16* dst, long long plane_stride, float* dst32, const float* x) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint4 hv, lv;
        uint32_t* hp = reinterpret_cast<uint32_t*>(&hv);
        uint32_t* lp = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            yv_split(x[8 * g + 2 * e], h0, l0);
            yv_split(x[8 * g + 2 * e + 1], h1, l1);
            hp[e] = pack_bf16(h0, h1);
            lp[e] = pack_bf16(l0, l1);
        }
        *reinterpret_cast<uint4*>(dst + 8 * g) = hv;
        *reinterpret_cast<uint4*>(dst + plane_stride + 8 * g) = lv;   // the lo plane is kept current in both modes
    }
    if (dst32) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(dst32 + 4 * g) = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
    }
}

template <int DH, int PASSES>
struct FwdCfg {
    static constexpr int PL = PASSES == 3 ? 2 : 1;
    static constexpr int DB = DH / 64;
    static constexpr uint32_t HEADER = 4096;                       // barriers, TMEM slot, row-statistic exchange
    static constexpr uint32_t SQ = 0;
    static constexpr uint32_t SK = SQ + PL * DB * Q_BLK;
    static constexpr uint32_t SV = SK + PL * DB * KV_BLK;
    static constexpr uint32_t SP = SV + PL * DB * KV_BLK;
    static constexpr uint32_t TILES = SP + PL * Q_BLK;
    static constexpr uint32_t SMEM = HEADER + 1024 + TILES;
    static constexpr int TMEM_COLS = 256;                           // S0 [0,64) S1 [64,128) O [128, 128 + DH)
    static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};

// ------------------------------------------------------------------------------------------- forward
template <int DH, int PASSES>
__global__ void __launch_bounds__(ATT_THREADS, 1)
yv_attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
    using C = FwdCfg<DH, PASSES>;
    constexpr int PL = C::PL, DB = C::DB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);         // 0 Q, 1 K, 2 V, 3-4 S[2], 5 P, 6 O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
    float* xm = reinterpret_cast<float*>(smem_raw + 128);           // [2 (chunk parity)][2 (column half)][128 rows]
    float* xl = xm + 2 * 2 * QT;                                    // [2][128]
    const uint32_t tiles = (smem_u32(smem_raw) + C::HEADER + 1023u) & ~1023u;
    const uint32_t sQ = tiles + C::SQ, sK = tiles + C::SK, sV = tiles + C::SV, sP = tiles + C::SP;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, head = blockIdx.y, pair = blockIdx.z;
    const int q0 = qtile * QT;
    const int nchunks = (p.Tk + KC - 1) / KC;
    yv_pdl_trigger();

    if (threadIdx.x == 0) {
        mbar_init(bar(0), 1);
        mbar_init(bar(1), 1);
        mbar_init(bar(2), 1);
        mbar_init(bar(3), 1);
        mbar_init(bar(4), 1);
        mbar_init(bar(5), 8);
        mbar_init(bar(6), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_v) : "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_o = tmem + 128;
    yv_pdl_wait();

    if (warp == 8) {
        if (lane == 0) {
            // ===================================== TMA + MMA issue =====================================
            auto load_kv = [&](const CUtensorMap* map, uint32_t dst, uint32_t b, int chunk) {
                mbar_expect_tx(b, PL * DB * KV_BLK);
#pragma unroll
                for (int pl = 0; pl < PL; ++pl)
#pragma unroll
                    for (int d = 0; d < DB; ++d)
                        tma_load_5d(dst + (pl * DB + d) * KV_BLK, map, b, d * 64, chunk * KC, head, pair, pl);
            };
            mbar_expect_tx(bar(0), PL * DB * Q_BLK);
#pragma unroll
            for (int pl = 0; pl < PL; ++pl)
#pragma unroll
                for (int d = 0; d < DB; ++d)
                    tma_load_5d(sQ + (pl * DB + d) * Q_BLK, &map_q, bar(0), d * 64, q0, head, pair, pl);
            load_kv(&map_k, sK, bar(1), 0);
            load_kv(&map_v, sV, bar(2), 0);
            mbar_wait(bar(0), 0);
            mbar_wait(bar(1), 0);
            tc_fence_after();
            mma_rows_x_rows<DH, PASSES>(tmem, sQ, Q_BLK, sK, KV_BLK);
            umma_commit(bar(3));
            const uint32_t idesc_pv = make_idesc(QT, DH, 0, 1);
            for (int j = 0; j < nchunks; ++j) {
                if (j + 1 < nchunks) {
                    mbar_wait(bar(3 + (j & 1)), (uint32_t)((j >> 1) & 1));   // S_j retired: the K buffer is free
                    load_kv(&map_k, sK, bar(1), j + 1);
                    mbar_wait(bar(1), (uint32_t)((j + 1) & 1));
                    tc_fence_after();
                    mma_rows_x_rows<DH, PASSES>(tmem + (uint32_t)(((j + 1) & 1) * KC), sQ, Q_BLK, sK, KV_BLK);
                    umma_commit(bar(3 + ((j + 1) & 1)));
                }
                mbar_wait(bar(5), (uint32_t)(j & 1));                        // P_j in shared memory, O rescaled
                mbar_wait(bar(2), (uint32_t)(j & 1));                        // V_j landed
                tc_fence_after();
                const int kc = min(KC, p.Tk - j * KC);
                const int ksteps = (kc + 15) >> 4;
                uint32_t accum = j > 0 ? 1u : 0u;
                for (int s = 0; s < ksteps; ++s)
                    mma_terms<PASSES>(tmem_o, desc_k(sP, s), desc_k(sP + Q_BLK, s), desc_mn(sV, KV_BLK, s),
                                      desc_mn(sV + DB * KV_BLK, KV_BLK, s), idesc_pv, accum);
                umma_commit(bar(6));
                if (j + 1 < nchunks) {
                    mbar_wait(bar(6), (uint32_t)(j & 1));                    // P V retired: V and P buffers are free
                    load_kv(&map_v, sV, bar(2), j + 1);
                }
            }
        }
    } else {
        // ========================================= softmax ============================================
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;                  // row of the tile = TMEM lane
        const int qrow = q0 + row;
        const bool row_ok = qrow < p.Tq;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t drop_row = (uint32_t)(((long long)(pair * p.heads + head) * p.Tq + qrow) * p.Tk);
        const float* mrow = p.mask ? p.mask + (long long)pair * p.Tk : nullptr;
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nchunks; ++j) {
            mbar_wait(bar(3 + (j & 1)), (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            uint32_t raw[32];
            tmem_ld32(tmem + lane_addr + (uint32_t)((j & 1) * KC + half * 32), raw);
            const int key0 = j * KC + half * 32;
            float x[32];
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int key = key0 + i;
                x[i] = -INFINITY;
                if (key < p.Tk) x[i] = __uint_as_float(raw[i]) * p.scale + (mrow ? __ldg(mrow + key) : 0.f);
                mx = fmaxf(mx, x[i]);
            }
            float* xmj = xm + (j & 1) * 2 * QT;
            xmj[half * QT + row] = mx;
            softmax_bar();
            mx = fmaxf(mx, xmj[(half ^ 1) * QT + row]);
            // online softmax with a lazy reference maximum: the accumulator is rescaled only when the row maximum
            // grew by more than RESCALE_THRESHOLD (both column halves take the same decision from the same numbers)
            const float m_new = (j == 0 || mx > m_ref + RESCALE_THRESHOLD) ? mx : m_ref;
            const float alpha = (j == 0) ? 1.f : __expf(m_ref - m_new);
            m_ref = m_new;
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                x[i] = __expf(x[i] - m_new);                  // exp(-inf) = 0 for keys past Tk
                sum += x[i];
            }
            l_run = l_run * alpha + sum;
            if (drop.thresh) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x[i] *= yv_drop_mul(drop, drop_row + (uint32_t)(key0 + i));
            }
            if (j > 0) {
                mbar_wait(bar(6), (uint32_t)((j - 1) & 1));   // P V of the previous chunk retired
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
                    for (int g = 0; g < DH / 64; ++g) {
                        uint32_t o[32];
                        const uint32_t a = tmem_o + lane_addr + (uint32_t)(half * (DH / 2) + g * 32);
                        tmem_ld32(a, o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(a, o);
                    }
                }
            }
            store_tile_row<PASSES>(sP, row, half, x);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(5));
        }
        // ---- epilogue: O / l -> planes (+ fp32), log-sum-exp for the backward
        mbar_wait(bar(6), (uint32_t)((nchunks - 1) & 1));
        tc_fence_after();
        xl[half * QT + row] = l_run;
        softmax_bar();
        const float l_tot = l_run + xl[(half ^ 1) * QT + row];
        const float inv = 1.f / l_tot;
        const long long grow = (long long)pair * p.Tq + qrow;
#pragma unroll 1
        for (int g = 0; g < DH / 64; ++g) {
            uint32_t o[32];
            const int col = half * (DH / 2) + g * 32;
            tmem_ld32(tmem_o + lane_addr + (uint32_t)col, o);
            if (row_ok) {
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(o[i]) * inv;
                store_row32<PASSES>(p.o.ptr + grow * p.o.ld + head * DH + col, p.o.plane_stride,
                                    p.o32 ? p.o32 + grow * p.o32_ld + head * DH + col : nullptr, y);
            }
        }
        if (half == 0 && row_ok && p.lse) p.lse[(long long)(pair * p.heads + head) * p.Tq + qrow] = m_ref + logf(l_tot);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------- backward
template <int DH, int PASSES>
struct BwdCfg {
    static constexpr int PL = PASSES == 3 ? 2 : 1;
    static constexpr int DB = DH / 64;
    static constexpr uint32_t HEADER = 1024;
    static constexpr uint32_t SQ = 0;
    static constexpr uint32_t SDO = SQ + PL * DB * Q_BLK;
    static constexpr uint32_t SK = SDO + PL * DB * Q_BLK;
    static constexpr uint32_t SV = SK + PL * DB * KV_BLK;
    static constexpr uint32_t ST = SV + PL * DB * KV_BLK;            // Pd, then dS
    static constexpr uint32_t TILES = ST + PL * Q_BLK;
    static constexpr uint32_t SMEM = HEADER + 1024 + TILES;
    static constexpr int TMEM_COLS = 512;   // S [0,64) dPd [64,128) dV^T [128,192) dK^T [192,256) dQ [256, 256 + DH)
    static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};

// rows of dO . O (the softmax-backward row term) for query row `grow`, from the global plane pairs
template <int DH>
YV_DEVINL float row_dot(const PlaneView& a, const PlaneView& b, long long grow, int head) {
    const __nv_bfloat16* pa = a.ptr + grow * a.ld + head * DH;
    const __nv_bfloat16* pb = b.ptr + grow * b.ld + head * DH;
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < DH; c += 8) {
        const uint4 ah = *reinterpret_cast<const uint4*>(pa + c), al = *reinterpret_cast<const uint4*>(pa + a.plane_stride + c);
        const uint4 bh = *reinterpret_cast<const uint4*>(pb + c), bl = *reinterpret_cast<const uint4*>(pb + b.plane_stride + c);
        const uint32_t* ahp = reinterpret_cast<const uint32_t*>(&ah);
        const uint32_t* alp = reinterpret_cast<const uint32_t*>(&al);
        const uint32_t* bhp = reinterpret_cast<const uint32_t*>(&bh);
        const uint32_t* blp = reinterpret_cast<const uint32_t*>(&bl);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            // bf16 -> fp32 is a 16-bit shift
            const float a0 = __uint_as_float(ahp[e] << 16) + __uint_as_float(alp[e] << 16);
            const float a1 = __uint_as_float(ahp[e] & 0xffff0000u) + __uint_as_float(alp[e] & 0xffff0000u);
            const float b0 = __uint_as_float(bhp[e] << 16) + __uint_as_float(blp[e] << 16);
            const float b1 = __uint_as_float(bhp[e] & 0xffff0000u) + __uint_as_float(blp[e] & 0xffff0000u);
            acc = fmaf(a0, b0, acc);
            acc = fmaf(a1, b1, acc);
        }
    }
    return acc;
}

template <int DH, int PASSES>
__global__ void __launch_bounds__(ATT_THREADS, 1)
yv_attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ AttnParams p) {
    using C = BwdCfg<DH, PASSES>;
    constexpr int PL = C::PL, DB = C::DB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);         // 0 Q+dO, 1 K+V, 2 S+dPd, 3 Pd, 4 dV, 5 dS, 6 dK+dQ
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
    uint32_t* last_flag = reinterpret_cast<uint32_t*>(smem_raw + 68);
    const uint32_t tiles = (smem_u32(smem_raw) + C::HEADER + 1023u) & ~1023u;
    const uint32_t sQ = tiles + C::SQ, sDO = tiles + C::SDO, sK = tiles + C::SK, sV = tiles + C::SV, sT = tiles + C::ST;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, head = blockIdx.y, pair = blockIdx.z;
    const int q0 = qtile * QT;
    const int nchunks = (p.Tk + KC - 1) / KC;
    yv_pdl_trigger();

    if (threadIdx.x == 0) {
        mbar_init(bar(0), 1);
        mbar_init(bar(1), 1);
        mbar_init(bar(2), 1);
        mbar_init(bar(3), 8);
        mbar_init(bar(4), 1);
        mbar_init(bar(5), 8);
        mbar_init(bar(6), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_do) : "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_s = tmem, tm_dp = tmem + 64, tm_dv = tmem + 128, tm_dk = tmem + 192, tm_dq = tmem + 256;
    yv_pdl_wait();

    if (warp == 8) {
        if (lane == 0) {
            // ===================================== TMA + MMA issue =====================================
            auto load_rows = [&](const CUtensorMap* map, uint32_t dst, uint32_t blk, uint32_t b, int row0) {
#pragma unroll
                for (int pl = 0; pl < PL; ++pl)
#pragma unroll
                    for (int d = 0; d < DB; ++d)
                        tma_load_5d(dst + (pl * DB + d) * blk, map, b, d * 64, row0, head, pair, pl);
            };
            mbar_expect_tx(bar(0), 2 * PL * DB * Q_BLK);
            load_rows(&map_q, sQ, Q_BLK, bar(0), q0);
            load_rows(&map_do, sDO, Q_BLK, bar(0), q0);
            mbar_expect_tx(bar(1), 2 * PL * DB * KV_BLK);
            load_rows(&map_k, sK, KV_BLK, bar(1), 0);
            load_rows(&map_v, sV, KV_BLK, bar(1), 0);
            mbar_wait(bar(0), 0);
            const uint32_t idesc_t = make_idesc(128, KC, 1, 1);          // dV^T / dK^T: both operands MN-major
            const uint32_t idesc_dq = make_idesc(QT, DH, 0, 1);
            for (int j = 0; j < nchunks; ++j) {
                const uint32_t ph = (uint32_t)(j & 1);
                mbar_wait(bar(1), ph);
                tc_fence_after();
                mma_rows_x_rows<DH, PASSES>(tm_s, sQ, Q_BLK, sK, KV_BLK);      // S   = Q  K^T
                mma_rows_x_rows<DH, PASSES>(tm_dp, sDO, Q_BLK, sV, KV_BLK);    // dPd = dO V^T
                umma_commit(bar(2));
                mbar_wait(bar(3), ph);                                          // Pd tile written
                tc_fence_after();
                {   // dV^T [DH x 64] = dO^T . Pd : contraction over the 128 query rows
                    uint32_t accum = 0;
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mma_terms<PASSES>(tm_dv, desc_mn(sDO, Q_BLK, s), desc_mn(sDO + DB * Q_BLK, Q_BLK, s),
                                          desc_mn(sT, Q_BLK, s), desc_mn(sT + Q_BLK, Q_BLK, s), idesc_t, accum);
                }
                umma_commit(bar(4));
                mbar_wait(bar(5), ph);                                          // dS tile written (dV^T retired before)
                tc_fence_after();
                {   // dK^T [DH x 64] = Q^T . dS
                    uint32_t accum = 0;
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mma_terms<PASSES>(tm_dk, desc_mn(sQ, Q_BLK, s), desc_mn(sQ + DB * Q_BLK, Q_BLK, s),
                                          desc_mn(sT, Q_BLK, s), desc_mn(sT + Q_BLK, Q_BLK, s), idesc_t, accum);
                }
                {   // dQ [128 x DH] += dS . K : contraction over the keys of this chunk
                    const int kc = min(KC, p.Tk - j * KC);
                    const int ksteps = (kc + 15) >> 4;
                    uint32_t accum = j > 0 ? 1u : 0u;
                    for (int s = 0; s < ksteps; ++s)
                        mma_terms<PASSES>(tm_dq, desc_k(sT, s), desc_k(sT + Q_BLK, s), desc_mn(sK, KV_BLK, s),
                                          desc_mn(sK + DB * KV_BLK, KV_BLK, s), idesc_dq, accum);
                }
                umma_commit(bar(6));
                if (j + 1 < nchunks) {
                    mbar_wait(bar(6), ph);                                      // K, V and the dS tile are free
                    mbar_expect_tx(bar(1), 2 * PL * DB * KV_BLK);
                    load_rows(&map_k, sK, KV_BLK, bar(1), (j + 1) * KC);
                    load_rows(&map_v, sV, KV_BLK, bar(1), (j + 1) * KC);
                }
            }
        }
    } else {
        // ========================================= softmax ============================================
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const int qrow = q0 + row;
        const bool row_ok = qrow < p.Tq;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t drop_row = (uint32_t)(((long long)(pair * p.heads + head) * p.Tq + qrow) * p.Tk);
        const float* mrow = p.mask ? p.mask + (long long)pair * p.Tk : nullptr;
        const long long grow = (long long)pair * p.Tq + qrow;
        // rows past Tq (zero-filled by TMA) get lse = +inf: their probabilities and dS are exactly zero
        float lse = INFINITY, delta = 0.f;
        if (row_ok) {
            lse = p.lse[(long long)(pair * p.heads + head) * p.Tq + qrow];
            delta = row_dot<DH>(p.d_o, p.fwd_o, grow, head);
        }
        const int d_lane = quarter * 32 + lane;                     // TMEM lane of dV^T / dK^T = head dimension index
        float* dkv_base = p.dkv32 + (long long)pair * p.Tk * p.dkv_ld + head * DH + d_lane;
        for (int j = 0; j < nchunks; ++j) {
            const uint32_t ph = (uint32_t)(j & 1);
            const int key0 = j * KC + half * 32;
            mbar_wait(bar(2), ph);
            tc_fence_after();
            float pd[32], ds[32];
            {
                uint32_t raw[32];
                tmem_ld32(tm_s + lane_addr + (uint32_t)(half * 32), raw);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int key = key0 + i;
                    float pr = 0.f;
                    if (key < p.Tk)
                        pr = __expf(__uint_as_float(raw[i]) * p.scale + (mrow ? __ldg(mrow + key) : 0.f) - lse);
                    pd[i] = pr;
                }
                tmem_ld32(tm_dp + lane_addr + (uint32_t)(half * 32), raw);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float mult = drop.thresh ? yv_drop_mul(drop, drop_row + (uint32_t)(key0 + i)) : 1.f;
                    const float pr = pd[i];
                    ds[i] = p.scale * pr * (__uint_as_float(raw[i]) * mult - delta);
                    pd[i] = pr * mult;
                }
            }
            if (j > 0) mbar_wait(bar(6), (uint32_t)((j - 1) & 1));   // previous dK^T / dQ products retired: tile is free
            store_tile_row<PASSES>(sT, row, half, pd);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(3));
            mbar_wait(bar(4), ph);                                    // dV^T complete, the Pd tile has been consumed
            tc_fence_after();
            store_tile_row<PASSES>(sT, row, half, ds);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(5));
            // drain dV^T (lane = d, column = key) into the fp32 scratch while dK^T / dQ run
            if (DH == 128 || quarter < 2) {
                uint32_t raw[32];
                tmem_ld32(tm_dv + lane_addr + (uint32_t)(half * 32), raw);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + i < p.Tk)
                        atomicAdd(dkv_base + (long long)(key0 + i) * p.dkv_ld + p.dv_col, __uint_as_float(raw[i]));
            }
            mbar_wait(bar(6), ph);
            tc_fence_after();
            if (DH == 128 || quarter < 2) {
                uint32_t raw[32];
                tmem_ld32(tm_dk + lane_addr + (uint32_t)(half * 32), raw);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (key0 + i < p.Tk)
                        atomicAdd(dkv_base + (long long)(key0 + i) * p.dkv_ld + p.dk_col, __uint_as_float(raw[i]));
            }
            tc_fence_before();
        }
        // ---- dQ tile -> planes
#pragma unroll 1
        for (int g = 0; g < DH / 64; ++g) {
            uint32_t o[32];
            const int col = half * (DH / 2) + g * 32;
            tmem_ld32(tm_dq + lane_addr + (uint32_t)col, o);
            if (row_ok) {
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(o[i]);
                store_row32<PASSES>(p.dq.ptr + grow * p.dq.ld + head * DH + col, p.dq.plane_stride, nullptr, y);
            }
        }
        // ---- the last query tile of this (pair, head) converts the reduced dK / dV to planes
        __threadfence();
        softmax_bar();
        if (threadIdx.x == 0) *last_flag = (atomicAdd(p.tickets + pair * p.heads + head, 1u) == gridDim.x - 1) ? 1u : 0u;
        softmax_bar();
        if (*last_flag) {
            __threadfence();
            constexpr int V4 = DH / 4;
            for (int idx = threadIdx.x; idx < p.Tk * V4; idx += SM_THREADS) {
                const int key = idx / V4, c = (idx % V4) * 4;
                const long long krow = (long long)pair * p.Tk + key;
                const float* src = p.dkv32 + krow * p.dkv_ld + head * DH + c;
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (t ? p.dv_col : p.dk_col)));
                    const PlaneView& out = t ? p.dv : p.dk;
                    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                    yv_split(v.x, h0, l0); yv_split(v.y, h1, l1); yv_split(v.z, h2, l2); yv_split(v.w, h3, l3);
                    __nv_bfloat16* dst = out.ptr + krow * out.ld + head * DH + c;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
                    *reinterpret_cast<uint2*>(dst + out.plane_stride) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------- host
int view_map(CUtensorMap* map, const YvHeadView& v, int pairs, int heads, int dh, int passes, int box_rows, const char* which) {
    YV_CHECK(v.ptr != nullptr && v.rows > 0, "yv_attn: view %s is empty", which);
    YvOperand o;
    o.ptr = v.ptr;
    o.inner = dh;
    o.rows = v.rows;
    o.ld = v.ld;
    o.nb0 = heads;
    o.sb0 = dh;
    o.nb1 = pairs;
    o.sb1 = v.pair_stride;
    o.plane_stride = v.plane_stride;
    o.mn_major = 0;
    return make_map(map, o, passes, which, 64, box_rows);
}

PlaneView plane_view(const YvHeadView& v) {
    PlaneView r;
    r.ptr = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(v.ptr));
    r.ld = v.ld;
    r.plane_stride = v.plane_stride;
    return r;
}

// per-device one-time opt-in to > 48 KB of dynamic shared memory (cudaFuncSetAttribute is per device)
template <typename K>
int set_smem_once(K kernel, int bytes, unsigned long long* done_mask) {
    int dev = 0;
    YV_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return 0;
    YV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
    return 0;
}

template <int DH, int PASSES>
int launch_fwd(const YvAttnFwd* a, const AttnParams& p, cudaStream_t st) {
    using C = FwdCfg<DH, PASSES>;
    CUtensorMap mq, mk, mv;
    if (view_map(&mq, a->q, a->pairs, a->heads, DH, PASSES, QT, "Q")) return 1;
    if (view_map(&mk, a->k, a->pairs, a->heads, DH, PASSES, KC, "K")) return 1;
    if (view_map(&mv, a->v, a->pairs, a->heads, DH, PASSES, KC, "V")) return 1;
    static unsigned long long done = 0;
    if (set_smem_once(yv_attn_fwd_kernel<DH, PASSES>, (int)C::SMEM, &done)) return 2;
    dim3 grid((unsigned)((p.Tq + QT - 1) / QT), (unsigned)a->heads, (unsigned)a->pairs);
    YV_CUDA(yv_launch(yv_attn_fwd_kernel<DH, PASSES>, grid, dim3(ATT_THREADS), C::SMEM, st, mq, mk, mv, p));
    return 0;
}

template <int DH, int PASSES>
int launch_bwd(const YvAttnBwd* a, const AttnParams& p, cudaStream_t st) {
    using C = BwdCfg<DH, PASSES>;
    CUtensorMap mq, mk, mv, md;
    if (view_map(&mq, a->q, a->pairs, a->heads, DH, PASSES, QT, "Q")) return 1;
    if (view_map(&mk, a->k, a->pairs, a->heads, DH, PASSES, KC, "K")) return 1;
    if (view_map(&mv, a->v, a->pairs, a->heads, DH, PASSES, KC, "V")) return 1;
    if (view_map(&md, a->dout, a->pairs, a->heads, DH, PASSES, QT, "dO")) return 1;
    static unsigned long long done = 0;
    if (set_smem_once(yv_attn_bwd_kernel<DH, PASSES>, (int)C::SMEM, &done)) return 2;
    dim3 grid((unsigned)((p.Tq + QT - 1) / QT), (unsigned)a->heads, (unsigned)a->pairs);
    YV_CUDA(yv_launch(yv_attn_bwd_kernel<DH, PASSES>, grid, dim3(ATT_THREADS), C::SMEM, st, mq, mk, mv, md, p));
    return 0;
}

int check_view(const YvHeadView& v, int heads, int dh, const char* which) {
    YV_CHECK(v.ptr != nullptr, "yv_attn: %s is NULL", which);
    YV_CHECK(((uintptr_t)v.ptr & 15) == 0 && (v.ld & 7) == 0 && (v.plane_stride & 7) == 0 && (v.pair_stride & 7) == 0,
             "yv_attn: %s must be 16-byte aligned with ld / strides multiples of 8 elements", which);
    YV_CHECK(v.ld >= (int64_t)heads * dh, "yv_attn: %s ld=%lld is smaller than heads*dh=%d", which, (long long)v.ld, heads * dh);
    return 0;
}

}  // namespace

void yv_count_launch();

extern "C" int yv_attn_supported(int32_t dh, int32_t passes) {
    return (dh == 64 || dh == 128) && (passes == 1 || passes == 3);
}

extern "C" int yv_attn_fwd(const YvAttnFwd* a, yv_stream_t stream) {
    YV_CHECK(a != nullptr, "yv_attn_fwd: NULL args");
    YV_CHECK(yv_attn_supported(a->dh, a->passes), "yv_attn_fwd: head size %d / passes %d not supported (64 or 128; 1 or 3)",
             a->dh, a->passes);
    YV_CHECK(a->pairs > 0 && a->heads > 0 && a->q.rows > 0 && a->k.rows > 0 && a->k.rows == a->v.rows,
             "yv_attn_fwd: bad extents");
    YV_CHECK(a->out_planes != nullptr && (a->ld_out & 7) == 0 && (a->out_plane_stride & 7) == 0 &&
             ((uintptr_t)a->out_planes & 15) == 0, "yv_attn_fwd: output planes missing or misaligned");
    YV_CHECK(a->out32 == nullptr || ((a->ld_out32 & 3) == 0 && ((uintptr_t)a->out32 & 15) == 0),
             "yv_attn_fwd: fp32 output misaligned");
    YV_CHECK((long long)a->pairs * a->heads * a->q.rows * a->k.rows < 4294967296LL,
             "yv_attn_fwd: more than 2^32 probabilities per call (dropout counter range)");
    if (check_view(a->q, a->heads, a->dh, "Q") || check_view(a->k, a->heads, a->dh, "K") ||
        check_view(a->v, a->heads, a->dh, "V"))
        return 1;
    if (get_encode()) return 1;
    AttnParams p = {};
    p.Tq = a->q.rows; p.Tk = a->k.rows; p.heads = a->heads; p.pairs = a->pairs;
    p.scale = a->scale;
    p.mask = a->mask;
    p.drop_p = a->drop_p; p.drop_site = a->drop_site;
    p.rng = reinterpret_cast<const unsigned long long*>(a->rng);
    p.o.ptr = reinterpret_cast<__nv_bfloat16*>(a->out_planes); p.o.ld = a->ld_out; p.o.plane_stride = a->out_plane_stride;
    p.o32 = a->out32; p.o32_ld = a->ld_out32;
    p.lse = a->lse;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (a->dh == 128) rc = a->passes == 3 ? launch_fwd<128, 3>(a, p, st) : launch_fwd<128, 1>(a, p, st);
    else rc = a->passes == 3 ? launch_fwd<64, 3>(a, p, st) : launch_fwd<64, 1>(a, p, st);
    if (rc) return rc;
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}

extern "C" size_t yv_attn_bwd_workspace_bytes(int32_t pairs, int32_t heads, int32_t dh, int32_t Tk) {
    if (pairs <= 0 || heads <= 0 || dh <= 0 || Tk <= 0) return 0;
    // fp32 [pairs * Tk, 2 * heads * dh] (dK | dV) followed by one ticket per (pair, head)
    return (size_t)pairs * Tk * 2 * heads * dh * sizeof(float) + (size_t)pairs * heads * sizeof(uint32_t);
}

extern "C" int yv_attn_bwd(const YvAttnBwd* a, yv_stream_t stream) {
    YV_CHECK(a != nullptr, "yv_attn_bwd: NULL args");
    YV_CHECK(yv_attn_supported(a->dh, a->passes), "yv_attn_bwd: head size %d / passes %d not supported (64 or 128; 1 or 3)",
             a->dh, a->passes);
    YV_CHECK(a->pairs > 0 && a->heads > 0 && a->q.rows > 0 && a->k.rows > 0 && a->k.rows == a->v.rows &&
             a->dout.rows == a->q.rows && a->out.rows == a->q.rows && a->dq.rows == a->q.rows &&
             a->dk.rows == a->k.rows && a->dv.rows == a->k.rows, "yv_attn_bwd: bad extents");
    YV_CHECK(a->lse != nullptr, "yv_attn_bwd: the forward's row log-sum-exp is required");
    YV_CHECK((long long)a->pairs * a->heads * a->q.rows * a->k.rows < 4294967296LL,
             "yv_attn_bwd: more than 2^32 probabilities per call (dropout counter range)");
    if (check_view(a->q, a->heads, a->dh, "Q") || check_view(a->k, a->heads, a->dh, "K") ||
        check_view(a->v, a->heads, a->dh, "V") || check_view(a->dout, a->heads, a->dh, "dO") ||
        check_view(a->out, a->heads, a->dh, "O") || check_view(a->dq, a->heads, a->dh, "dQ") ||
        check_view(a->dk, a->heads, a->dh, "dK") || check_view(a->dv, a->heads, a->dh, "dV"))
        return 1;
    const size_t need = yv_attn_bwd_workspace_bytes(a->pairs, a->heads, a->dh, a->k.rows);
    YV_CHECK(a->workspace != nullptr && a->workspace_bytes >= need && ((uintptr_t)a->workspace & 15) == 0,
             "yv_attn_bwd: workspace of %zu zero-filled bytes required (got %zu)", need, (size_t)a->workspace_bytes);
    if (get_encode()) return 1;
    const int H = a->heads * a->dh;
    AttnParams p = {};
    p.Tq = a->q.rows; p.Tk = a->k.rows; p.heads = a->heads; p.pairs = a->pairs;
    p.scale = a->scale;
    p.mask = a->mask;
    p.drop_p = a->drop_p; p.drop_site = a->drop_site;
    p.rng = reinterpret_cast<const unsigned long long*>(a->rng);
    p.lse = const_cast<float*>(a->lse);
    p.d_o = plane_view(a->dout);
    p.fwd_o = plane_view(a->out);
    p.dq = plane_view(a->dq); p.dk = plane_view(a->dk); p.dv = plane_view(a->dv);
    p.dkv32 = reinterpret_cast<float*>(a->workspace);
    p.dkv_ld = 2 * H;
    p.dk_col = 0; p.dv_col = H;
    p.tickets = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(a->workspace) +
                                            (size_t)a->pairs * a->k.rows * 2 * H * sizeof(float));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (a->dh == 128) rc = a->passes == 3 ? launch_bwd<128, 3>(a, p, st) : launch_bwd<128, 1>(a, p, st);
    else rc = a->passes == 3 ? launch_bwd<64, 3>(a, p, st) : launch_bwd<64, 1>(a, p, st);
    if (rc) return rc;
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}
