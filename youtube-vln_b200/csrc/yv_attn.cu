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
//   warps 16, 17 (one lane each): TMA producers and tcgen05.mma issuers.  Issuing a 128x64x16 MMA costs one thread about
//                      60 cycles (descriptor moves to uniform registers), twice its tensor-pipe time, so the products
//                      of a chunk are split over two issuing threads: forward S = Q K^T (first, K double-buffered)
//                      and O += P V (second); backward S, dV^T and dPd, dK^T, dQ.  Q (and dO) tiles stay resident in
//                      shared memory, K / V chunks stream.  (The kernels are templates on the softmax warp count;
//                      YVB200_ATTN_WARPS=8 selects the 8-warp build, 32 keys per thread, for A/B timing.)
//   warps 0-15       : softmax.  Warp w owns TMEM lanes 32*(w&3).. (query rows) and key columns 16*(w>>2).. of the
//                      chunk: S comes out of TMEM with one tcgen05.ld.32x32b.x16, exp / dropout / hi-lo split run in
//                      registers, P goes back to shared memory in the 128B-swizzled K-major operand layout and is the
//                      A operand of the P V product (accumulated in TMEM across chunks, online-softmax rescaling of
//                      the accumulator only when a row maximum grows by more than 8).
// Operands are bf16 hi/lo plane pairs; products are contracted as lo*hi + hi*lo + hi*hi (PASSES = 3, parity mode) or
// hi*hi only (PASSES = 1).  A [rows x 64] 128B-swizzled tile is addressed as a K-major operand (contraction along the
// 64-element direction) or as an MN-major operand (contraction along rows) just by the descriptor, so the backward
// needs no transposed copies: dV^T = dO^T Pd, dK^T = Q^T dS and dQ = dS K all read the tiles the forward products use.
// dK / dV of one (pair, head) receive contributions from every query tile: each tile stores its partial sums to its own
// fp32 slab of the workspace (plain coalesced stores; global reductions run at ~1 lane per clock and SM) and the last CTA
// of that (pair, head) -- atomic ticket -- adds the slabs and writes the planes.
// Dropout masks use the same counter-based hash and element index (row * Tk + key) as yv_softmax_fwd / _bwd.
#include "yv_gemm_common.cuh"

namespace {

constexpr int QT = 128;            // query rows per CTA (UMMA M)
constexpr int KC = 64;             // keys per chunk (UMMA N of the score product, one 128-byte swizzled row)
// SW softmax warps (8 or 16) + 2 issuer warps (TMA + tcgen05.mma, one lane each).  Softmax warp w owns TMEM lanes
// 32 * (w & 3) .. (query rows) and the key columns [KPT * (w >> 2), +KPT) of a chunk, KPT = 64 / (SW / 4).
constexpr int att_threads(int sw) { return sw * 32 + 64; }
constexpr uint32_t Q_BLK = QT * 128;    // bytes of a [128 rows x 64 bf16] block
constexpr uint32_t KV_BLK = KC * 128;   // bytes of a [64 rows x 64 bf16] block
constexpr float RESCALE_THRESHOLD = 8.f;

// developer timing (tools/attn_timing.cu): clock64 stamps of CTA (0, 0, 0)
#ifdef YV_ATTN_TIMING
__device__ long long yv_adbg[128];
#define YV_AT(i) do { if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (i) < 128) yv_adbg[i] = clock64(); } while (0)
#else
#define YV_AT(i)
#endif

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
YV_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
YV_DEVINL void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int N> YV_DEVINL void tmem_ld(uint32_t taddr, uint32_t* v) { if (N == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v); }
template <int N> YV_DEVINL void tmem_st(uint32_t taddr, const uint32_t* v) { if (N == 32) tmem_st32(taddr, v); else tmem_st16(taddr, v); }
YV_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
YV_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
YV_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int SW> YV_DEVINL void softmax_bar() { asm volatile("bar.sync 1, %0;" ::"n"(SW * 32) : "memory"); }
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

// N consecutive keys (column group cg) of query row `row` as an operand tile row: element (row, key) of a [128 x 64] block at
// row * 128 + ((key / 8) ^ (row % 8)) * 16 + (key % 8) * 2 (the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B)
template <int PASSES, int N>
YV_DEVINL void store_tile_row(uint32_t tile, int row, int cg, const float* x) {
#pragma unroll
    for (int g = 0; g < N / 8; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) yv_split2(x[8 * g + 2 * e], x[8 * g + 2 * e + 1], hi[e], lo[e]);
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((((N / 8) * cg + g) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                     "r"(hi[3])
                     : "memory");
        if (PASSES == 3)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + Q_BLK + off), "r"(lo[0]), "r"(lo[1]),
                         "r"(lo[2]), "r"(lo[3])
                         : "memory");
    }
}

// Output tiles ([128 rows x DH], thread = row in registers) leave through a swizzled shared-memory staging area so that
// the global stores are whole rows: 16-byte chunk c of row r of plane pl sits at pl * 128 * DH * 2 + r * DH * 2 +
// ((c ^ (r & 7)) << 4).  stage_row writes N consecutive columns of one row, copy_out_rows writes the tile.
template <int DH, int N>
YV_DEVINL void stage_row(uint32_t stage, int row, int col0, const float* x) {
    constexpr uint32_t ROW_B = DH * 2, PLANE_B = QT * ROW_B;
#pragma unroll
    for (int g = 0; g < N / 8; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) yv_split2(x[8 * g + 2 * e], x[8 * g + 2 * e + 1], hi[e], lo[e]);
        const uint32_t off = (uint32_t)row * ROW_B + (uint32_t)((((col0 >> 3) + g) ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                     "r"(hi[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + PLANE_B + off), "r"(lo[0]), "r"(lo[1]),
                     "r"(lo[2]), "r"(lo[3])
                     : "memory");
    }
}
template <int DH, int SW>
YV_DEVINL void copy_out_rows(uint32_t stage, const PlaneView& out, long long grow0, int rows_valid, int head, int warp,
                             int lane) {
    constexpr uint32_t ROW_B = DH * 2, PLANE_B = QT * ROW_B;
    constexpr int CPR = DH / 8;                 // 16-byte chunks per row
    constexpr int RPI = 32 / CPR;               // rows per warp instruction
    const int c = lane % CPR, rsub = lane / CPR;
#pragma unroll 2
    for (int r = warp * RPI + rsub; r < rows_valid; r += SW * RPI) {
        const uint32_t off = (uint32_t)r * ROW_B + (uint32_t)((c ^ (r & 7)) << 4);
        __nv_bfloat16* dst = out.ptr + (grow0 + r) * out.ld + head * DH + c * 8;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stage + pl * PLANE_B + off));
            *reinterpret_cast<uint4*>(dst + pl * out.plane_stride) = v;
        }
    }
}

// additive mask values of N consecutive keys, -inf past Tk (branch-free: clamped address, predicated select), issued
// before the thread waits for the scores so that their latency is hidden behind the MMA
template <int N>
YV_DEVINL void load_mask(const float* mrow, int key0, int Tk, float* mk) {
    if (mrow) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int key = key0 + i;
            const float v = __ldg(mrow + min(key, Tk - 1));
            mk[i] = key < Tk ? v : -INFINITY;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) mk[i] = (key0 + i) < Tk ? 0.f : -INFINITY;
    }
}

template <int DH, int PASSES>
struct FwdCfg {
    static constexpr int PL = PASSES == 3 ? 2 : 1;
    static constexpr int DB = DH / 64;
    static constexpr uint32_t HEADER = 8192;                       // barriers, TMEM slot, row-statistic exchange
    static constexpr uint32_t SQ = 0;
    static constexpr uint32_t K_BUF = PL * DB * KV_BLK;             // one K chunk (all planes / head-dimension blocks)
    static constexpr uint32_t SK = SQ + PL * DB * Q_BLK;            // two K buffers
    static constexpr uint32_t SV = SK + 2 * K_BUF;
    static constexpr uint32_t SP = SV + PL * DB * KV_BLK;
    static constexpr uint32_t TILES = SP + PL * Q_BLK;
    static constexpr uint32_t SMEM = HEADER + 1024 + TILES;
    static constexpr int TMEM_COLS = 256;                           // S0 [0,64) S1 [64,128) O [128, 128 + DH)
    static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
    static_assert(TILES >= 2u * QT * DH * 2, "the operand tiles double as the output staging area");
};

// ------------------------------------------------------------------------------------------- forward
template <int DH, int PASSES, int SW>
__global__ void __launch_bounds__(att_threads(SW), 1)
yv_attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
    using C = FwdCfg<DH, PASSES>;
    constexpr int PL = C::PL, DB = C::DB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // barriers: 0 Q landed, 1-2 K[2] landed, 3 V landed, 4-5 S[2] retired, 6-7 S[2] read by the softmax warps,
    //           8 P written (+ O rescaled), 9 P V retired
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 96);
    constexpr int CG = SW / 4;                                      // column groups per chunk
    constexpr int KPT = KC / CG;                                    // keys per thread and chunk (32 or 16)
    constexpr int OW = DH / CG;                                     // accumulator columns per softmax warp
    constexpr int OG = OW < 32 ? OW : 32;                           // ... handled OG at a time
    float* xm = reinterpret_cast<float*>(smem_raw + 128);           // [2 (chunk parity)][CG][128 rows]
    float* xl = xm + 2 * CG * QT;                                   // [CG][128]
    static_assert(128 + (3 * CG * QT) * 4 <= C::HEADER, "row-statistic exchange does not fit the header");
    const uint32_t tiles = (smem_u32(smem_raw) + C::HEADER + 1023u) & ~1023u;
    const uint32_t sQ = tiles + C::SQ, sK = tiles + C::SK, sV = tiles + C::SV, sP = tiles + C::SP;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * i; };
    constexpr int B_Q = 0, B_K = 1, B_V = 3, B_S = 4, B_SFREE = 6, B_P = 8, B_O = 9;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, head = blockIdx.y, pair = blockIdx.z;
    const int q0 = qtile * QT;
    const int nchunks = (p.Tk + KC - 1) / KC;
    yv_pdl_trigger();

    if (threadIdx.x == 0) {
        mbar_init(bar(B_Q), 1);
        mbar_init(bar(B_K), 1);
        mbar_init(bar(B_K + 1), 1);
        mbar_init(bar(B_V), 1);
        mbar_init(bar(B_S), 1);
        mbar_init(bar(B_S + 1), 1);
        mbar_init(bar(B_SFREE), SW);
        mbar_init(bar(B_SFREE + 1), SW);
        mbar_init(bar(B_P), SW);
        mbar_init(bar(B_O), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_v) : "memory");
    }
    if (warp == SW) {
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

    auto load_kv = [&](const CUtensorMap* map, uint32_t dst, uint32_t b, int chunk) {
        mbar_expect_tx(b, PL * DB * KV_BLK);
#pragma unroll
        for (int pl = 0; pl < PL; ++pl)
#pragma unroll
            for (int d = 0; d < DB; ++d)
                tma_load_5d(dst + (pl * DB + d) * KV_BLK, map, b, d * 64, chunk * KC, head, pair, pl);
    };

    if (warp == SW) {
        if (lane == 0) {
            // ============================ Q / K loads, S = Q K^T (runs up to two chunks ahead) ============================
            mbar_expect_tx(bar(B_Q), PL * DB * Q_BLK);
#pragma unroll
            for (int pl = 0; pl < PL; ++pl)
#pragma unroll
                for (int d = 0; d < DB; ++d)
                    tma_load_5d(sQ + (pl * DB + d) * Q_BLK, &map_q, bar(B_Q), d * 64, q0, head, pair, pl);
            load_kv(&map_k, sK, bar(B_K), 0);
            if (nchunks > 1) load_kv(&map_k, sK + C::K_BUF, bar(B_K + 1), 1);
            YV_AT(0);
            mbar_wait(bar(B_Q), 0);
            for (int c = 0; c < nchunks; ++c) {
                const int b = c & 1;
                const uint32_t use = (uint32_t)((c >> 1) & 1);              // parity of this buffer's (c / 2)-th use
                if (c >= 1 && c + 1 < nchunks) {
                    // prefetch chunk c + 1 into the other buffer as soon as S_{c-1} (its last reader) has retired, i.e.
                    // BEFORE spending ~1.5k cycles on issuing S_c: the TMA latency hides behind the issue loop
                    mbar_wait(bar(B_S + (b ^ 1)), (uint32_t)(((c - 1) >> 1) & 1));
                    load_kv(&map_k, sK + (b ^ 1) * C::K_BUF, bar(B_K + (b ^ 1)), c + 1);
                }
                mbar_wait(bar(B_K + b), use);
                if (c >= 2) mbar_wait(bar(B_SFREE + b), use ^ 1u);          // softmax has read S_{c-2} out of this buffer
                YV_AT(8 + 8 * c + 0);
                tc_fence_after();
                mma_rows_x_rows<DH, PASSES>(tmem + (uint32_t)(b * KC), sQ, Q_BLK, sK + b * C::K_BUF, KV_BLK);
                umma_commit(bar(B_S + b));
                YV_AT(8 + 8 * c + 1);
            }
        }
    } else if (warp == SW + 1) {
        if (lane == 0) {
            // ======================================= V loads, O += P V =======================================
            load_kv(&map_v, sV, bar(B_V), 0);
            const uint32_t idesc_pv = make_idesc(QT, DH, 0, 1);
            for (int j = 0; j < nchunks; ++j) {
                const uint32_t ph = (uint32_t)(j & 1);
                mbar_wait(bar(B_P), ph);                                     // P_j in shared memory, O rescaled
                YV_AT(8 + 8 * j + 2);
                mbar_wait(bar(B_V), ph);                                     // V_j landed
                tc_fence_after();
                const int kc = min(KC, p.Tk - j * KC);
                const int ksteps = (kc + 15) >> 4;
                uint32_t accum = j > 0 ? 1u : 0u;
                for (int s = 0; s < ksteps; ++s)
                    mma_terms<PASSES>(tmem_o, desc_k(sP, s), desc_k(sP + Q_BLK, s), desc_mn(sV, KV_BLK, s),
                                      desc_mn(sV + DB * KV_BLK, KV_BLK, s), idesc_pv, accum);
                umma_commit(bar(B_O));
                if (j + 1 < nchunks) {
                    mbar_wait(bar(B_O), ph);                                 // P V retired: V and P buffers are free
                    YV_AT(8 + 8 * j + 3);
                    load_kv(&map_v, sV, bar(B_V), j + 1);
                }
            }
        }
    } else {
        // ========================================= softmax ============================================
        const int quarter = warp & 3, cg = warp >> 2;
        const int row = quarter * 32 + lane;                  // row of the tile = TMEM lane
        const int qrow = q0 + row;
        const bool row_ok = qrow < p.Tq;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t drop_row = (uint32_t)(((long long)(pair * p.heads + head) * p.Tq + qrow) * p.Tk);
        const float* mrow = p.mask ? p.mask + (long long)pair * p.Tk : nullptr;
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nchunks; ++j) {
            const int key0 = j * KC + cg * KPT;
            float x[KPT];
            load_mask<KPT>(mrow, key0, p.Tk, x);
            mbar_wait(bar(B_S + (j & 1)), (uint32_t)((j >> 1) & 1));
            if (threadIdx.x == 0) YV_AT(8 + 8 * j + 4);
            tc_fence_after();
            uint32_t raw[KPT];
            tmem_ld<KPT>(tmem + lane_addr + (uint32_t)((j & 1) * KC + cg * KPT), raw);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_SFREE + (j & 1)));       // this S buffer may take chunk j + 2
            if (threadIdx.x == 0) YV_AT(64 + 8 * j + 0);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                x[i] = fmaf(__uint_as_float(raw[i]), p.scale, x[i]);      // -inf for keys past Tk
                mx = fmaxf(mx, x[i]);
            }
            float* xmj = xm + (j & 1) * CG * QT;
            xmj[cg * QT + row] = mx;
            if (threadIdx.x == 0) YV_AT(64 + 8 * j + 1);
            softmax_bar<SW>();
#pragma unroll
            for (int g = 0; g < CG; ++g) mx = fmaxf(mx, xmj[g * QT + row]);
            if (threadIdx.x == 0) YV_AT(8 + 8 * j + 5);
            // online softmax with a lazy reference maximum: the accumulator is rescaled only when the row maximum
            // grew by more than RESCALE_THRESHOLD (all column groups take the same decision from the same numbers)
            const float m_new = (j == 0 || mx > m_ref + RESCALE_THRESHOLD) ? mx : m_ref;
            const float alpha = (j == 0) ? 1.f : __expf(m_ref - m_new);
            m_ref = m_new;
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                x[i] = __expf(x[i] - m_new);                  // exp(-inf) = 0 for keys past Tk
                sum += x[i];
            }
            l_run = l_run * alpha + sum;
            if (drop.thresh) {
#pragma unroll
                for (int i = 0; i < KPT; ++i) x[i] *= yv_drop_mul(drop, drop_row + (uint32_t)(key0 + i));
            }
            if (threadIdx.x == 0) YV_AT(8 + 8 * j + 6);
            if (j > 0) {
                mbar_wait(bar(B_O), (uint32_t)((j - 1) & 1)); // P V of the previous chunk retired
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
                    for (int g = 0; g < OW / OG; ++g) {
                        uint32_t o[OG];
                        const uint32_t a = tmem_o + lane_addr + (uint32_t)(cg * OW + g * OG);
                        tmem_ld<OG>(a, o);
#pragma unroll
                        for (int i = 0; i < OG; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st<OG>(a, o);
                    }
                }
            }
            if (threadIdx.x == 0) YV_AT(64 + 8 * j + 2);
            store_tile_row<PASSES, KPT>(sP, row, cg, x);
            if (threadIdx.x == 0) YV_AT(64 + 8 * j + 3);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_P));
            if (threadIdx.x == 0) YV_AT(8 + 8 * j + 7);
        }
        // ---- epilogue: O / l -> planes (+ fp32), log-sum-exp for the backward.  Every MMA has retired and no TMA load
        // is in flight: the operand tiles are free and serve as the output staging area.
        mbar_wait(bar(B_O), (uint32_t)((nchunks - 1) & 1));
        if (threadIdx.x == 0) YV_AT(2);
        tc_fence_after();
        xl[cg * QT + row] = l_run;
        softmax_bar<SW>();
        float l_tot = 0.f;
#pragma unroll
        for (int g = 0; g < CG; ++g) l_tot += xl[g * QT + row];
        const float inv = 1.f / l_tot;
        const long long grow = (long long)pair * p.Tq + qrow;
#pragma unroll 1
        for (int g = 0; g < OW / OG; ++g) {
            uint32_t o[OG];
            const int col = cg * OW + g * OG;
            tmem_ld<OG>(tmem_o + lane_addr + (uint32_t)col, o);
            float y[OG];
#pragma unroll
            for (int i = 0; i < OG; ++i) y[i] = __uint_as_float(o[i]) * inv;
            stage_row<DH, OG>(tiles, row, col, y);
            if (p.o32 && row_ok) {
                float* dst32 = p.o32 + grow * p.o32_ld + head * DH + col;
#pragma unroll
                for (int i = 0; i < OG / 4; ++i)
                    *reinterpret_cast<float4*>(dst32 + 4 * i) = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
            }
        }
        if (cg == 0 && row_ok && p.lse) p.lse[(long long)(pair * p.heads + head) * p.Tq + qrow] = m_ref + logf(l_tot);
        softmax_bar<SW>();
        copy_out_rows<DH, SW>(tiles, p.o, (long long)pair * p.Tq + q0, min(QT, p.Tq - q0), head, warp, lane);
        if (threadIdx.x == 0) YV_AT(3);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == SW)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(C::TMEM_COLS) : "memory");
}

// dK / dV planes of one (pair, head) = sum over the query tiles' slabs.  NS > 0: slab count known at compile time, all
// NS * UNR loads of a batch are issued before the first add (the last CTA of a (pair, head) runs this alone).
template <int DH, int NS, int SW>
YV_DEVINL void sum_slabs(const AttnParams& p, int pair, int head, int nslabs) {
    constexpr int SM_THREADS = SW * 32;
    constexpr int V4 = DH / 4;
    constexpr int UNR = NS == 1 ? 8 : 4;
    const long long slab_stride = (long long)p.pairs * p.Tk * p.dkv_ld;
    const int total = p.Tk * V4;
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
        const int col = t ? p.dv_col : p.dk_col;
        const PlaneView& out = t ? p.dv : p.dk;
#pragma unroll 1
        for (int base = threadIdx.x; base < total; base += SM_THREADS * UNR) {
            float4 acc[UNR];
            if (NS > 0) {
                float4 w[UNR][NS > 0 ? NS : 1];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = min(base + u * SM_THREADS, total - 1);     // clamped: every load is unconditional
                    const int key = idx / V4, c = (idx % V4) * 4;
                    const float* src = p.dkv32 + ((long long)pair * p.Tk + key) * p.dkv_ld + col + head * DH + c;
#pragma unroll
                    for (int sl = 0; sl < (NS > 0 ? NS : 1); ++sl)
                        w[u][sl] = __ldcg(reinterpret_cast<const float4*>(src + sl * slab_stride));
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    acc[u] = w[u][0];
#pragma unroll
                    for (int sl = 1; sl < (NS > 0 ? NS : 1); ++sl) {
                        acc[u].x += w[u][sl].x; acc[u].y += w[u][sl].y; acc[u].z += w[u][sl].z; acc[u].w += w[u][sl].w;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int idx = min(base + u * SM_THREADS, total - 1);
                    const int key = idx / V4, c = (idx % V4) * 4;
                    const float* src = p.dkv32 + ((long long)pair * p.Tk + key) * p.dkv_ld + col + head * DH + c;
                    acc[u] = __ldcg(reinterpret_cast<const float4*>(src));
                    for (int sl = 1; sl < nslabs; ++sl) {
                        const float4 x = __ldcg(reinterpret_cast<const float4*>(src + sl * slab_stride));
                        acc[u].x += x.x; acc[u].y += x.y; acc[u].z += x.z; acc[u].w += x.w;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int idx = base + u * SM_THREADS;
                if (idx < total) {
                    const int key = idx / V4, c = (idx % V4) * 4;
                    uint32_t h0, l0, h1, l1;
                    yv_split2(acc[u].x, acc[u].y, h0, l0);
                    yv_split2(acc[u].z, acc[u].w, h1, l1);
                    __nv_bfloat16* dst = out.ptr + ((long long)pair * p.Tk + key) * out.ld + head * DH + c;
                    *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                    *reinterpret_cast<uint2*>(dst + out.plane_stride) = make_uint2(l0, l1);
                }
            }
        }
    }
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
    static constexpr uint32_t ST = SV + PL * DB * KV_BLK;            // Pd tile
    // dS tile: V is dead once dPd = dO V^T has retired, so with 128-wide heads (where shared memory is full) the dS
    // tile lives in the V buffer; 64-wide heads have room for a tile of their own
    static constexpr uint32_t SD = DH == 128 ? SV : ST + PL * Q_BLK;
    static constexpr uint32_t TILES = (DH == 128 ? ST : SD) + PL * Q_BLK;
    static constexpr uint32_t SMEM = HEADER + 1024 + TILES;
    static constexpr int TMEM_COLS = 512;   // S [0,64) dPd [64,128) dV^T [128,192) dK^T [192,256) dQ [256, 256 + DH)
    static_assert(DH != 128 || PL * DB * KV_BLK == PL * Q_BLK, "dS tile must fit the V buffer");
    static_assert(SMEM <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
    static_assert(TILES >= 2u * QT * DH * 2, "the operand tiles double as the output staging area");
};

// delta[r] = sum_d dO[r, d] * O[r, d] (the softmax-backward row term) for the 128 rows of a query tile, from the global
// plane pairs: a warp reads whole rows (8 bytes per lane and plane, coalesced), four rows in flight, shuffle reduction
template <int DH, int SW>
YV_DEVINL void row_dots(const PlaneView& a, const PlaneView& b, long long grow0, int rows_valid, int head, int warp, int lane,
                        float* delta) {
    constexpr int LPR = DH / 4;                 // lanes per row (4 bf16 = 8 bytes each)
    constexpr int RPI = 32 / LPR;               // rows per warp instruction (1 for DH = 128, 2 for DH = 64)
    const int c = (lane % LPR) * 4, rsub = lane / LPR;
    for (int r0 = warp * RPI * 4; r0 < QT; r0 += SW * RPI * 4) {
        uint2 ah[4], al[4], bh[4], bl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * RPI + rsub;
            ah[u] = al[u] = bh[u] = bl[u] = make_uint2(0u, 0u);
            if (r < rows_valid) {
                const __nv_bfloat16* pa = a.ptr + (grow0 + r) * a.ld + head * DH + c;
                const __nv_bfloat16* pb = b.ptr + (grow0 + r) * b.ld + head * DH + c;
                ah[u] = __ldg(reinterpret_cast<const uint2*>(pa));
                al[u] = __ldg(reinterpret_cast<const uint2*>(pa + a.plane_stride));
                bh[u] = __ldg(reinterpret_cast<const uint2*>(pb));
                bl[u] = __ldg(reinterpret_cast<const uint2*>(pb + b.plane_stride));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // bf16 -> fp32 is a 16-bit shift
            const float a0 = __uint_as_float(ah[u].x << 16) + __uint_as_float(al[u].x << 16);
            const float a1 = __uint_as_float(ah[u].x & 0xffff0000u) + __uint_as_float(al[u].x & 0xffff0000u);
            const float a2 = __uint_as_float(ah[u].y << 16) + __uint_as_float(al[u].y << 16);
            const float a3 = __uint_as_float(ah[u].y & 0xffff0000u) + __uint_as_float(al[u].y & 0xffff0000u);
            const float b0 = __uint_as_float(bh[u].x << 16) + __uint_as_float(bl[u].x << 16);
            const float b1 = __uint_as_float(bh[u].x & 0xffff0000u) + __uint_as_float(bl[u].x & 0xffff0000u);
            const float b2 = __uint_as_float(bh[u].y << 16) + __uint_as_float(bl[u].y << 16);
            const float b3 = __uint_as_float(bh[u].y & 0xffff0000u) + __uint_as_float(bl[u].y & 0xffff0000u);
            float acc = fmaf(a0, b0, fmaf(a1, b1, fmaf(a2, b2, a3 * b3)));
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if ((lane % LPR) == 0) delta[r0 + u * RPI + rsub] = acc;
        }
    }
}

template <int DH, int PASSES, int SW>
__global__ void __launch_bounds__(att_threads(SW), 1)
yv_attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ AttnParams p) {
    using C = BwdCfg<DH, PASSES>;
    constexpr int PL = C::PL, DB = C::DB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // barriers: 0 Q+dO landed, 1 K+V landed, 2 S retired, 3 dPd retired, 4 Pd+dS tiles written, 5 dV^T retired,
    //           6 dK^T+dQ retired
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 64);
    uint32_t* last_flag = reinterpret_cast<uint32_t*>(smem_raw + 68);
    float* delta_s = reinterpret_cast<float*>(smem_raw + 256);      // [128] row terms dO . O
    constexpr int B_QDO = 0, B_KV = 1, B_S = 2, B_DP = 3, B_T = 4, B_DV = 5, B_DKQ = 6;
    constexpr int CG = SW / 4;                                      // column groups per chunk
    constexpr int KPT = KC / CG;                                    // keys per thread and chunk (32 or 16)
    constexpr int OW = DH / CG;                                     // dQ columns per softmax warp
    constexpr int OG = OW < 32 ? OW : 32;
    const uint32_t tiles = (smem_u32(smem_raw) + C::HEADER + 1023u) & ~1023u;
    const uint32_t sQ = tiles + C::SQ, sDO = tiles + C::SDO, sK = tiles + C::SK, sV = tiles + C::SV, sT = tiles + C::ST,
                   sD = tiles + C::SD;
    const uint32_t bar0 = smem_u32(bars);
    auto bar = [&](int i) { return bar0 + 8u * i; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, head = blockIdx.y, pair = blockIdx.z;
    const int q0 = qtile * QT;
    const int nchunks = (p.Tk + KC - 1) / KC;
    yv_pdl_trigger();

    if (threadIdx.x == 0) {
        mbar_init(bar(B_QDO), 1);
        mbar_init(bar(B_KV), 1);
        mbar_init(bar(B_S), 1);
        mbar_init(bar(B_DP), 1);
        mbar_init(bar(B_T), SW);
        mbar_init(bar(B_DV), 1);
        mbar_init(bar(B_DKQ), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_v) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)&map_do) : "memory");
    }
    if (warp == SW) {
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

    const uint32_t idesc_t = make_idesc(128, KC, 1, 1);          // dV^T / dK^T: both operands MN-major
    if (warp == SW) {
        if (lane == 0) {
            // =========================== loads, S = Q K^T, dV^T = dO^T Pd ===========================
            auto load_rows = [&](const CUtensorMap* map, uint32_t dst, uint32_t blk, uint32_t b, int row0) {
#pragma unroll
                for (int pl = 0; pl < PL; ++pl)
#pragma unroll
                    for (int d = 0; d < DB; ++d)
                        tma_load_5d(dst + (pl * DB + d) * blk, map, b, d * 64, row0, head, pair, pl);
            };
            YV_AT(0);
            mbar_expect_tx(bar(B_QDO), 2 * PL * DB * Q_BLK);
            load_rows(&map_q, sQ, Q_BLK, bar(B_QDO), q0);
            load_rows(&map_do, sDO, Q_BLK, bar(B_QDO), q0);
            mbar_expect_tx(bar(B_KV), 2 * PL * DB * KV_BLK);
            load_rows(&map_k, sK, KV_BLK, bar(B_KV), 0);
            load_rows(&map_v, sV, KV_BLK, bar(B_KV), 0);
            mbar_wait(bar(B_QDO), 0);
            for (int j = 0; j < nchunks; ++j) {
                const uint32_t ph = (uint32_t)(j & 1);
                mbar_wait(bar(B_KV), ph);
                YV_AT(8 + 16 * j + 0);
                tc_fence_after();
                mma_rows_x_rows<DH, PASSES>(tm_s, sQ, Q_BLK, sK, KV_BLK);
                umma_commit(bar(B_S));
                YV_AT(8 + 16 * j + 1);
                mbar_wait(bar(B_T), ph);                                        // Pd and dS tiles written
                YV_AT(8 + 16 * j + 2);
                tc_fence_after();
                {   // dV^T [DH x 64] = dO^T . Pd : contraction over the 128 query rows
                    uint32_t accum = 0;
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mma_terms<PASSES>(tm_dv, desc_mn(sDO, Q_BLK, s), desc_mn(sDO + DB * Q_BLK, Q_BLK, s),
                                          desc_mn(sT, Q_BLK, s), desc_mn(sT + Q_BLK, Q_BLK, s), idesc_t, accum);
                }
                umma_commit(bar(B_DV));
                YV_AT(8 + 16 * j + 3);
                if (j + 1 < nchunks) {
                    mbar_wait(bar(B_DV), ph);                                   // Pd tile consumed
                    mbar_wait(bar(B_DKQ), ph);                                  // K and V (= dS tile) consumed
                    YV_AT(8 + 16 * j + 4);
                    mbar_expect_tx(bar(B_KV), 2 * PL * DB * KV_BLK);
                    load_rows(&map_k, sK, KV_BLK, bar(B_KV), (j + 1) * KC);
                    load_rows(&map_v, sV, KV_BLK, bar(B_KV), (j + 1) * KC);
                }
            }
        }
    } else if (warp == SW + 1) {
        if (lane == 0) {
            // =========================== dPd = dO V^T, dK^T = Q^T dS, dQ += dS K ===========================
            const uint32_t idesc_dq = make_idesc(QT, DH, 0, 1);
            mbar_wait(bar(B_QDO), 0);
            for (int j = 0; j < nchunks; ++j) {
                const uint32_t ph = (uint32_t)(j & 1);
                mbar_wait(bar(B_KV), ph);
                tc_fence_after();
                mma_rows_x_rows<DH, PASSES>(tm_dp, sDO, Q_BLK, sV, KV_BLK);
                umma_commit(bar(B_DP));
                YV_AT(8 + 16 * j + 5);
                mbar_wait(bar(B_T), ph);
                tc_fence_after();
                {   // dK^T [DH x 64] = Q^T . dS
                    uint32_t accum = 0;
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mma_terms<PASSES>(tm_dk, desc_mn(sQ, Q_BLK, s), desc_mn(sQ + DB * Q_BLK, Q_BLK, s),
                                          desc_mn(sD, Q_BLK, s), desc_mn(sD + Q_BLK, Q_BLK, s), idesc_t, accum);
                }
                {   // dQ [128 x DH] += dS . K : contraction over the keys of this chunk
                    const int kc = min(KC, p.Tk - j * KC);
                    const int ksteps = (kc + 15) >> 4;
                    uint32_t accum = j > 0 ? 1u : 0u;
                    for (int s = 0; s < ksteps; ++s)
                        mma_terms<PASSES>(tm_dq, desc_k(sD, s), desc_k(sD + Q_BLK, s), desc_mn(sK, KV_BLK, s),
                                          desc_mn(sK + DB * KV_BLK, KV_BLK, s), idesc_dq, accum);
                }
                umma_commit(bar(B_DKQ));
                YV_AT(8 + 16 * j + 6);
            }
        }
    } else {
        // ========================================= softmax ============================================
        const int quarter = warp & 3, cg = warp >> 2;
        const int row = quarter * 32 + lane;
        const int qrow = q0 + row;
        const bool row_ok = qrow < p.Tq;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        const YvDrop drop = yv_drop_make(p.rng, p.drop_site, p.drop_p);
        const uint32_t drop_row = (uint32_t)(((long long)(pair * p.heads + head) * p.Tq + qrow) * p.Tk);
        const float* mrow = p.mask ? p.mask + (long long)pair * p.Tk : nullptr;
        // rows past Tq (zero-filled by TMA) get lse = +inf: their probabilities and dS are exactly zero
        row_dots<DH, SW>(p.d_o, p.fwd_o, (long long)pair * p.Tq + q0, min(QT, p.Tq - q0), head, warp, lane, delta_s);
        const float lse = row_ok ? p.lse[(long long)(pair * p.heads + head) * p.Tq + qrow] : INFINITY;
        softmax_bar<SW>();
        const float delta = delta_s[row];
        // partial dK / dV of this query tile: slab `qtile` of the workspace, lane = head dimension index
        const int d_lane = quarter * 32 + lane;
        float* slab = p.dkv32 + ((long long)qtile * p.pairs + pair) * p.Tk * p.dkv_ld + head * DH + d_lane;
        for (int j = 0; j < nchunks; ++j) {
            const uint32_t ph = (uint32_t)(j & 1);
            const int key0 = j * KC + cg * KPT;
            float pd[KPT], ds[KPT];
            load_mask<KPT>(mrow, key0, p.Tk, pd);
            mbar_wait(bar(B_S), ph);
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 8);
            tc_fence_after();
            {
                uint32_t raw[KPT];
                tmem_ld<KPT>(tm_s + lane_addr + (uint32_t)(cg * KPT), raw);
#pragma unroll
                for (int i = 0; i < KPT; ++i)
                    pd[i] = __expf(fmaf(__uint_as_float(raw[i]), p.scale, pd[i]) - lse);   // 0 past Tk / past Tq
                mbar_wait(bar(B_DP), ph);
                tc_fence_after();
                tmem_ld<KPT>(tm_dp + lane_addr + (uint32_t)(cg * KPT), raw);
#pragma unroll
                for (int i = 0; i < KPT; ++i) {
                    const float mult = drop.thresh ? yv_drop_mul(drop, drop_row + (uint32_t)(key0 + i)) : 1.f;
                    const float pr = pd[i];
                    ds[i] = p.scale * pr * (__uint_as_float(raw[i]) * mult - delta);
                    pd[i] = pr * mult;
                }
            }
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 9);
            // (the products that read the previous chunk's tiles retired before this thread drained their results)
            store_tile_row<PASSES, KPT>(sT, row, cg, pd);
            store_tile_row<PASSES, KPT>(sD, row, cg, ds);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_T));
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 10);
            // drain dV^T and dK^T (lane = d, column = key): 128-byte coalesced rows of the slab
            const int nkeys = min(KPT, p.Tk - key0);              // warp-uniform
            mbar_wait(bar(B_DV), ph);
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 11);
            tc_fence_after();
            if (DH == 128 || quarter < 2) {
                uint32_t raw[KPT];
                tmem_ld<KPT>(tm_dv + lane_addr + (uint32_t)(cg * KPT), raw);
                float* dst = slab + (long long)key0 * p.dkv_ld + p.dv_col;
                if (nkeys >= KPT) {
#pragma unroll
                    for (int i = 0; i < KPT; ++i) dst[(long long)i * p.dkv_ld] = __uint_as_float(raw[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < KPT; ++i)
                        if (i < nkeys) dst[(long long)i * p.dkv_ld] = __uint_as_float(raw[i]);
                }
            }
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 12);
            mbar_wait(bar(B_DKQ), ph);
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 13);
            tc_fence_after();
            if (DH == 128 || quarter < 2) {
                uint32_t raw[KPT];
                tmem_ld<KPT>(tm_dk + lane_addr + (uint32_t)(cg * KPT), raw);
                float* dst = slab + (long long)key0 * p.dkv_ld + p.dk_col;
                if (nkeys >= KPT) {
#pragma unroll
                    for (int i = 0; i < KPT; ++i) dst[(long long)i * p.dkv_ld] = __uint_as_float(raw[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < KPT; ++i)
                        if (i < nkeys) dst[(long long)i * p.dkv_ld] = __uint_as_float(raw[i]);
                }
            }
            tc_fence_before();
            if (threadIdx.x == 0) YV_AT(8 + 16 * j + 14);
        }
        if (threadIdx.x == 0) YV_AT(1);
        // ---- dQ tile -> planes through the (now free) operand tiles
#pragma unroll 1
        for (int g = 0; g < OW / OG; ++g) {
            uint32_t o[OG];
            const int col = cg * OW + g * OG;
            tmem_ld<OG>(tm_dq + lane_addr + (uint32_t)col, o);
            float y[OG];
#pragma unroll
            for (int i = 0; i < OG; ++i) y[i] = __uint_as_float(o[i]);
            stage_row<DH, OG>(tiles, row, col, y);
        }
        __threadfence();                                      // this tile's slab is visible before the ticket is taken
        softmax_bar<SW>();
        if (threadIdx.x == 0) YV_AT(2);
        copy_out_rows<DH, SW>(tiles, p.dq, (long long)pair * p.Tq + q0, min(QT, p.Tq - q0), head, warp, lane);
        // ---- the last query tile of this (pair, head) adds the slabs and writes the dK / dV planes
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(p.tickets + pair * p.heads + head, 1u);
            const bool last = t == gridDim.x - 1;
            if (last) p.tickets[pair * p.heads + head] = 0u;  // self-resetting: the buffer can be reused by the next launch
            *last_flag = last ? 1u : 0u;
        }
        softmax_bar<SW>();
        if (threadIdx.x == 0) YV_AT(3);
        if (*last_flag) {
            __threadfence();
            const int nslabs = (int)gridDim.x;
            if (nslabs == 1) sum_slabs<DH, 1, SW>(p, pair, head, 1);
            else if (nslabs == 2) sum_slabs<DH, 2, SW>(p, pair, head, 2);
            else if (nslabs == 3) sum_slabs<DH, 3, SW>(p, pair, head, 3);
            else sum_slabs<DH, 0, SW>(p, pair, head, nslabs);
        }
    }

    if (threadIdx.x == 0) YV_AT(4);
    tc_fence_before();
    __syncthreads();
    if (warp == SW)
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

// per-device one-time opt-in to > 48 KB of dynamic shared memory (cudaFuncSetAttribute is a per-device setting and the
// reference's nn.DataParallel fallback calls forward from one host thread per device)
template <typename K>
int set_smem_once(K kernel, int bytes, bool (&done)[64], std::mutex& mu) {
    int dev = 0;
    YV_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev & 63]) return 0;
    YV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done[dev & 63] = true;
    return 0;
}

// softmax warps per CTA: 16 by default, YVB200_ATTN_WARPS=8 selects the 8-warp build (A/B timing)
const int g_attn_warps = []() { const char* e = getenv("YVB200_ATTN_WARPS"); return (e && atoi(e) == 8) ? 8 : 16; }();

template <int DH, int PASSES, int SW>
int launch_fwd(const YvAttnFwd* a, const AttnParams& p, cudaStream_t st) {
    using C = FwdCfg<DH, PASSES>;
    CUtensorMap mq, mk, mv;
    if (view_map(&mq, a->q, a->pairs, a->heads, DH, PASSES, QT, "Q")) return 1;
    if (view_map(&mk, a->k, a->pairs, a->heads, DH, PASSES, KC, "K")) return 1;
    if (view_map(&mv, a->v, a->pairs, a->heads, DH, PASSES, KC, "V")) return 1;
    static bool done[64] = {};
    static std::mutex mu;
    if (set_smem_once(yv_attn_fwd_kernel<DH, PASSES, SW>, (int)C::SMEM, done, mu)) return 2;
    dim3 grid((unsigned)((p.Tq + QT - 1) / QT), (unsigned)a->heads, (unsigned)a->pairs);
    YV_CUDA(yv_launch(yv_attn_fwd_kernel<DH, PASSES, SW>, grid, dim3(att_threads(SW)), C::SMEM, st, mq, mk, mv, p));
    return 0;
}

template <int DH, int PASSES, int SW>
int launch_bwd(const YvAttnBwd* a, const AttnParams& p, cudaStream_t st) {
    using C = BwdCfg<DH, PASSES>;
    CUtensorMap mq, mk, mv, md;
    if (view_map(&mq, a->q, a->pairs, a->heads, DH, PASSES, QT, "Q")) return 1;
    if (view_map(&mk, a->k, a->pairs, a->heads, DH, PASSES, KC, "K")) return 1;
    if (view_map(&mv, a->v, a->pairs, a->heads, DH, PASSES, KC, "V")) return 1;
    if (view_map(&md, a->dout, a->pairs, a->heads, DH, PASSES, QT, "dO")) return 1;
    static bool done[64] = {};
    static std::mutex mu;
    if (set_smem_once(yv_attn_bwd_kernel<DH, PASSES, SW>, (int)C::SMEM, done, mu)) return 2;
    dim3 grid((unsigned)((p.Tq + QT - 1) / QT), (unsigned)a->heads, (unsigned)a->pairs);
    YV_CUDA(yv_launch(yv_attn_bwd_kernel<DH, PASSES, SW>, grid, dim3(att_threads(SW)), C::SMEM, st, mq, mk, mv, md, p));
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
    if (g_attn_warps == 16) {
        if (a->dh == 128) rc = a->passes == 3 ? launch_fwd<128, 3, 16>(a, p, st) : launch_fwd<128, 1, 16>(a, p, st);
        else rc = a->passes == 3 ? launch_fwd<64, 3, 16>(a, p, st) : launch_fwd<64, 1, 16>(a, p, st);
    } else {
        if (a->dh == 128) rc = a->passes == 3 ? launch_fwd<128, 3, 8>(a, p, st) : launch_fwd<128, 1, 8>(a, p, st);
        else rc = a->passes == 3 ? launch_fwd<64, 3, 8>(a, p, st) : launch_fwd<64, 1, 8>(a, p, st);
    }
    if (rc) return rc;
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}

extern "C" size_t yv_attn_bwd_workspace_bytes(int32_t pairs, int32_t heads, int32_t dh, int32_t Tq, int32_t Tk) {
    if (pairs <= 0 || heads <= 0 || dh <= 0 || Tq <= 0 || Tk <= 0) return 0;
    // one fp32 slab [pairs * Tk, 2 * heads * dh] (dK | dV) per 128-query tile
    return (size_t)((Tq + QT - 1) / QT) * pairs * Tk * 2 * heads * dh * sizeof(float);
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
    const size_t need = yv_attn_bwd_workspace_bytes(a->pairs, a->heads, a->dh, a->q.rows, a->k.rows);
    YV_CHECK(a->workspace != nullptr && a->workspace_bytes >= need && ((uintptr_t)a->workspace & 15) == 0,
             "yv_attn_bwd: workspace of %zu bytes required (got %zu)", need, (size_t)a->workspace_bytes);
    YV_CHECK(a->tickets != nullptr, "yv_attn_bwd: tickets (pairs*heads zero-initialised uint32) required");
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
    p.tickets = a->tickets;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    if (g_attn_warps == 16) {
        if (a->dh == 128) rc = a->passes == 3 ? launch_bwd<128, 3, 16>(a, p, st) : launch_bwd<128, 1, 16>(a, p, st);
        else rc = a->passes == 3 ? launch_bwd<64, 3, 16>(a, p, st) : launch_bwd<64, 1, 16>(a, p, st);
    } else {
        if (a->dh == 128) rc = a->passes == 3 ? launch_bwd<128, 3, 8>(a, p, st) : launch_bwd<128, 1, 8>(a, p, st);
        else rc = a->passes == 3 ? launch_bwd<64, 3, 8>(a, p, st) : launch_bwd<64, 1, 8>(a, p, st);
    }
    if (rc) return rc;
    YV_CUDA(cudaGetLastError());
    yv_count_launch();
    return 0;
}
