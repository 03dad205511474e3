// yv_gemm: picks the tcgen05 GEMM variant for a problem.
//   yv_gemm_k32 / yv_gemm_k64 (yv_gemm.cu)   one CTA per 128x128 tile (cta_group::1)
//   yv_gemm_pair (yv_gemm_pair.cu)           two CTAs per 256 x {128,256} tile (cta_group::2)
#include <stdlib.h>

#include <atomic>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

extern "C" int yv_gemm_k32(const YvGemm* g, yv_stream_t stream);
extern "C" int yv_gemm_k64(const YvGemm* g, yv_stream_t stream);
extern "C" int yv_gemm_pair(const YvGemm* g, int pair_n, yv_stream_t stream);
extern "C" int yv_gemm_k32_splits(const YvGemm* g);
extern "C" int yv_gemm_k64_splits(const YvGemm* g);
extern "C" int yv_gemm_pair_splits(const YvGemm* g, int pair_n);

namespace {
// 0 = automatic; 32 / 64 = the single-CTA variants; 2 = CTA pairs (tile width chosen per problem);
// 128 / 256 = CTA pairs with that tile width.  Initial value from YVB200_GEMM_VARIANT.
std::atomic<int> g_variant{[]() { const char* e = getenv("YVB200_GEMM_VARIANT"); return e ? atoi(e) : 0; }()};
}  // namespace

extern "C" int yv_gemm_set_variant(int variant) {
    YV_CHECK(variant == 0 || variant == 32 || variant == 64 || variant == 2 || variant == 128 || variant == 256,
             "yv_gemm_set_variant: unknown variant %d", variant);
    g_variant.store(variant);
    return 0;
}

extern "C" int yv_gemm(const YvGemm* g, yv_stream_t stream) {
    YV_CHECK(g != nullptr, "yv_gemm: NULL args");
    const int force = g_variant.load(std::memory_order_relaxed);
    if (force == 32) return yv_gemm_k32(g, stream);
    if (force == 64) return yv_gemm_k64(g, stream);
    if (force == 2) return yv_gemm_pair(g, 0, stream);
    if (force == 128 || force == 256) return yv_gemm_pair(g, force, stream);
    // Measured on the cfg2 training step (B200, CUDA graph, bf16x3): every launch on the persistent 64-deep kernel
    // 10.70 ms, every launch on the half-SM 32-deep kernel 11.07 ms, CTA pairs wherever M allows 11.66 ms.  The
    // 64-deep ring keeps twice the operand bytes in flight per SM and halves the barrier round trips; its main loop
    // runs at ~80 % of the MMA issue rate against ~58 % for the 32-deep one (tools/gemm_timing.cu).  The CTA-pair
    // kernel reaches ~100 % in the main loop with 256-wide tiles but leaves half the SMs idle on the one-wave
    // problems of this step, so it stays opt-in (yv_gemm_set_variant / YVB200_GEMM_VARIANT=2).
    return yv_gemm_k64(g, stream);
}

// How many K splits yv_gemm would use for this problem (1 = none).  A split launch reduces partial sums into a
// zero-filled f32 output; a caller that zero-fills the output itself, off its critical path, sets g->out32_zeroed and
// saves the memset node in front of the kernel.
extern "C" int yv_gemm_splits(const YvGemm* g) {
    if (g == nullptr || g->M <= 0 || g->N <= 0 || g->K <= 0) return 1;
    const int force = g_variant.load(std::memory_order_relaxed);
    if (force == 32) return yv_gemm_k32_splits(g);
    if (force == 2) return yv_gemm_pair_splits(g, 0);
    if (force == 128 || force == 256) return yv_gemm_pair_splits(g, force);
    return yv_gemm_k64_splits(g);
}
