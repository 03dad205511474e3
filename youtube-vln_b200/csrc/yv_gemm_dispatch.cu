// yv_gemm: picks the tcgen05 GEMM variant for a problem (both are built from yv_gemm.cu).
#include <stdlib.h>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

extern "C" int yv_gemm_k32(const YvGemm* g, yv_stream_t stream);
extern "C" int yv_gemm_k64(const YvGemm* g, yv_stream_t stream);

extern "C" int yv_gemm(const YvGemm* g, yv_stream_t stream) {
    YV_CHECK(g != nullptr, "yv_gemm: NULL args");
    static const int force = []() { const char* e = getenv("YVB200_GEMM_VARIANT"); return e ? atoi(e) : 0; }();
    if (force == 32) return yv_gemm_k32(g, stream);
    if (force == 64) return yv_gemm_k64(g, stream);
    // several tiles per SM: the persistent variant hides every epilogue behind the next tile's main loop;
    // otherwise the half-SM variant lets CTAs of concurrently running launches share an SM
    const long long tiles = (long long)((g->M + 127) / 128) * ((g->N + 127) / 128) * g->a.nb0 * g->a.nb1;
    return tiles >= 2 * 148 + 74 ? yv_gemm_k64(g, stream) : yv_gemm_k32(g, stream);
}
