// Developer tool: per-phase clock64 timeline of CTA 0 of one yv_gemm variant.  Build (see tools/gemm_timing.sh):
//   nvcc -DYV_GEMM_TIMING -DYV_BLOCK_K=32|64 [-DYV_TIMING_PAIR=128|256] ... ; run on the GPU box.
#include "../youtube-vln_b200/csrc/yv_core.cu"
#ifdef YV_TIMING_PAIR
#include "../youtube-vln_b200/csrc/yv_gemm_pair.cu"
#define RUN(g) yv_gemm_pair(g, YV_TIMING_PAIR, 0)
#else
#include "../youtube-vln_b200/csrc/yv_gemm.cu"
#define RUN(g) YV_GEMM_ENTRY(g, 0)
#endif
#include <vector>
int main(int argc, char** argv) {
    int shapes[][4] = {{128, 128, 64, 3}, {128, 128, 4096, 3}, {256, 256, 4096, 3}, {2304, 1024, 1024, 3}, {2304, 1024, 1024, 1},
                       {2304, 3072, 1024, 3}, {640, 3072, 768, 3}, {640, 768, 768, 3}};
    for (auto& sh : shapes) {
        int M = sh[0], N = sh[1], K = sh[2], passes = sh[3];
        __nv_bfloat16 *a, *b, *pl; float* o;
        cudaMalloc(&a, 2ll * M * K * 2); cudaMalloc(&b, 2ll * N * K * 2); cudaMalloc(&o, 4ll * M * N); cudaMalloc(&pl, 2ll * M * N * 2);
        cudaMemset(a, 0, 2ll * M * K * 2); cudaMemset(b, 0, 2ll * N * K * 2);
        for (int mode = 0; mode < 2; ++mode) {           // 0: f32 output only, 1: f32 + planes
            YvGemm g; memset(&g, 0, sizeof(g));
            g.M = M; g.N = N; g.K = K; g.passes = passes; g.alpha = 1.f;
            g.a = {a, K, M, K, 1, 0, 1, 0, (int64_t)M * K, 0, 0};
            g.b = {b, K, N, K, 1, 0, 1, 0, (int64_t)N * K, 0, 0};
            g.out32 = o; g.ld_out = N;
            if (mode) { g.out_planes = pl; g.ld_pl = N; g.pl_plane_stride = (int64_t)M * N; }
            for (int it = 0; it < 2; ++it) {
                if (RUN(&g)) { printf("err %s\n", yv_last_error()); return 1; }
                cudaDeviceSynchronize();
                long long h[16];
                cudaMemcpyFromSymbol(h, yv_dbg, sizeof(h));
                if (it == 1)
                    printf("M=%d N=%d K=%d p=%d planes=%d: setup %lld first_full %lld mma_issued %lld epi_start %lld epi_end %lld exit %lld cyc\n",
                           M, N, K, passes, mode, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0]);
                if (it == 1) printf("      last chunk: tmem_ld done %lld staged %lld rows done %lld\n", h[8] - h[0], h[9] - h[0], h[10] - h[0]);
            }
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            for (int it = 0; it < 20; ++it) RUN(&g);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("   %.1f us per launch (%s)\n", ms / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(a); cudaFree(b); cudaFree(o); cudaFree(pl);
    }
    return 0;
}
