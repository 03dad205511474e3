// Developer tool: per-phase clock64 timeline of one yv_gemm CTA.  nvcc -DYV_GEMM_TIMING ... ; run on the GPU box.
#include "../youtube-vln_b200/csrc/yv_core.cu"
#include "../youtube-vln_b200/csrc/yv_gemm.cu"
#include <vector>
int main() {
    int Ms[] = {128, 128, 2304}, Ns[] = {128, 128, 1024}, Ks[] = {64, 4096, 1024};
    for (int t = 0; t < 3; ++t) {
        int M = Ms[t], N = Ns[t], K = Ks[t];
        __nv_bfloat16 *a, *b; float* o;
        cudaMalloc(&a, 2ll * M * K * 2); cudaMalloc(&b, 2ll * N * K * 2); cudaMalloc(&o, 4ll * M * N);
        cudaMemset(a, 0, 2ll * M * K * 2); cudaMemset(b, 0, 2ll * N * K * 2);
        YvGemm g; memset(&g, 0, sizeof(g));
        g.M = M; g.N = N; g.K = K; g.passes = 3; g.alpha = 1.f;
        g.a = {a, K, M, K, 1, 0, 1, 0, (int64_t)M * K, 0, 0};
        g.b = {b, K, N, K, 1, 0, 1, 0, (int64_t)N * K, 0, 0};
        g.out32 = o; g.ld_out = N;
        for (int it = 0; it < 3; ++it) {
            if (yv_gemm(&g, 0)) { printf("err %s\n", yv_last_error()); return 1; }
            cudaDeviceSynchronize();
            long long h[8];
            cudaMemcpyFromSymbol(h, yv_dbg, sizeof(h));
            printf("M=%d N=%d K=%d it%d: setup %lld  first_full %lld  mma_done %lld  epi_start %lld  epi_end %lld  exit %lld cycles\n",
                   M, N, K, it, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0]);
        }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) yv_gemm(&g, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("   %.1f us per launch (%s)\n", ms / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
