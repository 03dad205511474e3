// Developer tool: clock64 timeline of CTA (0,0,0) of the fused attention forward kernel (build: tools/attn_timing.sh).
#include "../youtube-vln_b200/csrc/yv_core.cu"
#include "../youtube-vln_b200/csrc/yv_attn.cu"
int main() {
    const int pairs = 8, heads = 8, dh = 128, Tq = 288, Tk = 288, H = heads * dh;
    __nv_bfloat16 *q, *k, *o; float* lse;
    const long long nq = 2ll * pairs * Tq * 3 * H, nk = 2ll * pairs * Tk * 3 * H;
    cudaMalloc(&q, nq * 2); cudaMalloc(&k, nk * 2); cudaMalloc(&o, 2ll * pairs * Tq * H * 2); cudaMalloc(&lse, 4ll * pairs * heads * Tq);
    cudaMemset(q, 0, nq * 2); cudaMemset(k, 0, nk * 2);
    for (int drop = 0; drop < 1; ++drop) {
        YvAttnFwd a; memset(&a, 0, sizeof(a));
        a.pairs = pairs; a.heads = heads; a.dh = dh; a.passes = 3;
        a.q = {q, 3 * H, (int64_t)pairs * Tq * 3 * H, (int64_t)Tq * 3 * H, Tq, 0};
        a.k = {k + H, 3 * H, (int64_t)pairs * Tk * 3 * H, (int64_t)Tk * 3 * H, Tk, 0};
        a.v = {k + 2 * H, 3 * H, (int64_t)pairs * Tk * 3 * H, (int64_t)Tk * 3 * H, Tk, 0};
        a.scale = 0.0883883f; a.out_planes = o; a.ld_out = H; a.out_plane_stride = (int64_t)pairs * Tq * H; a.lse = lse;
        for (int it = 0; it < 2; ++it) {
            if (yv_attn_fwd(&a, 0)) { printf("err %s\n", yv_last_error()); return 1; }
            cudaDeviceSynchronize();
        }
        long long h[128];
        cudaMemcpyFromSymbol(h, yv_adbg, sizeof(h));
        printf("fwd 288x288 dh128: Q,K landed %lld  epilogue start %lld end %lld\n", h[1] - h[0], h[2] - h[0], h[3] - h[0]);
        for (int j = 0; j < 5; ++j) {
            long long* c = h + 8 + 8 * j;
            printf("  chunk %d: S issue start %lld issued %lld | P waited %lld PV retired %lld | sm S ready %lld max xchg %lld exp done %lld P stored %lld\n",
                   j, c[0] - h[0], c[1] - h[0], c[2] - h[0], c[3] - h[0], c[4] - h[0], c[5] - h[0], c[6] - h[0], c[7] - h[0]);
            long long* d = h + 64 + 8 * j;
            printf("           sm: tmem_ld done %lld local max %lld | bar6 waited %lld tile stored %lld\n", d[0] - h[0], d[1] - h[0], d[2] - h[0], d[3] - h[0]);
        }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) yv_attn_fwd(&a, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("   %.1f us per launch (%s)\n", ms / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    {   // backward timeline
        __nv_bfloat16 *dO, *dq, *dkv; float* ws; unsigned* tk;
        cudaMalloc(&dO, 2ll * pairs * Tq * H * 2); cudaMalloc(&dq, 2ll * pairs * Tq * H * 2); cudaMalloc(&dkv, 2ll * pairs * Tk * 2 * H * 2);
        cudaMemset(dO, 0, 2ll * pairs * Tq * H * 2);
        size_t wsb = yv_attn_bwd_workspace_bytes(pairs, heads, dh, Tq, Tk);
        cudaMalloc(&ws, wsb); cudaMalloc(&tk, 4 * pairs * heads); cudaMemset(tk, 0, 4 * pairs * heads);
        cudaMemset(lse, 0, 4ll * pairs * heads * Tq);
        YvAttnBwd b; memset(&b, 0, sizeof(b));
        b.pairs = pairs; b.heads = heads; b.dh = dh; b.passes = 3;
        b.q = {q, 3 * H, (int64_t)pairs * Tq * 3 * H, (int64_t)Tq * 3 * H, Tq, 0};
        b.k = {k + H, 3 * H, (int64_t)pairs * Tk * 3 * H, (int64_t)Tk * 3 * H, Tk, 0};
        b.v = {k + 2 * H, 3 * H, (int64_t)pairs * Tk * 3 * H, (int64_t)Tk * 3 * H, Tk, 0};
        b.dout = {dO, H, (int64_t)pairs * Tq * H, (int64_t)Tq * H, Tq, 0};
        b.out = {o, H, (int64_t)pairs * Tq * H, (int64_t)Tq * H, Tq, 0};
        b.dq = {dq, H, (int64_t)pairs * Tq * H, (int64_t)Tq * H, Tq, 0};
        b.dk = {dkv, 2 * H, (int64_t)pairs * Tk * 2 * H, (int64_t)Tk * 2 * H, Tk, 0};
        b.dv = {dkv + H, 2 * H, (int64_t)pairs * Tk * 2 * H, (int64_t)Tk * 2 * H, Tk, 0};
        b.scale = 0.0883883f; b.lse = lse; b.workspace = ws; b.workspace_bytes = wsb; b.tickets = tk;
        for (int it = 0; it < 2; ++it) {
            if (yv_attn_bwd(&b, 0)) { printf("err %s\n", yv_last_error()); return 1; }
            cudaDeviceSynchronize();
        }
        long long h[128];
        cudaMemcpyFromSymbol(h, yv_adbg, sizeof(h));
        printf("bwd 288x288 dh128: loop end %lld  dQ staged %lld  ticket %lld  end %lld\n", h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0]);
        for (int j = 0; j < 5; ++j) {
            long long* c = h + 8 + 16 * j;
            printf("  chunk %d: A: KV landed %lld S issued %lld T waited %lld dV issued %lld next load %lld | B: dPd issued %lld dK,dQ issued %lld\n",
                   j, c[0] - h[0], c[1] - h[0], c[2] - h[0], c[3] - h[0], c[4] - h[0], c[5] - h[0], c[6] - h[0]);
            printf("           sm: S waited %lld math done %lld tiles stored %lld dV waited %lld dV drained %lld dKQ waited %lld dK drained %lld\n",
                   c[8] - h[0], c[9] - h[0], c[10] - h[0], c[11] - h[0], c[12] - h[0], c[13] - h[0], c[14] - h[0]);
        }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int it = 0; it < 20; ++it) yv_attn_bwd(&b, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("   %.1f us per launch (%s)\n", ms / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
