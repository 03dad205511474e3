// Row-wise / elementwise kernels of the ViLBERT hot path (HBM/L2-bound work; no tensor cores).
// One warp per row for LayerNorm / softmax (rows are 64..1152 floats), coalesced lane-strided columns.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/yvb200.h"
#include "yv_common.cuh"

void yv_count_launch();

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 hi/lo planes
// ------------------------------------------------------------------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                    long long ld_dst, long long plane_stride, long long rows, long long cols) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const bool vec = ((cols & 3) == 0) && ((ld_src & 3) == 0) && ((ld_dst & 3) == 0) && ((plane_stride & 3) == 0) &&
                     ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dst) & 7) == 0);
    if (vec) {
        const long long c4 = cols >> 2, total = rows * c4;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
             i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / c4, c = (i - r * c4) << 2;
            const float4 v = *reinterpret_cast<const float4*>(src + r * ld_src + c);
            __align__(8) __nv_bfloat16 h[4], l[4];
            yv_split(v.x, h[0], l[0]); yv_split(v.y, h[1], l[1]); yv_split(v.z, h[2], l[2]); yv_split(v.w, h[3], l[3]);
            *reinterpret_cast<uint2*>(dst + r * ld_dst + c) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(dst + plane_stride + r * ld_dst + c) = *reinterpret_cast<uint2*>(l);
        }
        return;
    }
    const long long total = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        __nv_bfloat16 h, l;
        yv_split(src[r * ld_src + c], h, l);
        dst[r * ld_dst + c] = h;
        dst[plane_stride + r * ld_dst + c] = l;
    }
}

constexpr int SPLIT_BLK = 2048;  // elements per block in the multi-tensor kernel
__global__ void split_multi_kernel(const YvSplitSeg* __restrict__ segs, int nseg, __nv_bfloat16* __restrict__ planes,
                                   long long plane_stride) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long blk = blockIdx.x;
    int lo = 0, hi = nseg - 1;                       // last segment with first_blk <= blk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].first_blk <= blk) lo = mid; else hi = mid - 1;
    }
    const YvSplitSeg s = segs[lo];
    const long long base = (blk - s.first_blk) * SPLIT_BLK;
    const float* src = s.src + base;
    __nv_bfloat16* dh = planes + s.dst_off + base;
    __nv_bfloat16* dl = dh + plane_stride;
    const long long n = min((long long)SPLIT_BLK, s.numel - base);
    const bool vec = ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dh) & 7) == 0) && ((plane_stride & 3) == 0);
    if (vec && n == SPLIT_BLK) {
        for (int i = threadIdx.x; i < SPLIT_BLK / 4; i += blockDim.x) {
            const float4 v = reinterpret_cast<const float4*>(src)[i];
            __align__(8) __nv_bfloat16 h[4], l[4];
            yv_split(v.x, h[0], l[0]); yv_split(v.y, h[1], l[1]); yv_split(v.z, h[2], l[2]); yv_split(v.w, h[3], l[3]);
            reinterpret_cast<uint2*>(dh)[i] = *reinterpret_cast<uint2*>(h);
            reinterpret_cast<uint2*>(dl)[i] = *reinterpret_cast<uint2*>(l);
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            __nv_bfloat16 h, l;
            yv_split(src[i], h, l);
            dh[i] = h;
            dl[i] = l;
        }
    }
}

// Fused multi-tensor AdamW (vilbert/optimization.py:141-187): one launch updates p / exp_avg / exp_avg_sq of every
// parameter and, for GEMM weights, re-splits the new value into the bf16 hi/lo planes the next forward reads.
// hyper = {lr, step_size (= lr * sqrt(1 - b2^t) / (1 - b1^t) or lr), beta1, beta2, eps, 1-beta1, 1-beta2}
constexpr int ADAM_BLK = 2048;
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const YvAdamSeg* __restrict__ segs, int nseg, const float* __restrict__ hyper) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long blk = blockIdx.x;
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].first_blk <= blk) lo = mid; else hi = mid - 1;
    }
    const YvAdamSeg s = segs[lo];
    const float lr = hyper[0], step_size = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4];
    const float omb1 = hyper[5], omb2 = hyper[6];   // 1-beta rounded from double on the host, as the reference does
    const float decay = lr * s.weight_decay;
    const long long base = (blk - s.first_blk) * ADAM_BLK;
    const long long n = min((long long)ADAM_BLK, s.numel - base);
    __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(s.plane_hi);
    __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(s.plane_lo);
    auto update = [&](float g, float& m, float& v, float& w) {
        m = m * b1 + omb1 * g;
        v = v * b2 + omb2 * (g * g);
        w = w - step_size * (m / (sqrtf(v) + eps));
        if (s.weight_decay > 0.f) w = w - decay * w;
    };
    const bool vec = (((uintptr_t)(s.p + base) | (uintptr_t)(s.g + base) | (uintptr_t)(s.m + base) | (uintptr_t)(s.v + base)) & 15) == 0 &&
                     (ph == nullptr || ((((uintptr_t)(ph + base) | (uintptr_t)(pl + base)) & 7) == 0));
    const long long n4 = vec ? (n >> 2) : 0;
    for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
        const long long j = base + 4 * i;
        const float4 g4 = *reinterpret_cast<const float4*>(s.g + j);
        float4 m4 = *reinterpret_cast<float4*>(s.m + j);
        float4 v4 = *reinterpret_cast<float4*>(s.v + j);
        float4 w4 = *reinterpret_cast<float4*>(s.p + j);
        update(g4.x, m4.x, v4.x, w4.x);
        update(g4.y, m4.y, v4.y, w4.y);
        update(g4.z, m4.z, v4.z, w4.z);
        update(g4.w, m4.w, v4.w, w4.w);
        *reinterpret_cast<float4*>(s.m + j) = m4;
        *reinterpret_cast<float4*>(s.v + j) = v4;
        *reinterpret_cast<float4*>(s.p + j) = w4;
        if (ph) {
            __align__(8) __nv_bfloat16 h[4], l[4];
            yv_split(w4.x, h[0], l[0]); yv_split(w4.y, h[1], l[1]); yv_split(w4.z, h[2], l[2]); yv_split(w4.w, h[3], l[3]);
            *reinterpret_cast<uint2*>(ph + j) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(pl + j) = *reinterpret_cast<uint2*>(l);
        }
    }
    for (long long i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
        const long long j = base + i;
        float m = s.m[j], v = s.v[j], w = s.p[j];
        update(s.g[j], m, v, w);
        s.m[j] = m;
        s.v[j] = v;
        s.p[j] = w;
        if (ph) {
            __nv_bfloat16 h, l;
            yv_split(w, h, l);
            ph[j] = h;
            pl[j] = l;
        }
    }
}

__global__ void rng_advance_kernel(unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait(); rng[1] += 1ULL; }

// ------------------------------------------------------------------------------------------------
// LayerNorm
// ------------------------------------------------------------------------------------------------
constexpr int LNF_ROWS = 4;
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ y32, __nv_bfloat16* __restrict__ yp, long long plane_stride,
                     float* __restrict__ stats, long long M, int C, float drop_p, unsigned drop_site,
                     const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    // thread t owns columns [4t, 4t+4); LNF_ROWS rows per block; two block reductions (mean, then centred variance)
    __shared__ float red[8][LNF_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int c0 = tid * 4;
    const bool active = c0 < C;
    const long long row0 = blockIdx.x * (long long)LNF_ROWS;
    float v[LNF_ROWS][4], part[LNF_ROWS];
#pragma unroll
    for (int r = 0; r < LNF_ROWS; ++r) {
        const long long row = row0 + r;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && row < M) t = *reinterpret_cast<const float4*>(x + row * C + c0);
        v[r][0] = t.x; v[r][1] = t.y; v[r][2] = t.z; v[r][3] = t.w;
        part[r] = (t.x + t.y) + (t.z + t.w);
    }
    float mean[LNF_ROWS], rstd[LNF_ROWS];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int r = 0; r < LNF_ROWS; ++r) part[r] = yv_warp_sum(part[r]);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < LNF_ROWS; ++r) red[warp][r] = part[r];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LNF_ROWS; ++r) {
            float t = 0.f;
            for (int w = 0; w < nwarps; ++w) t += red[w][r];
            if (pass == 0) {
                mean[r] = t / C;
                float q = 0.f;
                if (active) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float d = v[r][j] - mean[r];
                        q += d * d;
                    }
                }
                part[r] = q;
            } else {
                rstd[r] = 1.f / sqrtf(t / C + eps);
            }
        }
        __syncthreads();
    }
    if (!active) return;
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + c0);
    const float4 b4 = *reinterpret_cast<const float4*>(beta + c0);
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
#pragma unroll
    for (int r = 0; r < LNF_ROWS; ++r) {
        const long long row = row0 + r;
        if (row >= M) break;
        if (stats && tid == 0) {
            stats[2 * row] = mean[r];
            stats[2 * row + 1] = rstd[r];
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j] = gv[j] * ((v[r][j] - mean[r]) * rstd[r]) + bv[j];
            if (drop.thresh) o[j] *= yv_drop_mul(drop, (uint32_t)(row * C + c0 + j));
        }
        if (y32) *reinterpret_cast<float4*>(y32 + row * C + c0) = make_float4(o[0], o[1], o[2], o[3]);
        if (yp) {
            __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) yv_split(o[j], h4[j], l4[j]);
            *reinterpret_cast<uint2*>(yp + row * C + c0) = *reinterpret_cast<uint2*>(h4);
            *reinterpret_cast<uint2*>(yp + plane_stride + row * C + c0) = *reinterpret_cast<uint2*>(l4);
        }
    }
}

// Thread t owns columns [4t, 4t+4); a block walks the rows four at a time (grid-stride).  Row statistics need a
// block reduction (shuffle + one smem exchange per 4 rows); dgamma / dbeta / dbias live in 12 registers per thread
// and leave through one vector reduction per thread at the end.
constexpr int LNB_ROWS = 4;
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ stats, float post_p, unsigned post_site, const float* __restrict__ dx_add,
                     float* __restrict__ dx32, __nv_bfloat16* __restrict__ dxp, long long plane_stride, float pre_p,
                     unsigned pre_site, const unsigned long long* rng, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias, long long M, int C) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float red[8][2 * LNB_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int c0 = tid * 4;
    const bool active = c0 < C;
    const YvDrop dpost = yv_drop_make(rng, post_site, post_p);
    const YvDrop dpre = yv_drop_make(rng, pre_site, pre_p);
    float4 gm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) gm = *reinterpret_cast<const float4*>(gamma + c0);
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, abias[4] = {0.f, 0.f, 0.f, 0.f};
    const float invC = 1.f / C;
    for (long long row0 = blockIdx.x * (long long)LNB_ROWS; row0 < M; row0 += (long long)gridDim.x * LNB_ROWS) {
        float g[LNB_ROWS][4], xh[LNB_ROWS][4], part[2 * LNB_ROWS], rs[LNB_ROWS];
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            const long long row = row0 + r;
            part[2 * r] = part[2 * r + 1] = 0.f;
            rs[r] = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) g[r][j] = xh[r][j] = 0.f;
            if (active && row < M) {
                const float mean = stats[2 * row];
                rs[r] = stats[2 * row + 1];
                const float4 d4 = *reinterpret_cast<const float4*>(dy + row * C + c0);
                const float4 x4 = *reinterpret_cast<const float4*>(x + row * C + c0);
                float d[4] = {d4.x, d4.y, d4.z, d4.w};
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                const float gv[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (dpost.thresh) d[j] *= yv_drop_mul(dpost, (uint32_t)(row * C + c0 + j));
                    xh[r][j] = (xv[j] - mean) * rs[r];
                    ag[j] += d[j] * xh[r][j];
                    ab[j] += d[j];
                    g[r][j] = d[j] * gv[j];
                    part[2 * r] += g[r][j];
                    part[2 * r + 1] += g[r][j] * xh[r][j];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 2 * LNB_ROWS; ++i) part[i] = yv_warp_sum(part[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 2 * LNB_ROWS; ++i) red[warp][i] = part[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2 * LNB_ROWS; ++i) {
            float t = 0.f;
            for (int w = 0; w < nwarps; ++w) t += red[w][i];
            part[i] = t * invC;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LNB_ROWS; ++r) {
            const long long row = row0 + r;
            if (!(active && row < M)) continue;
            float o[4], om[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = rs[r] * (g[r][j] - part[2 * r] - xh[r][j] * part[2 * r + 1]);
            if (dx_add) {
                const float4 a4 = *reinterpret_cast<const float4*>(dx_add + row * C + c0);
                o[0] += a4.x; o[1] += a4.y; o[2] += a4.z; o[3] += a4.w;
            }
            if (dx32) *reinterpret_cast<float4*>(dx32 + row * C + c0) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                om[j] = o[j];
                if (dpre.thresh) om[j] *= yv_drop_mul(dpre, (uint32_t)(row * C + c0 + j));
                abias[j] += om[j];
            }
            if (dxp) {
                __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) yv_split(om[j], h4[j], l4[j]);
                *reinterpret_cast<uint2*>(dxp + row * C + c0) = *reinterpret_cast<uint2*>(h4);
                *reinterpret_cast<uint2*>(dxp + plane_stride + row * C + c0) = *reinterpret_cast<uint2*>(l4);
            }
        }
    }
    if (!active) return;
    if (dgamma) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dgamma + c0), "f"(ag[0]), "f"(ag[1]), "f"(ag[2]), "f"(ag[3]) : "memory");
    if (dbeta) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbeta + c0), "f"(ab[0]), "f"(ab[1]), "f"(ab[2]), "f"(ab[3]) : "memory");
    if (dbias) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dbias + c0), "f"(abias[0]), "f"(abias[1]), "f"(abias[2]), "f"(abias[3]) : "memory");
}

// ------------------------------------------------------------------------------------------------
// LayerNorm, warp-per-row variants (default).  A row (C <= 1024) lives in the registers of one warp: lane l owns the
// float4 column groups l, l + 32, ... so every access is a 512-byte coalesced warp transaction and the row statistics
// need warp shuffles only -- no shared memory, no block barrier on the dependency chain of the step.
// ------------------------------------------------------------------------------------------------
constexpr int LNW_WARPS = 8;      // rows in flight per block
constexpr int LNW_MAXV = 8;       // float4 groups per lane: C <= 32 * 4 * 8 = 1024

template <int NV>
__global__ void __launch_bounds__(LNW_WARPS * 32)
layernorm_fwd_warp_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float eps, float* __restrict__ y32, __nv_bfloat16* __restrict__ yp, long long plane_stride,
                          float* __restrict__ stats, long long M, int C, float drop_p, unsigned drop_site,
                          const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = blockIdx.x * (long long)LNW_WARPS + warp;
    if (row >= M) return;
    const float* xr = x + row * C;
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        v[j] = c < C ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = yv_warp_sum(sum) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        if (c < C) {
            const float a = v[j].x - mean, b = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
            q += (a * a + b * b) + (d * d + e * e);
        }
    }
    const float rstd = 1.f / sqrtf(yv_warp_sum(q) / C + eps);
    if (stats && lane == 0) {
        stats[2 * row] = mean;
        stats[2 * row + 1] = rstd;
    }
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        if (c >= C) continue;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
        float o[4] = {g4.x * ((v[j].x - mean) * rstd) + b4.x, g4.y * ((v[j].y - mean) * rstd) + b4.y,
                      g4.z * ((v[j].z - mean) * rstd) + b4.z, g4.w * ((v[j].w - mean) * rstd) + b4.w};
        if (drop.thresh) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] *= yv_drop_mul(drop, (uint32_t)(row * C + c + e));
        }
        if (y32) *reinterpret_cast<float4*>(y32 + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (yp) {
            __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) yv_split(o[e], h4[e], l4[e]);
            *reinterpret_cast<uint2*>(yp + row * C + c) = *reinterpret_cast<uint2*>(h4);
            *reinterpret_cast<uint2*>(yp + plane_stride + row * C + c) = *reinterpret_cast<uint2*>(l4);
        }
    }
}

// Backward: each warp walks rows (grid-stride over warps); the column sums dgamma / dbeta / dbias stay in registers per
// lane, are combined across the block's warps through shared memory and leave as one vector reduction per block and
// column group.
template <int NV, bool COLS>
__global__ void __launch_bounds__(LNW_WARPS * 32)
layernorm_bwd_warp_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                          const float* __restrict__ stats, float post_p, unsigned post_site,
                          const float* __restrict__ dx_add, float* __restrict__ dx32, __nv_bfloat16* __restrict__ dxp,
                          long long plane_stride, float pre_p, unsigned pre_site, const unsigned long long* rng,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, long long M,
                          int C) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float4 red[LNW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const YvDrop dpost = yv_drop_make(rng, post_site, post_p);
    const YvDrop dpre = yv_drop_make(rng, pre_site, pre_p);
    float4 gm[NV], ag[NV], ab[NV], abias[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 4;
        gm[j] = c < C ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        ag[j] = ab[j] = abias[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float invC = 1.f / C;
    const long long nwarps = (long long)gridDim.x * LNW_WARPS;
    for (long long row = blockIdx.x * (long long)LNW_WARPS + warp; row < M; row += nwarps) {
        const float mean = stats[2 * row], rs = stats[2 * row + 1];
        float4 g[NV], xh[NV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            g[j] = xh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) {
                float4 d = *reinterpret_cast<const float4*>(dy + row * C + c);
                const float4 xv = *reinterpret_cast<const float4*>(x + row * C + c);
                if (dpost.thresh) {
                    const uint32_t i0 = (uint32_t)(row * C + c);
                    d.x *= yv_drop_mul(dpost, i0); d.y *= yv_drop_mul(dpost, i0 + 1);
                    d.z *= yv_drop_mul(dpost, i0 + 2); d.w *= yv_drop_mul(dpost, i0 + 3);
                }
                xh[j] = make_float4((xv.x - mean) * rs, (xv.y - mean) * rs, (xv.z - mean) * rs, (xv.w - mean) * rs);
                if (COLS) {
                    ag[j].x += d.x * xh[j].x; ag[j].y += d.y * xh[j].y; ag[j].z += d.z * xh[j].z; ag[j].w += d.w * xh[j].w;
                    ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
                }
                g[j] = make_float4(d.x * gm[j].x, d.y * gm[j].y, d.z * gm[j].z, d.w * gm[j].w);
                s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
                s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
            }
        }
        s1 = yv_warp_sum(s1) * invC;
        s2 = yv_warp_sum(s2) * invC;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 4;
            if (c >= C) continue;
            float o[4] = {rs * (g[j].x - s1 - xh[j].x * s2), rs * (g[j].y - s1 - xh[j].y * s2),
                          rs * (g[j].z - s1 - xh[j].z * s2), rs * (g[j].w - s1 - xh[j].w * s2)};
            if (dx_add) {
                const float4 a4 = *reinterpret_cast<const float4*>(dx_add + row * C + c);
                o[0] += a4.x; o[1] += a4.y; o[2] += a4.z; o[3] += a4.w;
            }
            if (dx32) *reinterpret_cast<float4*>(dx32 + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
            if (dpre.thresh) {
                const uint32_t i0 = (uint32_t)(row * C + c);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] *= yv_drop_mul(dpre, i0 + e);
            }
            if (COLS) { abias[j].x += o[0]; abias[j].y += o[1]; abias[j].z += o[2]; abias[j].w += o[3]; }
            if (dxp) {
                __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) yv_split(o[e], h4[e], l4[e]);
                *reinterpret_cast<uint2*>(dxp + row * C + c) = *reinterpret_cast<uint2*>(h4);
                *reinterpret_cast<uint2*>(dxp + plane_stride + row * C + c) = *reinterpret_cast<uint2*>(l4);
            }
        }
    }
    if (!COLS) return;
    // block-level column sums: one array and one column group at a time through 4 KB of shared memory
#pragma unroll
    for (int which = 0; which < 3; ++which) {
        float* out = which == 0 ? dgamma : (which == 1 ? dbeta : dbias);
        if (out == nullptr) continue;                        // block-uniform
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (j * 128 >= C) break;                         // block-uniform
            red[warp][lane] = which == 0 ? ag[j] : (which == 1 ? ab[j] : abias[j]);
            __syncthreads();
            if (warp == 0) {
                float4 t = red[0][lane];
#pragma unroll
                for (int w = 1; w < LNW_WARPS; ++w) {
                    const float4 u = red[w][lane];
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                const int c = (j * 32 + lane) * 4;
                if (c < C)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c), "f"(t.x), "f"(t.y),
                                 "f"(t.z), "f"(t.w)
                                 : "memory");
            }
            __syncthreads();
        }
    }
}

// dgamma[c] += sum_m dy[m, c] * xhat[m, c],  dbeta[c] += sum_m dy[m, c]  -- the column reductions of the LayerNorm
// backward as a kernel of their own: nothing on the dependency chain of the backward pass needs them, so the step issues
// them on a trailing stream next to the weight gradients while layernorm_bwd_warp_kernel<NV, false> (dx only, a third of
// the registers, four blocks per SM) stays on the chain.  blockIdx.x = 128-column group, blockIdx.y = row split.
__global__ void __launch_bounds__(LNW_WARPS * 32)
layernorm_bwd_cols_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C, int rows_per_split) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float4 red[2][LNW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4;
    const long long r0 = (long long)blockIdx.y * rows_per_split;
    const long long r1 = min(M, r0 + rows_per_split);
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
    if (c < C) {
        for (long long row = r0 + warp; row < r1; row += LNW_WARPS) {
            const float mean = __ldg(stats + 2 * row), rs = __ldg(stats + 2 * row + 1);
            const float4 d = *reinterpret_cast<const float4*>(dy + row * C + c);
            const float4 xv = *reinterpret_cast<const float4*>(x + row * C + c);
            ag.x += d.x * ((xv.x - mean) * rs); ag.y += d.y * ((xv.y - mean) * rs);
            ag.z += d.z * ((xv.z - mean) * rs); ag.w += d.w * ((xv.w - mean) * rs);
            ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
        }
    }
    red[0][warp][lane] = ag;
    red[1][warp][lane] = ab;
    __syncthreads();
    if (warp < 2 && c < C) {
        float4 t = red[warp][0][lane];
#pragma unroll
        for (int w = 1; w < LNW_WARPS; ++w) {
            const float4 u = red[warp][w][lane];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        float* out = warp == 0 ? dgamma : dbeta;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// attention softmax
// ------------------------------------------------------------------------------------------------
// one warp per row, the row lives in registers (cols <= 32 * MAXPL): one read and one write of the scores
template <int MAXPL>
__global__ void __launch_bounds__(THREADS)
softmax_fwd_kernel(float* __restrict__ s, long long ld_s, const float* __restrict__ mask, long long rows, int cols,
                   long long rows_per_pair, float scale, __nv_bfloat16* __restrict__ pp, long long ld_p,
                   long long plane_stride, float drop_p, unsigned drop_site, const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * (long long)WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* sr = s + row * ld_s;
    const float* mr = mask ? mask + (row / rows_per_pair) * cols : nullptr;
    float v[MAXPL];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
        const int c = lane + 32 * i;
        v[i] = -INFINITY;
        if (c < cols) v[i] = sr[c] * scale + (mr ? mr[c] : 0.f);
        mx = fmaxf(mx, v[i]);
    }
    mx = yv_warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
        v[i] = (lane + 32 * i < cols) ? expf(v[i] - mx) : 0.f;
        sum += v[i];
    }
    sum = yv_warp_sum(sum);
    const float inv = 1.f / sum;
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
        const int c = lane + 32 * i;
        if (c >= cols) break;
        const float pr = v[i] * inv;
        sr[c] = pr;
        float pd = pr;
        if (drop.thresh) pd *= yv_drop_mul(drop, (uint32_t)(row * cols + c));
        __nv_bfloat16 h, l;
        yv_split(pd, h, l);
        pp[row * ld_p + c] = h;
        pp[plane_stride + row * ld_p + c] = l;
    }
}

template <int MAXPL>
__global__ void __launch_bounds__(THREADS)
softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dpd, long long ld_s, long long rows, int cols,
                   float scale, __nv_bfloat16* __restrict__ dsp, long long ld_p, long long plane_stride, float drop_p,
                   unsigned drop_site, const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * (long long)WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* pr = p + row * ld_s;
    const float* dr = dpd + row * ld_s;
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
    float pv[MAXPL], dv[MAXPL];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
        const int c = lane + 32 * i;
        pv[i] = dv[i] = 0.f;
        if (c < cols) {
            pv[i] = pr[c];
            dv[i] = dr[c];
            if (drop.thresh) dv[i] *= yv_drop_mul(drop, (uint32_t)(row * cols + c));
            dot += dv[i] * pv[i];
        }
    }
    dot = yv_warp_sum(dot);
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
        const int c = lane + 32 * i;
        if (c >= cols) break;
        const float ds = scale * pv[i] * (dv[i] - dot);
        __nv_bfloat16 h, l;
        yv_split(ds, h, l);
        dsp[row * ld_p + c] = h;
        dsp[plane_stride + row * ld_p + c] = l;
    }
}

// Vectorised variants (cols, leading dimensions and plane stride multiples of 4, 16-byte aligned rows): lane l owns the
// float4 column groups l, l + 32, ... -- 512-byte warp loads, 8-byte plane stores, a quarter of the memory
// instructions of the scalar kernels above.  Same dropout keys (row * cols + column) as the scalar kernels.
template <int NV>
__global__ void __launch_bounds__(THREADS)
softmax_fwd_vec_kernel(float* __restrict__ s, long long ld_s, const float* __restrict__ mask, long long rows, int cols,
                       long long rows_per_pair, float scale, __nv_bfloat16* __restrict__ pp, long long ld_p,
                       long long plane_stride, float drop_p, unsigned drop_site, const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * (long long)WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* sr = s + row * ld_s;
    const float* mr = mask ? mask + (row / rows_per_pair) * cols : nullptr;
    float4 v[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (c < cols) {
            const float4 x = *reinterpret_cast<const float4*>(sr + c);
            const float4 m = mr ? __ldg(reinterpret_cast<const float4*>(mr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[i] = make_float4(x.x * scale + m.x, x.y * scale + m.y, x.z * scale + m.z, x.w * scale + m.w);
        }
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
    mx = yv_warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if ((lane + 32 * i) * 4 < cols)
            v[i] = make_float4(expf(v[i].x - mx), expf(v[i].y - mx), expf(v[i].z - mx), expf(v[i].w - mx));
        else
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    sum = yv_warp_sum(sum);
    const float inv = 1.f / sum;
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c >= cols) break;
        float pr[4] = {v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv};
        *reinterpret_cast<float4*>(sr + c) = make_float4(pr[0], pr[1], pr[2], pr[3]);
        __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float pd = pr[e];
            if (drop.thresh) pd *= yv_drop_mul(drop, (uint32_t)(row * cols + c + e));
            yv_split(pd, h4[e], l4[e]);
        }
        *reinterpret_cast<uint2*>(pp + row * ld_p + c) = *reinterpret_cast<uint2*>(h4);
        *reinterpret_cast<uint2*>(pp + plane_stride + row * ld_p + c) = *reinterpret_cast<uint2*>(l4);
    }
}

template <int NV>
__global__ void __launch_bounds__(THREADS)
softmax_bwd_vec_kernel(const float* __restrict__ p, const float* __restrict__ dpd, long long ld_s, long long rows, int cols,
                       float scale, __nv_bfloat16* __restrict__ dsp, long long ld_p, long long plane_stride, float drop_p,
                       unsigned drop_site, const unsigned long long* rng) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * (long long)WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* pr = p + row * ld_s;
    const float* dr = dpd + row * ld_s;
    const YvDrop drop = yv_drop_make(rng, drop_site, drop_p);
    float4 pv[NV], dv[NV];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        pv[i] = dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < cols) {
            pv[i] = *reinterpret_cast<const float4*>(pr + c);
            dv[i] = *reinterpret_cast<const float4*>(dr + c);
            if (drop.thresh) {
                const uint32_t i0 = (uint32_t)(row * cols + c);
                dv[i].x *= yv_drop_mul(drop, i0); dv[i].y *= yv_drop_mul(drop, i0 + 1);
                dv[i].z *= yv_drop_mul(drop, i0 + 2); dv[i].w *= yv_drop_mul(drop, i0 + 3);
            }
            dot += (dv[i].x * pv[i].x + dv[i].y * pv[i].y) + (dv[i].z * pv[i].z + dv[i].w * pv[i].w);
        }
    }
    dot = yv_warp_sum(dot);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (lane + 32 * i) * 4;
        if (c >= cols) break;
        const float ds[4] = {scale * pv[i].x * (dv[i].x - dot), scale * pv[i].y * (dv[i].y - dot),
                             scale * pv[i].z * (dv[i].z - dot), scale * pv[i].w * (dv[i].w - dot)};
        __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) yv_split(ds[e], h4[e], l4[e]);
        *reinterpret_cast<uint2*>(dsp + row * ld_p + c) = *reinterpret_cast<uint2*>(h4);
        *reinterpret_cast<uint2*>(dsp + plane_stride + row * ld_p + c) = *reinterpret_cast<uint2*>(l4);
    }
}

// ------------------------------------------------------------------------------------------------
// embeddings
// ------------------------------------------------------------------------------------------------
__global__ void embed_text_fwd_kernel(const long long* __restrict__ tok, const long long* __restrict__ seg,
                                      const float* __restrict__ word, const float* __restrict__ pos,
                                      const float* __restrict__ type, float* __restrict__ out, long long M, int T, int H) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long m = blockIdx.x;
    const float* w = word + tok[m] * (long long)H;
    const float* ps = pos + (m % T) * (long long)H;
    const float* ty = type + seg[m] * (long long)H;
    for (int h = threadIdx.x; h < H; h += blockDim.x) out[m * H + h] = w[h] + ps[h] + ty[h];
}

__global__ void embed_text_bwd_kernel(const long long* __restrict__ tok, const long long* __restrict__ seg,
                                      const float* __restrict__ dout, float* __restrict__ dword, float* __restrict__ dpos,
                                      float* __restrict__ dtype_, long long M, int T, int H, int padding_idx) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long m = blockIdx.x;
    const long long t = tok[m];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const float g = dout[m * H + h];
        if (dword && t != padding_idx) atomicAdd(dword + t * H + h, g);
        if (dpos) atomicAdd(dpos + (m % T) * (long long)H + h, g);
        if (dtype_) atomicAdd(dtype_ + seg[m] * (long long)H + h, g);
    }
}

__global__ void embed_loc_fwd_kernel(const float* __restrict__ loc, const float* __restrict__ w5, const float* __restrict__ b5,
                                     const float* __restrict__ w4, const float* __restrict__ b4, const float* __restrict__ w2,
                                     const float* __restrict__ b2, const float* __restrict__ seq, float* __restrict__ out,
                                     long long M, int H) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long m = blockIdx.x;
    __shared__ float l[12];
    if (threadIdx.x < 12) l[threadIdx.x] = loc[m * 12 + threadIdx.x];
    __syncthreads();
    const long long sidx = (long long)l[11];
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        // same association order as the reference: ((a + b) + c) + d with a,b,c = W.x + bias
        float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) a += w5[h * 5 + j] * l[j];
#pragma unroll
        for (int j = 0; j < 4; ++j) b += w4[h * 4 + j] * l[5 + j];
#pragma unroll
        for (int j = 0; j < 2; ++j) c += w2[h * 2 + j] * l[9 + j];
        out[m * H + h] = (((a + b5[h]) + (b + b4[h])) + (c + b2[h])) + seq[sidx * H + h];
    }
}

constexpr int LOC_ROWS = 32;
constexpr int LOC_SPLIT = 2;                   // CTAs per row block (each owns H / LOC_SPLIT columns): 144 CTAs at cfg2
// One CTA: 32 rows x a column range.  The frame-embedding gradient (dseq) is accumulated in registers for as long as
// consecutive rows belong to the same frame (rows arrive frame by frame: one reduction per run instead of one per row --
// the per-row reductions were most of the kernel's 67 us at the very end of the backward chain).
__global__ void __launch_bounds__(256)
embed_loc_bwd_kernel(const float* __restrict__ loc, const float* __restrict__ dout, float* __restrict__ dw5,
                     float* __restrict__ db5, float* __restrict__ dw4, float* __restrict__ db4,
                     float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ dseq, long long M,
                     int H) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float l[LOC_ROWS][12];
    const long long r0 = blockIdx.x * (long long)LOC_ROWS;
    const int nr = (int)min((long long)LOC_ROWS, M - r0);
    for (int i = threadIdx.x; i < nr * 12; i += blockDim.x) l[i / 12][i % 12] = loc[r0 * 12 + i];
    __syncthreads();
    const int hs = (H + LOC_SPLIT - 1) / LOC_SPLIT;
    const int h1 = min(H, (int)(blockIdx.y + 1) * hs);
    for (int h = blockIdx.y * hs + threadIdx.x; h < h1; h += blockDim.x) {
        float acc[11];
#pragma unroll
        for (int j = 0; j < 11; ++j) acc[j] = 0.f;
        float accb = 0.f, accs = 0.f;
        int cur = (int)l[0][11];
        for (int r = 0; r < nr; ++r) {
            const float g = dout[(r0 + r) * H + h];
            accb += g;
#pragma unroll
            for (int j = 0; j < 11; ++j) acc[j] += g * l[r][j];
            const int si = (int)l[r][11];
            if (si != cur) {                                 // (block-uniform)
                atomicAdd(dseq + (long long)cur * H + h, accs);
                accs = 0.f;
                cur = si;
            }
            accs += g;
        }
        atomicAdd(dseq + (long long)cur * H + h, accs);
#pragma unroll
        for (int j = 0; j < 5; ++j) atomicAdd(dw5 + h * 5 + j, acc[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dw4 + h * 4 + j, acc[5 + j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) atomicAdd(dw2 + h * 2 + j, acc[9 + j]);
        atomicAdd(db5 + h, accb);
        atomicAdd(db4 + h, accb);
        atomicAdd(db2 + h, accb);
    }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 64;
__global__ void colsum_kernel(const float* __restrict__ x, long long ld, long long rows, int cols, float* __restrict__ out) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long long r0 = blockIdx.y * (long long)CS_ROWS;
    const long long r1 = min(rows, r0 + CS_ROWS);
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += x[r * ld + c];
    atomicAdd(out + c, s);
}

// planes variant: x = hi + lo.  Thread owns 8 columns (one 16-byte load per plane per row), a block covers
// 1024 columns x CSP_ROWS rows; partial sums leave through vector reductions.
constexpr int CSP_ROWS = 16;
__global__ void __launch_bounds__(128)
colsum_planes_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long plane_stride, long long rows, int cols,
                     float* __restrict__ out) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const int c0 = (blockIdx.x * 128 + threadIdx.x) * 8;
    if (c0 >= cols) return;
    const long long r0 = blockIdx.y * (long long)CSP_ROWS;
    const long long r1 = min(rows, r0 + CSP_ROWS);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c0 + 8 <= cols && (ld & 7) == 0 && (plane_stride & 7) == 0) {
#pragma unroll 4
        for (long long r = r0; r < r1; ++r) {
            const uint4 h = *reinterpret_cast<const uint4*>(x + r * ld + c0);
            const uint4 l = *reinterpret_cast<const uint4*>(x + plane_stride + r * ld + c0);
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
            const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 a = __bfloat1622float2(hp[j]), b = __bfloat1622float2(lp[j]);
                acc[2 * j] += a.x + b.x;
                acc[2 * j + 1] += a.y + b.y;
            }
        }
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c0), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + c0 + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
    } else {
        for (int j = 0; j < 8 && c0 + j < cols; ++j) {
            float t = 0.f;
            for (long long r = r0; r < r1; ++r)
                t += __bfloat162float(x[r * ld + c0 + j]) + __bfloat162float(x[plane_stride + r * ld + c0 + j]);
            atomicAdd(out + c0 + j, t);
        }
    }
}

// dpre = dy * act'(aux) -> planes   (GELU: aux = pre-activation; ReLU: aux = forward output); optional column sums
// of dpre (the bias gradient).  Thread owns 4 columns, a block covers 1024 columns x ABS_ROWS rows.
constexpr int ABS_ROWS = 4;
__global__ void __launch_bounds__(256)
act_bwd_split_kernel(const float* __restrict__ dy, long long ld_dy, const float* __restrict__ aux, long long ld_aux, int act,
                     __nv_bfloat16* __restrict__ dst, long long ld_dst, long long plane_stride, long long rows, long long cols,
                     float* __restrict__ dbias) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long c0 = (blockIdx.x * 256LL + threadIdx.x) * 4;
    if (c0 >= cols) return;
    const long long r0 = blockIdx.y * (long long)ABS_ROWS;
    const long long r1 = min(rows, r0 + ABS_ROWS);
    const bool vec = (c0 + 4 <= cols) && ((ld_dy & 3) == 0) && ((ld_aux & 3) == 0) && ((ld_dst & 3) == 0) &&
                     ((plane_stride & 3) == 0) && ((((uintptr_t)dy | (uintptr_t)aux) & 15) == 0) && ((((uintptr_t)dst) & 7) == 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = r0; r < r1; ++r) {
        float d[4] = {0.f, 0.f, 0.f, 0.f}, a[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec) {
            const float4 d4 = *reinterpret_cast<const float4*>(dy + r * ld_dy + c0);
            d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
            if (act != YV_ACT_NONE) {
                const float4 a4 = *reinterpret_cast<const float4*>(aux + r * ld_aux + c0);
                a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
            }
        } else {
            for (int j = 0; j < 4 && c0 + j < cols; ++j) {
                d[j] = dy[r * ld_dy + c0 + j];
                if (act != YV_ACT_NONE) a[j] = aux[r * ld_aux + c0 + j];
            }
        }
        __align__(8) __nv_bfloat16 h4[4], l4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (act == YV_ACT_GELU) d[j] *= yv_gelu_grad(a[j]);
            else if (act == YV_ACT_RELU) d[j] = a[j] > 0.f ? d[j] : 0.f;
            acc[j] += d[j];
            yv_split(d[j], h4[j], l4[j]);
        }
        if (vec) {
            *reinterpret_cast<uint2*>(dst + r * ld_dst + c0) = *reinterpret_cast<uint2*>(h4);
            *reinterpret_cast<uint2*>(dst + plane_stride + r * ld_dst + c0) = *reinterpret_cast<uint2*>(l4);
        } else {
            for (int j = 0; j < 4 && c0 + j < cols; ++j) {
                dst[r * ld_dst + c0 + j] = h4[j];
                dst[plane_stride + r * ld_dst + c0 + j] = l4[j];
            }
        }
    }
    if (dbias)
        for (int j = 0; j < 4 && c0 + j < cols; ++j) atomicAdd(dbias + c0 + j, acc[j]);
}

// ------------------------------------------------------------------------------------------------
// losses: one block per row
// ------------------------------------------------------------------------------------------------
__device__ float block_max(float v, float* sh) {
    v = yv_warp_max(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
    __syncthreads();
    return r;
}
__device__ float block_sum(float v, float* sh) {
    v = yv_warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(THREADS)
ce_loss_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target, int cols,
               float* __restrict__ loss_sum, float* __restrict__ count) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float sh[WARPS];
    const long long row = blockIdx.x;
    const long long t = target[row];
    if (t < 0) return;
    if (t >= cols) __trap();           // label outside the vocabulary (ATen raises a device assert here too)
    const float* lr = logits + row * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += THREADS) mx = fmaxf(mx, lr[c]);
    mx = block_max(mx, sh);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += THREADS) s += expf(lr[c] - mx);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) {
        atomicAdd(loss_sum, (mx + logf(s)) - lr[t]);
        atomicAdd(count, 1.f);
    }
}

__global__ void __launch_bounds__(THREADS)
ce_grad_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ target, int cols,
               const float* __restrict__ count, const float* __restrict__ gscale, float* __restrict__ dl32,
               __nv_bfloat16* __restrict__ dlp, long long ld_p, long long plane_stride) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float sh[WARPS];
    const long long row = blockIdx.x;
    const long long t = target[row];
    if (t >= cols) __trap();
    const float* lr = logits + row * ld;
    float mx = 0.f, inv = 0.f, g = 0.f;
    if (t >= 0) {
        mx = -INFINITY;
        for (int c = threadIdx.x; c < cols; c += THREADS) mx = fmaxf(mx, lr[c]);
        mx = block_max(mx, sh);
        float s = 0.f;
        for (int c = threadIdx.x; c < cols; c += THREADS) s += expf(lr[c] - mx);
        s = block_sum(s, sh);
        inv = 1.f / s;
        g = (gscale ? gscale[0] : 1.f) / fmaxf(count[0], 1.f);
    }
    for (int c = threadIdx.x; c < cols; c += THREADS) {
        float d = 0.f;
        if (t >= 0) d = g * (expf(lr[c] - mx) * inv - (c == t ? 1.f : 0.f));
        if (dl32) dl32[row * ld + c] = d;
        if (dlp) {
            __nv_bfloat16 h, l;
            yv_split(d, h, l);
            dlp[row * ld_p + c] = h;
            dlp[plane_stride + row * ld_p + c] = l;
        }
    }
}

__global__ void __launch_bounds__(THREADS)
kl_loss_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target, long long ld_t,
               const long long* __restrict__ mask, int cols, float* __restrict__ loss_sum, float* __restrict__ count) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float sh[WARPS];
    const long long row = blockIdx.x;
    if (mask[row] == 0) return;
    const float* lr = logits + row * ld;
    const float* tr = target + row * ld_t;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += THREADS) mx = fmaxf(mx, lr[c]);
    mx = block_max(mx, sh);
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += THREADS) s += expf(lr[c] - mx);
    s = block_sum(s, sh);
    const float lse = mx + logf(s);
    float acc = 0.f;
    for (int c = threadIdx.x; c < cols; c += THREADS) {
        const float t = tr[c];
        if (t > 0.f) acc += t * (logf(t) - (lr[c] - lse));
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        atomicAdd(loss_sum, acc * (float)mask[row]);
        atomicAdd(count, (float)mask[row]);
    }
}

__global__ void __launch_bounds__(THREADS)
kl_grad_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ target, long long ld_t,
               const long long* __restrict__ mask, int cols, const float* __restrict__ count,
               const float* __restrict__ gscale, float* __restrict__ dl32, __nv_bfloat16* __restrict__ dlp, long long ld_p,
               long long plane_stride) {
    yv_pdl_trigger();
    yv_pdl_wait();
    __shared__ float sh[WARPS];
    const long long row = blockIdx.x;
    const float mk = (float)mask[row];
    const float* lr = logits + row * ld;
    const float* tr = target + row * ld_t;
    float mx = 0.f, inv = 0.f, g = 0.f, tsum = 0.f;
    if (mk != 0.f) {
        mx = -INFINITY;
        for (int c = threadIdx.x; c < cols; c += THREADS) mx = fmaxf(mx, lr[c]);
        mx = block_max(mx, sh);
        float s = 0.f, ts = 0.f;
        for (int c = threadIdx.x; c < cols; c += THREADS) {
            s += expf(lr[c] - mx);
            ts += tr[c];
        }
        s = block_sum(s, sh);
        tsum = block_sum(ts, sh);
        inv = 1.f / s;
        g = mk * (gscale ? gscale[0] : 1.f) / fmaxf(count[0], 1.f);
    }
    for (int c = threadIdx.x; c < cols; c += THREADS) {
        float d = 0.f;
        if (mk != 0.f) d = g * (expf(lr[c] - mx) * inv * tsum - tr[c]);
        if (dl32) dl32[row * ld + c] = d;
        if (dlp) {
            __nv_bfloat16 h, l;
            yv_split(d, h, l);
            dlp[row * ld_p + c] = h;
            dlp[plane_stride + row * ld_p + c] = l;
        }
    }
}

inline cudaStream_t S(yv_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
// ------------------------------------------------------------------------------------------------
// batch masking (SURVEY 8f "next" #3): BERT token masking and ViLBERT region masking on the device.  The uniform draws
// are inputs, so the kernels are deterministic integer / byte work (bit-exact against the reference's CPU functions).
// Thresholds as the reference computes them: Python doubles, cast to float32 by the comparison with a float tensor.
// ------------------------------------------------------------------------------------------------
__device__ __constant__ float kMaskThresh[5] = {(float)0.85, (float)(0.85 + 0.15 * 0.8), (float)(0.85 + 0.15 * 0.9),
                                                (float)(0.85 + 0.15 * 0.1), (float)(0.85 * 0.9)};

__global__ void mask_tokens_kernel(long long* __restrict__ tokens, const unsigned char* __restrict__ mask,
                                   const float* __restrict__ p, const long long* __restrict__ random,
                                   const unsigned char* __restrict__ forced, long long mask_id,
                                   long long* __restrict__ targets, long long n) {
    yv_pdl_trigger();
    yv_pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float pe = p[i] * (mask[i] ? 1.f : 0.f);
        long long t = tokens[i], tg = -1;
        if (forced != nullptr && forced[i]) {        // an action word picked by the host: always [MASK]
            tg = t;
            t = mask_id;
            pe = kMaskThresh[4];
        }
        if (pe >= kMaskThresh[0]) {                  // supervised position: 80 % [MASK] ...
            tg = t;
            t = mask_id;
        }
        if (pe >= kMaskThresh[1]) t = random[i];     // ... 10 % random word ...
        if (pe >= kMaskThresh[2]) t = tg;            // ... 10 % unchanged
        tokens[i] = t;
        targets[i] = tg;
    }
}

// one block per region: targets = probs (supervised) or uniform, 90 % of the supervised regions get zero features
__global__ void __launch_bounds__(256)
mask_regions_kernel(float* __restrict__ features, const float* __restrict__ probs, const long long* __restrict__ mask,
                    const float* __restrict__ p, float* __restrict__ targets, long long* __restrict__ targets_mask,
                    long long rows, int F, int Cc) {
    yv_pdl_trigger();
    yv_pdl_wait();
    const long long r = blockIdx.x;
    if (r >= rows) return;
    const float pe = p[r] * (float)mask[r];
    const bool sup = pe >= kMaskThresh[0];
    const float uni = 1.0f / (float)Cc;
    const float* pr = probs + r * Cc;
    float* tr = targets + r * Cc;
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) tr[c] = sup ? pr[c] : uni;
    if (threadIdx.x == 0) targets_mask[r] = sup ? 1 : 0;
    if (pe >= kMaskThresh[3]) {
        float* fr = features + r * (long long)F;
        for (int c = threadIdx.x; c < F; c += blockDim.x) fr[c] = 0.f;
    }
}

// YVB200_SOFTMAX=scalar selects the scalar softmax kernels for every shape (kept for A/B timing)
const bool g_softmax_vec = []() { const char* e = getenv("YVB200_SOFTMAX"); return !(e && e[0] == 's'); }();
// YVB200_LN=block selects the older block-per-4-rows LayerNorm kernels (kept for A/B timing)
const bool g_ln_warp = []() { const char* e = getenv("YVB200_LN"); return !(e && e[0] == 'b'); }();

inline int grid_for(long long n, int per_block, int cap = 148 * 16) {
    long long g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace

#define YV_LAUNCHED()           \
    YV_CUDA(cudaGetLastError()); \
    yv_count_launch();           \
    return 0

extern "C" int yv_split_planes(const float* src, int64_t ld_src, void* planes, int64_t ld_dst, int64_t plane_stride,
                               int64_t rows, int64_t cols, yv_stream_t stream) {
    YV_CHECK(src && planes && rows > 0 && cols > 0, "yv_split_planes: bad arguments");
    YV_CUDA(yv_launch(split_planes_kernel, dim3(grid_for(rows * cols, 256 * 4)), dim3(256), 0, S(stream), 
        src, ld_src, reinterpret_cast<__nv_bfloat16*>(planes), ld_dst, plane_stride, rows, cols));
    YV_LAUNCHED();
}

extern "C" int yv_split_multi(const YvSplitSeg* segs_dev, int32_t nseg, int64_t total_blocks, void* planes,
                              int64_t plane_stride, yv_stream_t stream) {
    YV_CHECK(segs_dev && planes && nseg > 0 && total_blocks > 0, "yv_split_multi: bad arguments");
    YV_CHECK(total_blocks < 2147483647LL, "yv_split_multi: too many blocks");
    YV_CUDA(yv_launch(split_multi_kernel, dim3((unsigned)total_blocks), dim3(256), 0, S(stream), segs_dev, nseg,
                                                                      reinterpret_cast<__nv_bfloat16*>(planes), plane_stride));
    YV_LAUNCHED();
}

extern "C" int yv_adamw_multi(const YvAdamSeg* segs_dev, int32_t nseg, int64_t total_blocks, const float* hyper_dev,
                              yv_stream_t stream) {
    YV_CHECK(segs_dev && hyper_dev && nseg > 0 && total_blocks > 0, "yv_adamw_multi: bad arguments");
    YV_CHECK(total_blocks < 2147483647LL, "yv_adamw_multi: too many blocks");
    YV_CUDA(yv_launch(adamw_multi_kernel, dim3((unsigned)total_blocks), dim3(256), 0, S(stream), segs_dev, nseg, hyper_dev));
    YV_LAUNCHED();
}

extern "C" int yv_rng_advance(uint64_t* rng, yv_stream_t stream) {
    YV_CHECK(rng, "yv_rng_advance: NULL state");
    YV_CUDA(yv_launch(rng_advance_kernel, dim3(1), dim3(1), 0, S(stream), reinterpret_cast<unsigned long long*>(rng)));
    YV_LAUNCHED();
}

extern "C" int yv_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32, void* y_planes,
                                int64_t plane_stride, float* stats, int64_t M, int32_t C, float drop_p, uint32_t drop_site,
                                const uint64_t* rng, yv_stream_t stream) {
    YV_CHECK(x && gamma && beta && M > 0 && C > 0 && (y32 || y_planes), "yv_layernorm_fwd: bad arguments");
    YV_CHECK(C <= 1024 && C % 4 == 0, "yv_layernorm_fwd: hidden size %d must be a multiple of 4 and <= 1024", C);
    YV_CHECK(((((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)y32) & 15) == 0) &&
                 ((((uintptr_t)y_planes) & 7) == 0) && (plane_stride % 4 == 0),
             "yv_layernorm_fwd: pointers must be 16-byte aligned");
    if (g_ln_warp) {
        const dim3 grid((unsigned)((M + LNW_WARPS - 1) / LNW_WARPS)), block(LNW_WARPS * 32);
        __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y_planes);
        const unsigned long long* r = reinterpret_cast<const unsigned long long*>(rng);
        const int nv = (C + 127) / 128;
#define YV_LNF(NV) YV_CUDA(yv_launch(layernorm_fwd_warp_kernel<NV>, grid, block, 0, S(stream), x, gamma, beta, eps, y32, yp, \
                                     plane_stride, stats, M, C, drop_p, drop_site, r))
        if (nv <= 2) YV_LNF(2); else if (nv <= 4) YV_LNF(4); else if (nv <= 6) YV_LNF(6); else YV_LNF(8);
#undef YV_LNF
        YV_LAUNCHED();
    }
    const int threads = ((C / 4 + 31) / 32) * 32;
    YV_CUDA(yv_launch(layernorm_fwd_kernel, dim3((unsigned)((M + LNF_ROWS - 1) / LNF_ROWS)), dim3(threads), 0, S(stream), x, gamma,
                      beta, eps, y32, reinterpret_cast<__nv_bfloat16*>(y_planes), plane_stride, stats, M, C, drop_p, drop_site,
                      reinterpret_cast<const unsigned long long*>(rng)));
    YV_LAUNCHED();
}

extern "C" int yv_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* stats, float post_drop_p,
                                uint32_t post_drop_site, const float* dx_add, float* dx32, void* dx_planes,
                                int64_t plane_stride, float pre_drop_p, uint32_t pre_drop_site, const uint64_t* rng,
                                float* dgamma, float* dbeta, float* dbias, int64_t M, int32_t C, yv_stream_t stream) {
    YV_CHECK(dy && x && gamma && stats && M > 0 && C > 0, "yv_layernorm_bwd: bad arguments");
    YV_CHECK(C <= 1024 && C % 4 == 0, "yv_layernorm_bwd: hidden size %d must be a multiple of 4 and <= 1024", C);
    YV_CHECK(((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)gamma | (uintptr_t)dx_add | (uintptr_t)dx32 | (uintptr_t)dgamma |
                (uintptr_t)dbeta | (uintptr_t)dbias) & 15) == 0) && ((((uintptr_t)dx_planes) & 7) == 0) &&
                 (plane_stride % 4 == 0),
             "yv_layernorm_bwd: pointers must be 16-byte aligned");
    if (g_ln_warp) {
        // at most one block per SM: every block ends with 3 * C / 4 vector reductions into dgamma / dbeta / dbias
        const dim3 grid((unsigned)grid_for(M, LNW_WARPS, 148)), block(LNW_WARPS * 32);
        __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dx_planes);
        const unsigned long long* r = reinterpret_cast<const unsigned long long*>(rng);
        const int nv = (C + 127) / 128;
#define YV_LNB(NV) YV_CUDA(yv_launch(layernorm_bwd_warp_kernel<NV, true>, grid, block, 0, S(stream), dy, x, gamma, stats, post_drop_p, \
                                     post_drop_site, dx_add, dx32, dp, plane_stride, pre_drop_p, pre_drop_site, r, dgamma, dbeta, \
                                     dbias, M, C))
        if (nv <= 2) YV_LNB(2); else if (nv <= 4) YV_LNB(4); else if (nv <= 6) YV_LNB(6); else YV_LNB(8);
#undef YV_LNB
        YV_LAUNCHED();
    }
    const int threads = ((C / 4 + 31) / 32) * 32;
    const int grid = grid_for(M, LNB_ROWS, 148 * 4);
    YV_CUDA(yv_launch(layernorm_bwd_kernel, dim3(grid), dim3(threads), 0, S(stream), dy, x, gamma, stats, post_drop_p, post_drop_site, dx_add, dx32,
                                                          reinterpret_cast<__nv_bfloat16*>(dx_planes), plane_stride, pre_drop_p,
                                                          pre_drop_site, reinterpret_cast<const unsigned long long*>(rng), dgamma,
                                                          dbeta, dbias, M, C));
    YV_LAUNCHED();
}

extern "C" int yv_layernorm_bwd_dx(const float* dy, const float* x, const float* gamma, const float* stats,
                                   const float* dx_add, float* dx32, void* dx_planes, int64_t plane_stride, float pre_drop_p,
                                   uint32_t pre_drop_site, const uint64_t* rng, int64_t M, int32_t C, yv_stream_t stream) {
    YV_CHECK(dy && x && gamma && stats && M > 0 && C > 0, "yv_layernorm_bwd_dx: bad arguments");
    YV_CHECK(C <= 1024 && C % 4 == 0, "yv_layernorm_bwd_dx: hidden size %d must be a multiple of 4 and <= 1024", C);
    YV_CHECK(((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)gamma | (uintptr_t)dx_add | (uintptr_t)dx32) & 15) == 0) &&
                 ((((uintptr_t)dx_planes) & 7) == 0) && (plane_stride % 4 == 0),
             "yv_layernorm_bwd_dx: pointers must be 16-byte aligned");
    const dim3 grid((unsigned)grid_for(M, LNW_WARPS, 148 * 4)), block(LNW_WARPS * 32);
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dx_planes);
    const unsigned long long* r = reinterpret_cast<const unsigned long long*>(rng);
    const int nv = (C + 127) / 128;
#define YV_LNB(NV) YV_CUDA(yv_launch(layernorm_bwd_warp_kernel<NV, false>, grid, block, 0, S(stream), dy, x, gamma, stats, 0.f, \
                                     0u, dx_add, dx32, dp, plane_stride, pre_drop_p, pre_drop_site, r, (float*)nullptr, \
                                     (float*)nullptr, (float*)nullptr, M, C))
    if (nv <= 2) YV_LNB(2); else if (nv <= 4) YV_LNB(4); else if (nv <= 6) YV_LNB(6); else YV_LNB(8);
#undef YV_LNB
    YV_LAUNCHED();
}

extern "C" int yv_layernorm_bwd_cols(const float* dy, const float* x, const float* stats, float* dgamma, float* dbeta,
                                     int64_t M, int32_t C, yv_stream_t stream) {
    YV_CHECK(dy && x && stats && dgamma && dbeta && M > 0 && C > 0, "yv_layernorm_bwd_cols: bad arguments");
    YV_CHECK(C % 4 == 0 && ((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dgamma | (uintptr_t)dbeta) & 15) == 0),
             "yv_layernorm_bwd_cols: hidden size must be a multiple of 4 and pointers 16-byte aligned");
    const int groups = (C + 127) / 128;
    int splits = (2 * 148 + groups - 1) / groups;
    if ((long long)splits * LNW_WARPS > M) splits = (int)((M + LNW_WARPS - 1) / LNW_WARPS);
    const int rows_per_split = (int)((M + splits - 1) / splits);
    splits = (int)((M + rows_per_split - 1) / rows_per_split);
    YV_CUDA(yv_launch(layernorm_bwd_cols_kernel, dim3((unsigned)groups, (unsigned)splits), dim3(LNW_WARPS * 32), 0, S(stream),
                      dy, x, stats, dgamma, dbeta, M, C, rows_per_split));
    YV_LAUNCHED();
}

extern "C" int yv_softmax_fwd(float* s, int64_t ld_s, const float* mask, int64_t rows, int32_t cols, int64_t rows_per_pair,
                              float scale, void* p_planes, int64_t ld_p, int64_t plane_stride, float drop_p,
                              uint32_t drop_site, const uint64_t* rng, yv_stream_t stream) {
    YV_CHECK(s && p_planes && rows > 0 && cols > 0 && rows_per_pair > 0, "yv_softmax_fwd: bad arguments");
    YV_CHECK(cols <= 2048, "yv_softmax_fwd: %d keys per row not supported (max 2048)", cols);
    const dim3 grid((unsigned)((rows + WARPS - 1) / WARPS));
    auto* pp = reinterpret_cast<__nv_bfloat16*>(p_planes);
    auto* rp = reinterpret_cast<const unsigned long long*>(rng);
    const bool vec = g_softmax_vec && cols % 4 == 0 && ld_s % 4 == 0 && ld_p % 4 == 0 && plane_stride % 4 == 0 &&
                     (((uintptr_t)s | (uintptr_t)mask) & 15) == 0 && (((uintptr_t)p_planes) & 7) == 0;
    if (vec) {
#define YV_SMFV(NV) YV_CUDA(yv_launch(softmax_fwd_vec_kernel<NV>, grid, dim3(THREADS), 0, S(stream), s, ld_s, mask, rows, cols, \
                                      rows_per_pair, scale, pp, ld_p, plane_stride, drop_p, drop_site, rp))
        if (cols <= 128) YV_SMFV(1);
        else if (cols <= 384) YV_SMFV(3);
        else if (cols <= 640) YV_SMFV(5);
        else if (cols <= 1152) YV_SMFV(9);
        else YV_SMFV(16);
#undef YV_SMFV
        YV_LAUNCHED();
    }
#define YV_SMF(PL) YV_CUDA(yv_launch(softmax_fwd_kernel<PL>, grid, dim3(THREADS), 0, S(stream), s, ld_s, mask, rows, cols, \
                                     rows_per_pair, scale, pp, ld_p, plane_stride, drop_p, drop_site, rp))
    if (cols <= 96) YV_SMF(3);
    else if (cols <= 288) YV_SMF(9);
    else if (cols <= 576) YV_SMF(18);
    else if (cols <= 1152) YV_SMF(36);
    else YV_SMF(64);
#undef YV_SMF
    YV_LAUNCHED();
}

extern "C" int yv_softmax_bwd(const float* p, const float* dpd, int64_t ld_s, int64_t rows, int32_t cols, float scale,
                              void* ds_planes, int64_t ld_p, int64_t plane_stride, float drop_p, uint32_t drop_site,
                              const uint64_t* rng, yv_stream_t stream) {
    YV_CHECK(p && dpd && ds_planes && rows > 0 && cols > 0, "yv_softmax_bwd: bad arguments");
    YV_CHECK(cols <= 2048, "yv_softmax_bwd: %d keys per row not supported (max 2048)", cols);
    const dim3 grid((unsigned)((rows + WARPS - 1) / WARPS));
    auto* dp = reinterpret_cast<__nv_bfloat16*>(ds_planes);
    auto* rp = reinterpret_cast<const unsigned long long*>(rng);
    const bool vec = g_softmax_vec && cols % 4 == 0 && ld_s % 4 == 0 && ld_p % 4 == 0 && plane_stride % 4 == 0 &&
                     (((uintptr_t)p | (uintptr_t)dpd) & 15) == 0 && (((uintptr_t)ds_planes) & 7) == 0;
    if (vec) {
#define YV_SMBV(NV) YV_CUDA(yv_launch(softmax_bwd_vec_kernel<NV>, grid, dim3(THREADS), 0, S(stream), p, dpd, ld_s, rows, cols, \
                                      scale, dp, ld_p, plane_stride, drop_p, drop_site, rp))
        if (cols <= 128) YV_SMBV(1);
        else if (cols <= 384) YV_SMBV(3);
        else if (cols <= 640) YV_SMBV(5);
        else if (cols <= 1152) YV_SMBV(9);
        else YV_SMBV(16);
#undef YV_SMBV
        YV_LAUNCHED();
    }
#define YV_SMB(PL) YV_CUDA(yv_launch(softmax_bwd_kernel<PL>, grid, dim3(THREADS), 0, S(stream), p, dpd, ld_s, rows, cols, scale, \
                                     dp, ld_p, plane_stride, drop_p, drop_site, rp))
    if (cols <= 96) YV_SMB(3);
    else if (cols <= 288) YV_SMB(9);
    else if (cols <= 576) YV_SMB(18);
    else if (cols <= 1152) YV_SMB(36);
    else YV_SMB(64);
#undef YV_SMB
    YV_LAUNCHED();
}

extern "C" int yv_embed_text_fwd(const int64_t* tok, const int64_t* seg, const float* word, const float* pos,
                                 const float* type, float* out, int64_t M, int32_t T, int32_t H, yv_stream_t stream) {
    YV_CHECK(tok && seg && word && pos && type && out && M > 0, "yv_embed_text_fwd: bad arguments");
    YV_CUDA(yv_launch(embed_text_fwd_kernel, dim3((unsigned)M), dim3(256), 0, S(stream), reinterpret_cast<const long long*>(tok),
                                                              reinterpret_cast<const long long*>(seg), word, pos, type, out, M, T, H));
    YV_LAUNCHED();
}

extern "C" int yv_embed_text_bwd(const int64_t* tok, const int64_t* seg, const float* dout, float* dword, float* dpos,
                                 float* dtype_, int64_t M, int32_t T, int32_t H, int32_t padding_idx, yv_stream_t stream) {
    YV_CHECK(tok && seg && dout && M > 0, "yv_embed_text_bwd: bad arguments");
    YV_CUDA(yv_launch(embed_text_bwd_kernel, dim3((unsigned)M), dim3(256), 0, S(stream), reinterpret_cast<const long long*>(tok),
                                                              reinterpret_cast<const long long*>(seg), dout, dword, dpos, dtype_, M,
                                                              T, H, padding_idx));
    YV_LAUNCHED();
}

extern "C" int yv_embed_loc_fwd(const float* loc, const float* w5, const float* b5, const float* w4, const float* b4,
                                const float* w2, const float* b2, const float* seq, float* out, int64_t M, int32_t H,
                                yv_stream_t stream) {
    YV_CHECK(loc && w5 && b5 && w4 && b4 && w2 && b2 && seq && out && M > 0, "yv_embed_loc_fwd: bad arguments");
    YV_CUDA(yv_launch(embed_loc_fwd_kernel, dim3((unsigned)M), dim3(256), 0, S(stream), loc, w5, b5, w4, b4, w2, b2, seq, out, M, H));
    YV_LAUNCHED();
}

extern "C" int yv_embed_loc_bwd(const float* loc, const float* dout, float* dw5, float* db5, float* dw4, float* db4, float* dw2,
                                float* db2, float* dseq, int64_t M, int32_t H, yv_stream_t stream) {
    YV_CHECK(loc && dout && dw5 && db5 && dw4 && db4 && dw2 && db2 && dseq && M > 0, "yv_embed_loc_bwd: bad arguments");
    YV_CUDA(yv_launch(embed_loc_bwd_kernel, dim3((unsigned)((M + LOC_ROWS - 1) / LOC_ROWS), LOC_SPLIT), dim3(256), 0, S(stream), loc, dout, dw5, db5, dw4, db4, dw2,
                                                                                          db2, dseq, M, H));
    YV_LAUNCHED();
}

// ------------------------------------------------------------------------------------------- exchange: chunk mean
// out[i] = (sum over ranks r = 0..world-1, in that order, of x_r[i]) / world, where x_rank is `own` (also the output)
// and x_r of a peer is the row of `stage` the copy engines filled for it: peer (rank - step) mod world sits in row step-1.
// Streaming, 16 bytes per lane, a bounded grid (the backward pass is running on the same SMs).
__global__ void __launch_bounds__(512) mean_chunks_kernel(float* __restrict__ own, const float* __restrict__ stage,
                                                          long long stage_stride, int world, int rank, long long n4) {
    yv_pdl_wait();
    const float inv = 1.0f / (float)world;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float* src = r == rank ? own : stage + (long long)(((rank - r + world) % world) - 1) * stage_stride;
            const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + i);
            if (r == 0) acc = v;
            else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        }
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
        reinterpret_cast<float4*>(own)[i] = acc;
    }
}

extern "C" int yv_mean_chunks(float* own, const float* stage, int64_t stage_stride, int32_t world, int32_t rank, int64_t n,
                              yv_stream_t stream) {
    YV_CHECK(own && world >= 1 && rank >= 0 && rank < world && n > 0, "yv_mean_chunks: bad arguments");
    YV_CHECK(world == 1 || stage, "yv_mean_chunks: no staging rows");
    YV_CHECK((n & 3) == 0 && (stage_stride & 3) == 0 && (((uintptr_t)own | (uintptr_t)stage) & 15) == 0,
             "yv_mean_chunks: chunks must be 16-byte aligned multiples of 4 elements");
    const long long n4 = n / 4;
    const unsigned blocks = (unsigned)(n4 < 74 * 512 ? (n4 + 511) / 512 : 74);
    YV_CUDA(yv_launch(mean_chunks_kernel, dim3(blocks), dim3(512), 0, S(stream), own, stage, (long long)stage_stride,
                      (int)world, (int)rank, n4));
    YV_LAUNCHED();
}

extern "C" int yv_colsum(const float* x, int64_t ld, int64_t rows, int32_t cols, float* out, int32_t accumulate,
                         yv_stream_t stream) {
    YV_CHECK(x && out && rows > 0 && cols > 0, "yv_colsum: bad arguments");
    if (!accumulate) YV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, S(stream)));
    dim3 grid((cols + 127) / 128, (unsigned)((rows + CS_ROWS - 1) / CS_ROWS));
    YV_CUDA(yv_launch(colsum_kernel, dim3(grid), dim3(128), 0, S(stream), x, ld, rows, cols, out));
    YV_LAUNCHED();
}

extern "C" int yv_colsum_planes(const void* planes, int64_t ld, int64_t plane_stride, int64_t rows, int32_t cols, float* out,
                                int32_t accumulate, yv_stream_t stream) {
    YV_CHECK(planes && out && rows > 0 && cols > 0, "yv_colsum_planes: bad arguments");
    if (!accumulate) YV_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, S(stream)));
    dim3 grid((cols + 1023) / 1024, (unsigned)((rows + CSP_ROWS - 1) / CSP_ROWS));
    YV_CUDA(yv_launch(colsum_planes_kernel, dim3(grid), dim3(128), 0, S(stream), reinterpret_cast<const __nv_bfloat16*>(planes), ld, plane_stride, rows,
                                                      cols, out));
    YV_LAUNCHED();
}

extern "C" int yv_act_bwd_split(const float* dy, int64_t ld_dy, const float* aux, int64_t ld_aux, int32_t act, void* planes,
                                int64_t ld_dst, int64_t plane_stride, int64_t rows, int64_t cols, float* dbias,
                                yv_stream_t stream) {
    YV_CHECK(dy && planes && rows > 0 && cols > 0, "yv_act_bwd_split: bad arguments");
    YV_CHECK(act == YV_ACT_NONE || aux, "yv_act_bwd_split: act %d needs aux", act);
    if (dbias) YV_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * cols, S(stream)));
    dim3 grid((unsigned)((cols + 1023) / 1024), (unsigned)((rows + ABS_ROWS - 1) / ABS_ROWS));
    YV_CUDA(yv_launch(act_bwd_split_kernel, dim3(grid), dim3(256), 0, S(stream), dy, ld_dy, aux, ld_aux, act, reinterpret_cast<__nv_bfloat16*>(planes), ld_dst,
                                                      plane_stride, rows, cols, dbias));
    YV_LAUNCHED();
}

extern "C" int yv_ce_loss(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int32_t cols, float* loss_sum,
                          float* count, yv_stream_t stream) {
    YV_CHECK(logits && target && loss_sum && count && rows > 0 && cols > 0, "yv_ce_loss: bad arguments");
    YV_CUDA(yv_launch(ce_loss_kernel, dim3((unsigned)rows), dim3(THREADS), 0, S(stream), logits, ld, reinterpret_cast<const long long*>(target), cols,
                                                             loss_sum, count));
    YV_LAUNCHED();
}

extern "C" int yv_ce_grad(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int32_t cols, const float* count,
                          const float* gscale, float* dl32, void* dl_planes, int64_t ld_p, int64_t plane_stride,
                          yv_stream_t stream) {
    YV_CHECK(logits && target && count && (dl32 || dl_planes) && rows > 0 && cols > 0, "yv_ce_grad: bad arguments");
    YV_CUDA(yv_launch(ce_grad_kernel, dim3((unsigned)rows), dim3(THREADS), 0, S(stream), logits, ld, reinterpret_cast<const long long*>(target), cols, count,
                                                             gscale, dl32, reinterpret_cast<__nv_bfloat16*>(dl_planes), ld_p,
                                                             plane_stride));
    YV_LAUNCHED();
}

extern "C" int yv_kl_loss(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* mask, int64_t rows,
                          int32_t cols, float* loss_sum, float* count, yv_stream_t stream) {
    YV_CHECK(logits && target && mask && loss_sum && count && rows > 0 && cols > 0, "yv_kl_loss: bad arguments");
    YV_CUDA(yv_launch(kl_loss_kernel, dim3((unsigned)rows), dim3(THREADS), 0, S(stream), logits, ld, target, ld_t, reinterpret_cast<const long long*>(mask),
                                                             cols, loss_sum, count));
    YV_LAUNCHED();
}

extern "C" int yv_kl_grad(const float* logits, int64_t ld, const float* target, int64_t ld_t, const int64_t* mask, int64_t rows,
                          int32_t cols, const float* count, const float* gscale, float* dl32, void* dl_planes, int64_t ld_p,
                          int64_t plane_stride, yv_stream_t stream) {
    YV_CHECK(logits && target && mask && count && (dl32 || dl_planes) && rows > 0 && cols > 0, "yv_kl_grad: bad arguments");
    YV_CUDA(yv_launch(kl_grad_kernel, dim3((unsigned)rows), dim3(THREADS), 0, S(stream), logits, ld, target, ld_t, reinterpret_cast<const long long*>(mask),
                                                             cols, count, gscale, dl32, reinterpret_cast<__nv_bfloat16*>(dl_planes),
                                                             ld_p, plane_stride));
    YV_LAUNCHED();
}

extern "C" int yv_mask_tokens(int64_t* tokens, const uint8_t* mask, const float* p, const int64_t* random_tokens,
                              const uint8_t* forced, int64_t mask_id, int64_t* targets, int64_t n, yv_stream_t stream) {
    YV_CHECK(tokens && mask && p && random_tokens && targets && n > 0, "yv_mask_tokens: bad arguments");
    YV_CUDA(yv_launch(mask_tokens_kernel, dim3(grid_for(n, 256)), dim3(256), 0, S(stream),
                      reinterpret_cast<long long*>(tokens), mask, p, reinterpret_cast<const long long*>(random_tokens), forced,
                      (long long)mask_id, reinterpret_cast<long long*>(targets), (long long)n));
    YV_LAUNCHED();
}

extern "C" int yv_mask_regions(float* features, const float* probs, const int64_t* mask, const float* p, float* targets,
                               int64_t* targets_mask, int64_t rows, int32_t F, int32_t Cc, yv_stream_t stream) {
    YV_CHECK(features && probs && mask && p && targets && targets_mask && rows > 0 && F > 0 && Cc > 0,
             "yv_mask_regions: bad arguments");
    YV_CHECK(rows < 2147483647LL, "yv_mask_regions: too many regions");
    YV_CUDA(yv_launch(mask_regions_kernel, dim3((unsigned)rows), dim3(256), 0, S(stream), features, probs,
                      reinterpret_cast<const long long*>(mask), p, targets, reinterpret_cast<long long*>(targets_mask),
                      (long long)rows, (int)F, (int)Cc));
    YV_LAUNCHED();
}
