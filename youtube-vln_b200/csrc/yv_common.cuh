// Shared device helpers for the yvb200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define YV_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// error plumbing (host): every C-ABI entry returns 0 / non-zero and leaves a message behind
// ---------------------------------------------------------------------------------------------
void yv_set_error(const char* fmt, ...);
#define YV_CHECK(cond, ...)                \
    do {                                   \
        if (!(cond)) {                     \
            yv_set_error(__VA_ARGS__);     \
            return 1;                      \
        }                                  \
    } while (0)
#define YV_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            yv_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                         __LINE__);                                                         \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: stateless hash of (element index, site key, step key)
//   rng[0] = seed, rng[1] = step counter (advanced by yv_rng_advance once per training step, so a
//   captured CUDA graph draws fresh masks on every replay).  Backward recomputes the same mask.
// ---------------------------------------------------------------------------------------------
YV_DEVINL uint32_t yv_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

struct YvDrop {
    uint32_t k0, k1, thresh;
    float scale;  // 1/(1-p); thresh == 0 -> dropout disabled
};

YV_DEVINL YvDrop yv_drop_make(const unsigned long long* rng, uint32_t site, float p) {
    YvDrop d;
    d.thresh = 0;
    d.scale = 1.f;
    d.k0 = d.k1 = 0;
    if (p > 0.f && rng != nullptr) {
        unsigned long long seed = rng[0], step = rng[1];
        d.k0 = yv_mix32((uint32_t)seed ^ (site * 0x9E3779B1U));
        d.k1 = yv_mix32((uint32_t)(seed >> 32) + (uint32_t)step * 0x85EBCA77U + (uint32_t)(step >> 32));
        double t = (double)p * 4294967296.0;
        d.thresh = t >= 4294967295.0 ? 0xFFFFFFFFU : (uint32_t)t;
        d.scale = 1.f / (1.f - p);
    }
    return d;
}

// multiplier applied to element `idx`: 0 (dropped) or 1/(1-p) (kept).  One avalanche round keyed twice: the site key
// enters before the first multiply, the step key between the two multiplies (10 integer ops per element; the fused
// attention kernels hash every probability twice per step, forward and backward).
YV_DEVINL uint32_t yv_drop_hash(const YvDrop& d, uint32_t idx) {
    uint32_t x = idx ^ d.k0;
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x += d.k1;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
YV_DEVINL float yv_drop_mul(const YvDrop& d, uint32_t idx) {
    if (d.thresh == 0) return 1.f;
    return yv_drop_hash(d, idx) >= d.thresh ? d.scale : 0.f;
}

// two fp32 values -> packed bf16x2 hi and lo words (element a in the low half): x ~= hi + lo
YV_DEVINL void yv_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
YV_DEVINL float yv_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
YV_DEVINL float yv_gelu_grad(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// split an fp32 value into a bf16 "hi" plane and a bf16 "lo" residual plane (x ~= hi + lo)
YV_DEVINL void yv_split(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

YV_DEVINL float yv_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
YV_DEVINL float yv_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel lets its successor start launching at once and waits for
// its predecessors (completion + memory visibility) before touching global memory
// ---------------------------------------------------------------------------------------------
YV_DEVINL void yv_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
YV_DEVINL void yv_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool yv_pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t yv_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = yv_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
