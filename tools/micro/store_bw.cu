// Developer microbenchmark: per-SM global store throughput on B200 for the access patterns an epilogue can use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/store_bw.cu -o tools/bin/store_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ long long g_cyc[8];

// mode 0: each warp store = 4 rows x 128 B (float4 per lane), rows `ld` floats apart   (current epilogue, f32)
// mode 1: each warp store = 1 row x 512 B contiguous (float4 per lane)
// mode 2: each warp store = 4 rows x 64 B (uint2 per lane)                              (current epilogue, planes)
// mode 3: bulk async copy smem -> global, 4 KB per instruction (one thread per warp)
// mode 4: each warp store = 2 rows x 256 B
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long ld, int iters, int nctas_timed) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* base = out + (long long)blockIdx.x * 128 * ld;      // this CTA owns 128 rows x 128 cols, repeatedly
    for (int i = threadIdx.x; i < 8192; i += 256) reinterpret_cast<float*>(smem)[i] = (float)i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float* tile = base + (long long)(it & 7) * 128;         // 8 column blocks of 128 floats
        float4 v = make_float4(1.f * it, 2.f, 3.f, 4.f);
        if (MODE == 0) {
            // warp w: rows 16w..16w+15 (4 iterations of 4 rows) x 2 column halves of 32 floats... cover 128x128 tile
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    int row = warp * 16 + i * 4 + (lane >> 3);
                    *reinterpret_cast<float4*>(tile + (long long)row * ld + c * 32 + (lane & 7) * 4) = v;
                }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int row = warp * 16 + i;
                *reinterpret_cast<float4*>(tile + (long long)row * ld + lane * 4) = v;
            }
        } else if (MODE == 2) {
            // bf16 planes: 128x128 bf16 x 2 planes = same 64 KB; 4 rows x 64 B per warp store
            uint16_t* pt = reinterpret_cast<uint16_t*>(tile);
            for (int pl = 0; pl < 2; ++pl)
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int row = warp * 16 + i * 4 + (lane >> 3);
                        *reinterpret_cast<uint2*>(pt + ((long long)row * ld + pl * 64) * 2 + c * 32 + (lane & 7) * 4) =
                            make_uint2(it, lane);
                    }
        } else if (MODE == 3) {
            if (lane == 0) {
                // 16 rows of 512 B per warp -> 16 bulk copies of 512 B each (rows are not contiguous in global)
                for (int i = 0; i < 16; ++i) {
                    int row = warp * 16 + i;
                    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem + (row & 63) * 512);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(tile + (long long)row * ld), "r"(s) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int row = warp * 16 + i * 2 + (lane >> 4);
                *reinterpret_cast<float4*>(tile + (long long)row * ld + (lane & 15) * 4) = v;
                *reinterpret_cast<float4*>(tile + (long long)row * ld + 64 + (lane & 15) * 4) = v;
            }
        }
    }
    if (MODE == 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) g_cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long ld, int grid) {
    const int iters = 64;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE><<<grid, 256, 65536>>>(out, ld, iters, grid);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256, 65536>>>(out, ld, iters, grid);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpyFromSymbol(&cyc, g_cyc, 8);
    double bytes = 65536.0 * iters;
    printf("%-34s grid %3d: %6.1f B/clk per SM (CTA 0), %7.1f GB/s chip (event), err=%s\n", name, grid, bytes / cyc,
           bytes * grid / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long ld = 1024;
    float* out; cudaMalloc(&out, 148ll * 128 * ld * 4);
    for (int grid : {1, 148}) {
        run<0>("4 rows x 128 B float4", out, ld, grid);
        run<1>("1 row x 512 B float4", out, ld, grid);
        run<4>("2 rows x 256 B float4", out, ld, grid);
        run<2>("4 rows x 64 B uint2 (planes)", out, ld, grid);
        run<3>("bulk smem->global 512 B rows", out, ld, grid);
    }
    return 0;
}
