// ubench_pipes.cu -- three questions round 2 needed answered on this machine (sm_100a):
//   1. does SHFL share the LSU data path with LDS / LDG?   (lds / shfl / lds+shfl: is the mix max() or sum()?)
//   2. does a TMA bulk reduce of a 128-byte row (cp.reduce.async.bulk.global.shared::cta.add.f32) get through the L2's
//      atomic units faster than 8 lanes x red.global.add.v4.f32?
//   3. what do packed bf16 reds (red.global.add.noftz.v4.bf16x2: a 64-byte bf16 row per 8 lanes) cost next to fp32 ones?
// One JSON line per case.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// ---- 1. LDS.128 (four random 128-byte rows per warp instruction) and SHFL.IDX, alone and interleaved ----
// n_lds LDS.128 and n_shfl SHFL per inner step; rows from a 32 KB window.
template <int N_LDS, int N_SHFL>
__global__ void __launch_bounds__(kThreads) pipe_mix(int iters, float *sink, long long *cycles) {
    __shared__ __align__(16) float smem[8192];
    for (int i = threadIdx.x; i < 8192; i += kThreads) smem[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, grp = lane >> 3;
    uint32_t seed = (blockIdx.x * kThreads + threadIdx.x / 8 * 8 + grp) * 2654435761u + 7u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float s0 = (float)lane, s1 = 1.f, s2 = 2.f, s3 = 3.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (N_LDS) {
#pragma unroll
                for (int k = 0; k < N_LDS; ++k) {
                    const uint32_t off = (lcg(seed) & 255u) * 128u + (lane & 7) * 16;
                    const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(smem) + off);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            if (N_SHFL) {
#pragma unroll
                for (int k = 0; k < N_SHFL; k += 4) {      // four independent chains
                    s0 += __shfl_sync(0xffffffffu, s0, (lane + 1 + u) & 31);
                    s1 += __shfl_sync(0xffffffffu, s1, (lane + 2 + u) & 31);
                    s2 += __shfl_sync(0xffffffffu, s2, (lane + 3 + u) & 31);
                    s3 += __shfl_sync(0xffffffffu, s3, (lane + 4 + u) & 31);
                }
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x + acc.y + acc.z + acc.w + s0 + s1 + s2 + s3 == 123.456f) sink[0] = acc.x;
}

// ---- 2 / 3. reds of whole rows into a window of `rows` rows (all CTAs the same window: L2) ----
enum RedMode { RED_V4_F32 = 0, BULK_F32 = 1, RED_V4_BF16X2 = 2, BULK_BF16 = 3 };

template <int MODE>
__global__ void __launch_bounds__(kThreads) red_rows(char *buf, int rows, int iters, long long *cycles) {
    __shared__ __align__(128) float src[kThreads / 32][4][32];      // one 128-byte source row per lane group
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 3, cl = lane & 7;
    src[warp][grp][cl * 4 + 0] = 1.f; src[warp][grp][cl * 4 + 1] = 2.f; src[warp][grp][cl * 4 + 2] = 3.f; src[warp][grp][cl * 4 + 3] = 4.f;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the bulk (async proxy) reads
    const uint32_t mask = rows - 1;
    uint32_t seed = ((blockIdx.x * (kThreads / 32) + warp) * 8 + grp) * 2654435761u + 99u;
    const int row_bytes = (MODE == RED_V4_BF16X2 || MODE == BULK_BF16) ? 64 : 128;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const size_t off = (size_t)(lcg(seed) & mask) * row_bytes;
            if (MODE == RED_V4_F32) {
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                             :: "l"(buf + off + cl * 16), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
            } else if (MODE == RED_V4_BF16X2) {
                // 8 lanes x 8 bytes = one 64-byte bf16 row: v2 of packed pairs (4 channels per lane, like the fp32 kernels)
                asm volatile("red.relaxed.gpu.global.add.noftz.v2.bf16x2 [%0], {%1,%2};"
                             :: "l"(buf + off + cl * 8), "r"(0x3f803f80u), "r"(0x40004000u) : "memory");
            } else if (MODE == BULK_F32) {
                if (cl == 0) {      // one lane per group issues the whole row
                    const uint32_t s = (uint32_t)__cvta_generic_to_shared(&src[warp][grp][0]);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;"
                                 :: "l"(buf + off), "r"(s) : "memory");
                }
            } else if (MODE == BULK_BF16) {
                if (cl == 0) {
                    const uint32_t s = (uint32_t)__cvta_generic_to_shared(&src[warp][grp][0]);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], 64;"
                                 :: "l"(buf + off), "r"(s) : "memory");
                }
            }
        }
        if (MODE == BULK_F32 || MODE == BULK_BF16) {
            if (cl == 0) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
        }
    }
    if (MODE == BULK_F32 || MODE == BULK_BF16) {
        if (cl == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    float *sink; long long *cyc;
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * 4096));
    const int iters = 2000;
    // ---- 1: one CTA of 8 warps per SM x 4 CTAs: enough warps to saturate the pipes
    {
        const int grid = sms * 4;
        auto report = [&](const char *name, int n_lds, int n_shfl, float ms) {
            const double steps = (double)iters * 8;                       // inner steps per warp
            const double warps_per_sm = 4.0 * kThreads / 32;
            const double cycles = ms * 1e-3 * khz * 1e3;
            printf("{\"case\": \"%s\", \"lds128_per_step\": %d, \"shfl_per_step\": %d, \"ms\": %.4f, \"sm_cycles_per_step_per_warp\": %.3f, "
                   "\"cycles_per_lds128\": %.3f, \"cycles_per_shfl\": %.3f}\n", name, n_lds, n_shfl, ms, cycles / (steps * warps_per_sm),
                   n_lds ? cycles / (steps * warps_per_sm * n_lds) : 0.0, n_shfl ? cycles / (steps * warps_per_sm * n_shfl) : 0.0);
        };
        report("lds128_only", 1, 0, time_ms([&] { pipe_mix<1, 0><<<grid, kThreads>>>(iters, sink, cyc); }));
        report("shfl_only_x4", 0, 4, time_ms([&] { pipe_mix<0, 4><<<grid, kThreads>>>(iters, sink, cyc); }));
        report("shfl_only_x8", 0, 8, time_ms([&] { pipe_mix<0, 8><<<grid, kThreads>>>(iters, sink, cyc); }));
        report("lds128_plus_shfl_x4", 1, 4, time_ms([&] { pipe_mix<1, 4><<<grid, kThreads>>>(iters, sink, cyc); }));
        report("lds128_plus_shfl_x8", 1, 8, time_ms([&] { pipe_mix<1, 8><<<grid, kThreads>>>(iters, sink, cyc); }));
        report("lds128_x2_plus_shfl_x4", 2, 4, time_ms([&] { pipe_mix<2, 4><<<grid, kThreads>>>(iters, sink, cyc); }));
    }
    // ---- 2 / 3: chip-wide red throughput into a 32k-row window (4 MB fp32 / 2 MB bf16: L2 resident, spread over all slices)
    {
        char *buf;
        const int rows = 32768;
        CK(cudaMalloc(&buf, (size_t)rows * 128));
        CK(cudaMemset(buf, 0, (size_t)rows * 128));
        const int grid = sms * 4, it2 = 400;
        auto report = [&](const char *name, int row_bytes, float ms) {
            const double nrows = (double)grid * (kThreads / 8) * it2 * 8;
            printf("{\"case\": \"%s\", \"rows\": %.0f, \"row_bytes\": %d, \"ms\": %.4f, \"rows_per_us\": %.1f, \"payload_TBps\": %.3f, "
                   "\"sectors_per_us\": %.1f}\n", name, nrows, row_bytes, ms, nrows / (ms * 1e3), nrows * row_bytes / (ms * 1e-3) / 1e12,
                   nrows * (row_bytes / 32) / (ms * 1e3));
        };
        report("red_v4_f32_row128", 128, time_ms([&] { red_rows<RED_V4_F32><<<grid, kThreads>>>(buf, rows, it2, cyc); }));
        report("bulk_reduce_f32_row128", 128, time_ms([&] { red_rows<BULK_F32><<<grid, kThreads>>>(buf, rows, it2, cyc); }));
        report("red_v2_bf16x2_row64", 64, time_ms([&] { red_rows<RED_V4_BF16X2><<<grid, kThreads>>>(buf, rows, it2, cyc); }));
        report("bulk_reduce_bf16_row64", 64, time_ms([&] { red_rows<BULK_BF16><<<grid, kThreads>>>(buf, rows, it2, cyc); }));
        // correctness of the bulk path: every element of the window is a multiple of its lane pattern
        float h[4];
        CK(cudaMemset(buf, 0, (size_t)rows * 128));
        red_rows<BULK_F32><<<1, kThreads>>>(buf, 1, 1, cyc);      // 32 groups x 8 rows into row 0: expect 256 x {1,2,3,4}
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
        printf("{\"case\": \"bulk_reduce_f32_check\", \"got\": [%.1f, %.1f, %.1f, %.1f], \"want\": [256, 512, 768, 1024]}\n", h[0], h[1], h[2], h[3]);
    }
    return 0;
}
