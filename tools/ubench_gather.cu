// ubench_gather.cu -- what does one gathered / scattered 128-byte row cost on a B200 SM?
//
// The MSDeformAttn kernels gather 64 rows (and, backward, scatter 64 rows) of 128 B per (query, head).
// This microbenchmark measures the per-row cost of the candidate access shapes so that the kernel
// design (and the "SM-local gather" roofline quoted in DESIGN.md) rests on numbers from this machine:
//   ldg128   8 lanes x 16 B per row, 4 independent rows per warp instruction      (LDG.E.128)
//   ldg256   4 lanes x 32 B per row, 8 rows per warp instruction                  (LDG.E.ENL2.256)
//   ldg32    32 lanes x 4 B, 1 row per warp instruction                           (LDG.E)
//   lds128   as ldg128 but from shared memory                                     (LDS.128)
//   red128   8 lanes x red.global.add.v4.f32 per row, 4 rows per instruction      (REDG.E.ADD.F32x4)
//   red32    32 lanes x red.global.add.f32, 1 row per instruction
// each with rows drawn at random from a window of `rows` rows per CTA (small window: L1 / one L2 line set;
// large shared window: L2).  Output: one JSON line per case with ns, cycles per row per SM, TB/s.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kUnroll = 8;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
}

enum Mode { LDG128 = 0, LDG256 = 1, LDG32 = 2, LDS128 = 3, RED128 = 4, RED32 = 5, LDG128_NC = 6, LDG64 = 7, ST128 = 8, RED64 = 9, LDG128_HALF = 10, LDG64_HALF = 11 };

// window_rows: power of two.  shared_window: 0 -> each CTA has its own window (L1-resident when small),
// 1 -> all CTAs draw from the same window of window_rows rows (L2).
template <int MODE>
__global__ void __launch_bounds__(kThreads) bench(float *buf, int window_rows, int shared_window, int iters,
                                                  float *sink, long long *cycles) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    char *base = reinterpret_cast<char *>(buf) + (shared_window ? 0 : (size_t)blockIdx.x * window_rows * 128);
    const uint32_t mask = window_rows - 1;
    int grp, sub_bytes;
    if (MODE == LDG256) { grp = lane >> 2; sub_bytes = (lane & 3) * 32; }
    else if (MODE == LDG128_HALF) { grp = lane >> 2; sub_bytes = (lane & 3) * 16; }      // 8 rows, the first 64 B of each (a bf16 row)
    else if (MODE == LDG64_HALF) { grp = lane >> 3; sub_bytes = (lane & 7) * 8; }        // 4 rows, the first 64 B of each
    else if (MODE == LDG32 || MODE == RED32) { grp = 0; sub_bytes = lane * 4; }
    else if (MODE == LDG64 || MODE == RED64) { grp = lane >> 4; sub_bytes = (lane & 15) * 8; }
    else { grp = lane >> 3; sub_bytes = (lane & 7) * 16; }
    uint32_t seed = (warp_global * 8 + grp) * 2654435761u + 12345u;
    if (MODE == LDS128) {
        for (int i = threadIdx.x; i < window_rows * 32; i += kThreads) smem[i] = (float)i;
        __syncthreads();
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t off[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) off[u] = (lcg(seed) & mask) * 128u + sub_bytes;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (MODE == LDG128 || MODE == LDG128_HALF) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(base + off[u]));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else if (MODE == LDG128_NC) {
                float4 v;
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(base + off[u]));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else if (MODE == LDG256) {
                float a, b, c, d, e, f, g, h;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(a), "=f"(b), "=f"(c), "=f"(d), "=f"(e), "=f"(f), "=f"(g), "=f"(h) : "l"(base + off[u]));
                acc.x += a + e; acc.y += b + f; acc.z += c + g; acc.w += d + h;
            } else if (MODE == LDG64 || MODE == LDG64_HALF) {
                const float2 v = __ldg(reinterpret_cast<const float2 *>(base + off[u]));
                acc.x += v.x; acc.y += v.y;
            } else if (MODE == ST128) {
                *reinterpret_cast<float4 *>(base + off[u]) = make_float4(1.f, 2.f, 3.f, (float)it);
            } else if (MODE == RED64) {
                asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1,%2};" :: "l"(base + off[u]), "f"(1.f), "f"(2.f) : "memory");
            } else if (MODE == LDG32) {
                acc.x += __ldg(reinterpret_cast<const float *>(base + off[u]));
            } else if (MODE == LDS128) {
                const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(smem) + off[u]);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            } else if (MODE == RED128) {
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                             :: "l"(base + off[u]), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
            } else if (MODE == RED32) {
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(base + off[u]), "f"(1.f) : "memory");
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

// LDG.128, 4 rows per warp instruction, of which `n_far` rows come from the big shared (L2) window and the rest from
// the CTA's private L1-resident window: what does ONE missing row cost a request?
__global__ void __launch_bounds__(kThreads) bench_mixed(float *buf, int near_rows, int far_rows, int n_far, int iters,
                                                        float *sink, long long *cycles) {
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int grp = lane >> 3, sub_bytes = (lane & 7) * 16;
    // private windows live after the far window
    char *far_base = reinterpret_cast<char *>(buf);
    char *near_base = far_base + (size_t)far_rows * 128 + (size_t)blockIdx.x * near_rows * 128;
    const bool far = grp < n_far;
    char *base = far ? far_base : near_base;
    const uint32_t mask = (far ? far_rows : near_rows) - 1;
    uint32_t seed = (warp_global * 8 + grp) * 2654435761u + 12345u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // warm the private window
    for (int i = threadIdx.x; i < near_rows * 8; i += kThreads) acc.x += __ldg(reinterpret_cast<const float4 *>(near_base) + i).x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t off[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) off[u] = (lcg(seed) & mask) * 128u + sub_bytes;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(base + off[u]));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

void run_mixed(float *buf, int n_far, int ctas_per_sm, int iters, int sms, float *sink, long long *d_cycles) {
    const int grid = sms * ctas_per_sm;
    const int near_rows = 128, far_rows = 1 << 17;
    float best = 1e30f;
    long long cyc_max = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        bench_mixed<<<grid, kThreads>>>(buf, near_rows, far_rows, n_far, iters, sink, d_cycles);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
            best = ms;
            long long *h = (long long *)malloc(sizeof(long long) * grid);
            CK(cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            cyc_max = 0;
            for (int i = 0; i < grid; ++i) cyc_max = h[i] > cyc_max ? h[i] : cyc_max;
            free(h);
        }
    }
    const double instr_per_sm = (double)ctas_per_sm * (kThreads / 32) * iters * kUnroll;
    printf("{\"case\": \"ldg128_mixed\", \"rows_from_L2_of_4\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"cycles_cta_max\": %lld, "
           "\"cycles_per_row_per_sm\": %.3f, \"cycles_per_instr_per_sm\": %.3f}\n",
           n_far, ctas_per_sm * kThreads / 32, best, cyc_max, cyc_max / (instr_per_sm * 4), cyc_max / instr_per_sm);
    fflush(stdout);
}

template <int MODE>
void run(const char *name, float *buf, size_t buf_bytes, int window_rows, int shared_window, int ctas_per_sm, int iters,
         int sms, float *sink, long long *d_cycles, const char *note) {
    const int grid = sms * ctas_per_sm;
    if (!shared_window && (size_t)grid * window_rows * 128 > buf_bytes) { printf("buffer too small for %s\n", name); return; }
    const size_t smem = MODE == LDS128 ? (size_t)window_rows * 128 : 0;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    long long cyc_max = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        bench<MODE><<<grid, kThreads, smem>>>(buf, window_rows, shared_window, iters, sink, d_cycles);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
            best = ms;
            long long *h = (long long *)malloc(sizeof(long long) * grid);
            CK(cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            cyc_max = 0;
            for (int i = 0; i < grid; ++i) cyc_max = h[i] > cyc_max ? h[i] : cyc_max;
            free(h);
        }
    }
    const int rows_per_instr = (MODE == LDG256 || MODE == LDG128_HALF) ? 8 : (MODE == LDG32 || MODE == RED32) ? 1 : (MODE == LDG64 || MODE == RED64) ? 2 : 4;
    const double instr_per_sm = (double)ctas_per_sm * (kThreads / 32) * iters * kUnroll;
    const double rows_per_sm = instr_per_sm * rows_per_instr;
    const double bytes = rows_per_sm * sms * 128.0;
    printf("{\"case\": \"%s\", \"window_rows\": %d, \"shared_window\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, "
           "\"cycles_cta_max\": %lld, \"cycles_per_row_per_sm\": %.3f, \"cycles_per_instr_per_sm\": %.3f, \"TBps\": %.3f, \"note\": \"%s\"}\n",
           name, window_rows, shared_window, ctas_per_sm * kThreads / 32, best, cyc_max, cyc_max / rows_per_sm,
           cyc_max / instr_per_sm, bytes / (best * 1e-3) / 1e12, note);
    fflush(stdout);
}

int main() {
    int sms = 148;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t buf_bytes = (size_t)1 << 30;
    float *buf, *sink;
    long long *d_cycles;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 0, buf_bytes));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMalloc(&d_cycles, sizeof(long long) * sms * 32));
    const int iters = 400;
    if (getenv("UBENCH_HALF")) {      // is the L1 cost of a request per 128-byte LINE touched or per BYTE delivered?  (bf16 rows are half lines)
        for (int cps : {4, 8}) {
            run<LDG128>("ldg128_4rows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "4 full rows per instruction (fp32 layout)");
            run<LDG64_HALF>("ldg64_4halfrows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "4 half rows (64 B) per instruction: the bf16 kernels' shape; TBps column counts 128 B per row");
            run<LDG128_HALF>("ldg128_8halfrows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "8 half rows (64 B) per instruction; TBps column counts 128 B per row");
            run<LDG128_HALF>("ldg128_8halfrows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "8 half rows per instruction, shared 16 MB window");
        }
        return 0;
    }
    if (getenv("UBENCH_FULL")) {
    for (int cps : {2, 4, 8}) {
        run<LDG128>("ldg128_4rows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "private 16 KB window per CTA");
        run<LDG256>("ldg256_8rows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "private 16 KB window per CTA");
        run<LDG32>("ldg32_1row_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "private 16 KB window per CTA");
        run<LDS128>("lds128_4rows", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "16 KB of shared memory per CTA");
        run<LDG128>("ldg128_4rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2)");
        run<LDG128_NC>("ldg128nc_4rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2), L1::no_allocate");
        run<LDG256>("ldg256_8rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2)");
        run<LDG32>("ldg32_1row_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2)");
        run<RED128>("red128_4rows_win", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "private 16 KB window per CTA");
        run<RED128>("red128_4rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2)");
        run<RED32>("red32_1row_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "shared 16 MB window (L2)");
        run<RED128>("red128_4rows_HBM", buf, buf_bytes, 1 << 22, 1, cps, iters, sms, sink, d_cycles, "shared 512 MB window (> L2)");
    }
    }
    // second series: what a partially missing request costs; narrower requests; stores; reds on part of the chip
    for (int cps : {4, 8}) {
        if (getenv("UBENCH_HOT")) break;
        for (int n_far = 0; n_far <= 4; ++n_far) run_mixed(buf, n_far, cps, iters, sms, sink, d_cycles);
        run<LDG64>("ldg64_2rows_L1", buf, buf_bytes, 128, 0, cps, iters, sms, sink, d_cycles, "16 lanes x 8 B per row, private window");
        run<LDG64>("ldg64_2rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "16 lanes x 8 B per row, shared 16 MB window");
        run<ST128>("st128_4rows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "plain stores, shared 16 MB window");
        run<RED64>("red64_2halfrows_L2", buf, buf_bytes, 1 << 17, 1, cps, iters, sms, sink, d_cycles, "red.v2.f32: two 128 B rows per instruction");
    }
    // hot spots: every CTA reds into the SAME few rows (a coarse pyramid level has 60 .. 720 rows per head)
    for (int rows : {64, 512, 4096, 32768})
        run<RED128>("red128_4rows_hot", buf, buf_bytes, rows, 1, 4, iters, sms, sink, d_cycles, "all CTAs share this many rows");
    for (int part : {37, 74, 111}) {
        char note[64];
        snprintf(note, sizeof(note), "only %d of %d SMs issue reds", part, sms);
        run<RED128>("red128_4rows_L2_partial", buf, buf_bytes, 1 << 17, 1, 4, iters, part, sink, d_cycles, note);
        run<LDG128>("ldg128_4rows_L2_partial", buf, buf_bytes, 1 << 17, 1, 4, iters, part, sink, d_cycles, note);
    }
    return 0;
}
