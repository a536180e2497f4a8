// msda_sm100.cu -- multi-scale deformable attention forward / backward for B200 (sm_100a).
//
// Written from scratch for Blackwell; the operator it implements is the reference's
// models/ops/src/cuda/ms_deform_im2col_cuda.cuh (fwd :237-299, bwd :301-403, bilinear helpers :33-159)
// behind the C ABI of include/msda_sm100.h.  See DESIGN.md for the data layout and the roofline
// of each kernel.
//
// Why it looks the way it does.  Per (query, head) the op gathers L*P*4 = 64 rows of 32 channels
// (128 B each in fp32): 8 KB of gather for 448 B of compulsory traffic.  The gather is served by the
// SM's L1/shared-memory data path (one 128 B wavefront per row), not by HBM, so the kernels are
// organised around three things:
//   1. locality: a CTA works on an 8x8 *spatial tile* of queries of ONE head at a time, so the rows its
//      16 warps gather overlap and stay resident in L1 (the whole of `value` stays resident in the
//      126 MB L2);
//   2. few, wide memory instructions: 8 lanes x 128-bit cover one 32-channel row, the four 8-lane
//      groups of a warp work on four x-adjacent queries, so one LDG.128 / RED.128 moves four rows;
//   3. few issue slots: the per-point geometry (pixel coordinates, bilinear weights, corner offsets,
//      validity) is computed once per point by one lane, staged in shared memory, and read back by the
//      lanes that gather with one or two broadcast LDS.128 per point.
// The backward replaces the reference's per-channel scalar atomics and its serial 32-term reductions by
// red.global.add.v4.f32 (one instruction per four rows) and an 8-lane transposing shuffle reduction.
//
// No tensor cores: the op is a gather / scatter with ~0.2 flop per byte.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "msda_sm100.h"

namespace {

// ------------------------------------------------------------------------------------------------
// host-side state: last error (per thread), launch counter, tuning knobs, cached SM counts
// ------------------------------------------------------------------------------------------------
thread_local char g_err[256] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_fwd_ctas_per_sm{0};   // 0 = kernel default
std::atomic<int> g_bwd_ctas_per_sm{0};
std::atomic<int> g_force_generic{0};
std::atomic<int> g_force_linear{0};     // experiments: never use the tiled query walk
std::atomic<int> g_skip_scatter{0};     // experiments: see Dims::debug_skip_scatter

int fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int fail_cuda(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
constexpr int kMaxLevels = 16;   // levels the tiled kernels keep in shared memory
constexpr int kWarps = 16;       // warps per CTA of the tiled kernels (512 threads)
constexpr int kTile = 8;         // a CTA pass covers an 8 x 8 tile of queries: 16 warps x 4 lane groups
constexpr int kPad = 9;          // 8 staged points per lane group + 1 slot of padding (bank spread)

struct Dims {
    int N, S, M, L, Lq, P;       // D is a template parameter / 32 for the tiled kernels
    int tiled;                   // 1: Lq == S, walk queries as spatial tiles of their own level
    int debug_skip_scatter;      // experiments only: backward omits the grad_value reds (wrong grad_value)
};

struct LevelTable {              // shared memory, filled once per CTA from the int64 device tensors
    int H[kMaxLevels], W[kMaxLevels];
    int start[kMaxLevels];       // first row of the level
    int tiles_x[kMaxLevels];
    int tile_cum[kMaxLevels + 1];
    int dense;                   // 1 if the levels tile [0, Lq) exactly: start[l] == sum_{k<l} H_k*W_k, total == Lq
};

__device__ __forceinline__ float2 ld_stream_f2(const float *p) {
    float2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
    float r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// grad_value[p .. p+3] += s * {lo.x, lo.y, hi.x, hi.y}: one REDG.E.ADD.F32x4 (sm_90+), no return value.
__device__ __forceinline__ void red_add_row(const char *p, float s, float2 lo, float2 hi) {
    const float2 a = __fmul2_rn(make_float2(s, s), lo), b = __fmul2_rn(make_float2(s, s), hi);
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}
__device__ __forceinline__ const char *at(const char *base, int off) { return base + (uint32_t)off; }

// One sample point as the lanes that gather need it.
//   o[4]    : byte offsets (inside the frame+head slice of `value`) of the corners (y0,x0), (y0,x0+1),
//             (y0+1,x0), (y0+1,x0+1).  Rows / columns that fall outside the level are CLAMPED
//             onto the nearest valid one and their bilinear weight is zeroed instead: every load is in
//             bounds and unpredicated, and zero padding comes out of the weights (cuh:56-78).
//   w[4]    : bilinear weights hy*hx, hy*lx, ly*hx, ly*lx, zero for invalid corners / out-of-range points.
struct PointGeo {
    int o00, o01, o10, o11;
    float w00, w01, w10, w11;
    float lx, ly;
    int valid;      // bit i: corner i contributes (00, 01, 10, 11)
};

// Pixel coordinates exactly as the reference's compiled kernel forms them: fma(loc, size, -0.5)
// (ms_deform_im2col_cuda.cuh:285-286; nvcc contracts the expression -- SASS of the reference op built for
// sm_100a: `FFMA R29, R12, R29, -0.5` -- so borderline points fall into the same bilinear cell as there),
// range test of :288, corner tests of :56/:62/:68/:74.
__device__ __forceinline__ PointGeo point_geometry(float loc_x, float loc_y, int H, int W, int level_base_bytes,
                                                   int pixel_bytes) {
    PointGeo g;
    const float fw = (float)W, fh = (float)H;
    const float x = fmaf(loc_x, fw, -0.5f);
    const float y = fmaf(loc_y, fh, -0.5f);
    const bool in_range = (y > -1.f) && (x > -1.f) && (y < fh) && (x < fw);   // false for NaN / inf
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;     // saturating conversion; clamped below when out of range
    g.lx = in_range ? x - xf : 0.f;
    g.ly = in_range ? y - yf : 0.f;
    const float hx = 1.f - g.lx, hy = 1.f - g.ly;
    const bool xa = in_range && x0 >= 0, xb = in_range && x0 + 1 <= W - 1;
    const bool ya = in_range && y0 >= 0, yb = in_range && y0 + 1 <= H - 1;
    const int xs = min(max(x0, -1), W - 1), ys = min(max(y0, -1), H - 1);   // [-1, size-1]: no overflow below
    const int xc0 = max(xs, 0), xc1 = min(xs + 1, W - 1), yc0 = max(ys, 0), yc1 = min(ys + 1, H - 1);
    const int dx = (xc1 - xc0) * pixel_bytes;
    g.o00 = level_base_bytes + (yc0 * W + xc0) * pixel_bytes;
    g.o10 = level_base_bytes + (yc1 * W + xc0) * pixel_bytes;
    g.o01 = g.o00 + dx;
    g.o11 = g.o10 + dx;
    g.w00 = (xa && ya) ? hy * hx : 0.f;
    g.w01 = (xb && ya) ? hy * g.lx : 0.f;
    g.w10 = (xa && yb) ? g.ly * hx : 0.f;
    g.w11 = (xb && yb) ? g.ly * g.lx : 0.f;
    g.valid = (int)(xa && ya) | ((int)(xb && ya) << 1) | ((int)(xa && yb) << 2) | ((int)(xb && yb) << 3);
    return g;
}

// Which four queries (one per 8-lane group) a warp handles in pass `tile`, and whether they exist.
struct QuerySel {
    int q;
    bool valid;
};
__device__ __forceinline__ QuerySel select_query(bool tiled, const Dims &d, const LevelTable &lt, int tile, int warp,
                                                 int grp) {
    QuerySel s;
    if (tiled) {
        int lv = 0;
        while (lv + 1 < d.L && tile >= lt.tile_cum[lv + 1]) ++lv;
        const int t = tile - lt.tile_cum[lv];
        const int ty = t / lt.tiles_x[lv], tx = t - ty * lt.tiles_x[lv];
        const int y = ty * kTile + (warp >> 1), x = tx * kTile + ((warp & 1) << 2) + grp;
        s.valid = (y < lt.H[lv]) && (x < lt.W[lv]);
        s.q = lt.start[lv] + y * lt.W[lv] + x;
    } else {
        s.q = tile * (kWarps * 4) + warp * 4 + grp;
        s.valid = s.q < d.Lq;
    }
    return s;
}

__device__ __forceinline__ void load_level_table(LevelTable &lt, const int64_t *shapes, const int64_t *start, int L,
                                                 int Lq) {
    if (threadIdx.x < L) {
        lt.H[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        lt.W[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        lt.start[threadIdx.x] = (int)start[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cum = 0, rows = 0, dense = 1;
        for (int l = 0; l < L; ++l) {
            const int tx = (lt.W[l] + kTile - 1) / kTile, ty = (lt.H[l] + kTile - 1) / kTile;
            lt.tiles_x[l] = tx;
            lt.tile_cum[l] = cum;
            cum += tx * ty;
            dense &= (lt.start[l] == rows) && lt.H[l] > 0 && lt.W[l] > 0;
            rows += lt.H[l] * lt.W[l];
        }
        lt.tile_cum[L] = cum;
        lt.dense = dense && rows == Lq;
    }
    __syncthreads();
}

// Packed fp32 pairs: Blackwell's FFMA2 / FMUL2 do two lanes of fp32 math per issue slot, with a scalar
// operand broadcast for free -- the kernels are issue-bound, so every row update is written in pairs.
__device__ __forceinline__ float2 fma2s(float s, float2 v, float2 acc) { return __ffma2_rn(make_float2(s, s), v, acc); }
__device__ __forceinline__ float2 mul2s(float s, float2 v) { return __fmul2_rn(make_float2(s, s), v); }
struct Row {            // four channels as two pairs
    float2 lo, hi;
};
template <typename VT> struct RowIO;
template <> struct RowIO<float> {
    static constexpr int kBytes = 16;          // bytes of four channels
    static __device__ __forceinline__ Row load(const char *p) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
        return Row{make_float2(v.x, v.y), make_float2(v.z, v.w)};
    }
    static __device__ __forceinline__ void store(char *p, Row r) {
        *reinterpret_cast<float4 *>(p) = make_float4(r.lo.x, r.lo.y, r.hi.x, r.hi.y);
    }
};
template <> struct RowIO<__nv_bfloat16> {
    static constexpr int kBytes = 8;
    static __device__ __forceinline__ Row load(const char *p) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
        return Row{make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u)),
                   make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u))};
    }
    static __device__ __forceinline__ void store(char *p, Row r) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r.lo.x, r.lo.y), hi = __floats2bfloat162_rn(r.hi.x, r.hi.y);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t *>(&lo);
        raw.y = *reinterpret_cast<uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(p) = raw;
    }
};

// ------------------------------------------------------------------------------------------------
// Tiled forward, D = 32.  grid: persistent, CTA b takes tasks b, b+grid, ... ; a task is
// (frame n, query tile, head m), m fastest.  ROUNDS = ceil(L*P / 8); a round stages 8 points per lane
// group (slots past L*P carry zero weights), then gathers them branch-free.
// ------------------------------------------------------------------------------------------------
template <typename VT, int ROUNDS>
__global__ void __launch_bounds__(kWarps * 32, 2)
msda_fwd_tiled(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
               const float *__restrict__ loc, const float *__restrict__ attn, VT *__restrict__ out, Dims d) {
    __shared__ LevelTable lt;
    __shared__ float4 s_w[kWarps][4][kPad];   // a*w00, a*w01, a*w10, a*w11
    __shared__ int4 s_o[kWarps][4][kPad];     // byte offsets of the four corners
    load_level_table(lt, shapes, start, d.L, d.Lq);
    const bool tiled = d.tiled && lt.dense;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int pts = d.L * d.P;
    const int pixel_bytes = d.M * 32 * (int)sizeof(VT);
    int lvl[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) lvl[r] = min((8 * r + cl) / d.P, d.L - 1);

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kWarps * 4 - 1) / (kWarps * 4);
    const int64_t total = (int64_t)d.N * tiles * d.M;
    for (int64_t task = blockIdx.x; task < total; task += gridDim.x) {
        const int m = (int)(task % d.M);
        const int tile = (int)((task / d.M) % tiles);
        const int64_t n = task / ((int64_t)d.M * tiles);
        const QuerySel qs = select_query(tiled, d, lt, tile, warp, grp);
        if (!__any_sync(0xffffffffu, qs.valid)) continue;
        const int64_t row = (n * d.Lq + qs.q) * d.M + m;              // (n, q, m)
        // frame n, head m, this lane's four channels
        const char *vb = reinterpret_cast<const char *>(value) +
                         ((n * d.S * d.M + m) * 32 + cl * 4) * (int64_t)sizeof(VT);
        float2 acc_lo = make_float2(0.f, 0.f), acc_hi = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pt = 8 * r + cl;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            int4 o = make_int4(0, 0, 0, 0);
            if (qs.valid && pt < pts) {
                const float2 xy = ld_stream_f2(loc + (row * pts + pt) * 2);
                const float a = ld_stream_f1(attn + row * pts + pt);
                const int l = lvl[r];
                const PointGeo g = point_geometry(xy.x, xy.y, lt.H[l], lt.W[l], lt.start[l] * pixel_bytes, pixel_bytes);
                w = make_float4(a * g.w00, a * g.w01, a * g.w10, a * g.w11);
                o = make_int4(g.o00, g.o01, g.o10, g.o11);
            }
            s_w[warp][grp][cl] = w;
            s_o[warp][grp][cl] = o;
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const float4 pw = s_w[warp][grp][it];
                const int4 po = s_o[warp][grp][it];
                const Row v00 = RowIO<VT>::load(at(vb, po.x));
                const Row v01 = RowIO<VT>::load(at(vb, po.y));
                const Row v10 = RowIO<VT>::load(at(vb, po.z));
                const Row v11 = RowIO<VT>::load(at(vb, po.w));
                acc_lo = fma2s(pw.x, v00.lo, acc_lo); acc_hi = fma2s(pw.x, v00.hi, acc_hi);
                acc_lo = fma2s(pw.y, v01.lo, acc_lo); acc_hi = fma2s(pw.y, v01.hi, acc_hi);
                acc_lo = fma2s(pw.z, v10.lo, acc_lo); acc_hi = fma2s(pw.z, v10.hi, acc_hi);
                acc_lo = fma2s(pw.w, v11.lo, acc_lo); acc_hi = fma2s(pw.w, v11.hi, acc_hi);
            }
            __syncwarp();
        }
        if (qs.valid)
            RowIO<VT>::store(reinterpret_cast<char *>(out) + (row * 32 + cl * 4) * (int64_t)sizeof(VT), Row{acc_lo, acc_hi});
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled backward, D = 32.  Same task walk as the forward.  grad_value (fp32) must be zero on entry.
//
// Per point and per lane (4 channels) the loop only forms the four corner dot products
//     p_i = sum_c grad_out[c] * v_i[c]
// and the scatter rows (a * w_i) * grad_out.  The 8-lane sums of p_i land, transposed, in the lane that
// owns the point, which finishes with scalars (cuh:123-158 regrouped by corner):
//     grad_attn = sum_i w_i p_i
//     grad_x    = W * a * ( hy (p01 - p00) + ly (p11 - p10) )      (invalid corners dropped)
//     grad_y    = H * a * ( hx (p10 - p00) + lx (p11 - p01) )
// ------------------------------------------------------------------------------------------------
// Sum v[0..7] across the 8 lanes of a group so that lane `cl` ends with the total of v[cl]:
// 7 shuffles instead of 24 for eight separate butterfly reductions.
__device__ __forceinline__ float transpose_reduce8(const float (&v)[8], int cl) {
    float u[4], t[2];
    const bool b2 = cl & 4, b1 = cl & 2, b0 = cl & 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b2 ? v[j + 4] : v[j], send = b2 ? v[j] : v[j + 4];
        u[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b1 ? u[j + 2] : u[j], send = b1 ? u[j] : u[j + 2];
        t[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? t[1] : t[0], send = b0 ? t[0] : t[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

__device__ __forceinline__ float dot4(const Row &a, const Row &b) {
    const float2 t = __ffma2_rn(a.hi, b.hi, __fmul2_rn(a.lo, b.lo));
    return t.x + t.y;
}

template <typename VT, int ROUNDS>
__global__ void __launch_bounds__(kWarps * 32, 1)
msda_bwd_tiled(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ start, const float *__restrict__ loc, const float *__restrict__ attn,
               float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn, Dims d) {
    __shared__ LevelTable lt;
    __shared__ float4 s_w[kWarps][4][kPad];   // a*w00, a*w01, a*w10, a*w11   (scatter weights)
    __shared__ int4 s_o[kWarps][4][kPad];     // byte offsets of the four corners
    load_level_table(lt, shapes, start, d.L, d.Lq);
    const bool tiled = d.tiled && lt.dense;
    constexpr int kGradScale = 4 / (int)sizeof(VT);   // grad_value is fp32: its byte offsets are this x value's

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int pts = d.L * d.P;
    const int pixel_bytes = d.M * 32 * (int)sizeof(VT);
    int lvl[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) lvl[r] = min((8 * r + cl) / d.P, d.L - 1);

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kWarps * 4 - 1) / (kWarps * 4);
    const int64_t total = (int64_t)d.N * tiles * d.M;
    for (int64_t task = blockIdx.x; task < total; task += gridDim.x) {
        const int m = (int)(task % d.M);
        const int tile = (int)((task / d.M) % tiles);
        const int64_t n = task / ((int64_t)d.M * tiles);
        const QuerySel qs = select_query(tiled, d, lt, tile, warp, grp);
        if (!__any_sync(0xffffffffu, qs.valid)) continue;
        const int64_t row = (n * d.Lq + qs.q) * d.M + m;
        const int64_t slice = (n * d.S * d.M + m) * 32 + cl * 4;      // element offset of frame n, head m, 4 channels
        const char *vb = reinterpret_cast<const char *>(value) + slice * (int64_t)sizeof(VT);
        const char *gb = reinterpret_cast<const char *>(grad_value) + slice * 4;
        Row go{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (qs.valid) go = RowIO<VT>::load(reinterpret_cast<const char *>(grad_out) + (row * 32 + cl * 4) * (int64_t)sizeof(VT));
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pt = 8 * r + cl;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            int4 o = make_int4(0, 0, 0, 0);
            PointGeo own;                                   // this lane's own point, kept for the epilogue
            own.w00 = own.w01 = own.w10 = own.w11 = own.lx = own.ly = 0.f;
            own.valid = 0;
            float a_own = 0.f, fw_own = 0.f, fh_own = 0.f;
            if (qs.valid && pt < pts) {
                const float2 xy = ld_stream_f2(loc + (row * pts + pt) * 2);
                a_own = ld_stream_f1(attn + row * pts + pt);
                const int l = lvl[r];
                own = point_geometry(xy.x, xy.y, lt.H[l], lt.W[l], lt.start[l] * pixel_bytes, pixel_bytes);
                fw_own = (float)lt.W[l];
                fh_own = (float)lt.H[l];
                w = make_float4(a_own * own.w00, a_own * own.w01, a_own * own.w10, a_own * own.w11);
                o = make_int4(own.o00, own.o01, own.o10, own.o11);
            }
            s_w[warp][grp][cl] = w;
            s_o[warp][grp][cl] = o;
            __syncwarp();
            float p00[8], p01[8], p10[8], p11[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const float4 pw = s_w[warp][grp][it];
                const int4 po = s_o[warp][grp][it];
                const Row v00 = RowIO<VT>::load(at(vb, po.x));
                const Row v01 = RowIO<VT>::load(at(vb, po.y));
                const Row v10 = RowIO<VT>::load(at(vb, po.z));
                const Row v11 = RowIO<VT>::load(at(vb, po.w));
                p00[it] = dot4(go, v00);
                p01[it] = dot4(go, v01);
                p10[it] = dot4(go, v10);
                p11[it] = dot4(go, v11);
                // scatter: grad_value[corner] += (a * w_corner) * grad_out              (cuh:125,134,143,152)
                // A point with no contributing corner (out of range, zero attention) is skipped as a whole;
                // otherwise its clamped corners receive +0, which is harmless.
                const uint32_t any_w = (__float_as_uint(pw.x) | __float_as_uint(pw.y) | __float_as_uint(pw.z) |
                                        __float_as_uint(pw.w)) << 1;
                if (any_w != 0u && !d.debug_skip_scatter) {
                    red_add_row(at(gb, po.x * kGradScale), pw.x, go.lo, go.hi);
                    red_add_row(at(gb, po.y * kGradScale), pw.y, go.lo, go.hi);
                    red_add_row(at(gb, po.z * kGradScale), pw.z, go.lo, go.hi);
                    red_add_row(at(gb, po.w * kGradScale), pw.w, go.lo, go.hi);
                }
            }
            __syncwarp();
            float q00 = transpose_reduce8(p00, cl), q01 = transpose_reduce8(p01, cl);
            float q10 = transpose_reduce8(p10, cl), q11 = transpose_reduce8(p11, cl);
            if (qs.valid && pt < pts) {
                // clamped (invalid) corners carry someone else's row: drop them (zero padding, cuh:56-78)
                q00 = (own.valid & 1) ? q00 : 0.f;
                q01 = (own.valid & 2) ? q01 : 0.f;
                q10 = (own.valid & 4) ? q10 : 0.f;
                q11 = (own.valid & 8) ? q11 : 0.f;
                const float hx = 1.f - own.lx, hy = 1.f - own.ly;
                const float ga = own.w00 * q00 + own.w01 * q01 + own.w10 * q10 + own.w11 * q11;       // :156
                const float gx = hy * (q01 - q00) + own.ly * (q11 - q10);                             // :157
                const float gy = hx * (q10 - q00) + own.lx * (q11 - q01);                             // :158
                grad_attn[row * pts + pt] = ga;
                *reinterpret_cast<float2 *>(grad_loc + (row * pts + pt) * 2) =
                    make_float2(fw_own * a_own * gx, fh_own * a_own * gy);
            }
        }
    }
}

// fp32 accumulator -> bf16 (only for the bf16 backward when the caller wants a bf16 grad_value)
__global__ void msda_f32_to_bf16(const float4 *__restrict__ src, uint2 *__restrict__ dst, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t *>(&lo);
        raw.y = *reinterpret_cast<uint32_t *>(&hi);
        dst[i] = raw;
    }
}

// ------------------------------------------------------------------------------------------------
// Generic kernels: any D, L, P; float or double.  Correctness path for the shapes the tiled kernels
// do not cover (the reference's tests use D in {30, 32, 64, 71, 1025, 2048, 3096} in fp64, test.py:85).
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Cell {
    int64_t c[4];     // element offsets of the four corners inside the frame (head 0, channel 0), -1 if invalid
    T lx, ly;
    bool in_range;
};

template <typename T>
__device__ __forceinline__ Cell<T> generic_cell(T loc_x, T loc_y, int H, int W, int64_t level_start, int MD) {
    Cell<T> g;
    const T x = loc_x * (T)W - (T)0.5, y = loc_y * (T)H - (T)0.5;
    g.in_range = (y > (T)-1) && (x > (T)-1) && (y < (T)H) && (x < (T)W);
    const T xf = floor(x), yf = floor(y);
    const int x0 = g.in_range ? (int)xf : 0, y0 = g.in_range ? (int)yf : 0;
    g.lx = g.in_range ? x - xf : (T)0;
    g.ly = g.in_range ? y - yf : (T)0;
    const bool xa = g.in_range && x0 >= 0, xb = g.in_range && x0 + 1 <= W - 1, ya = y0 >= 0, yb = y0 + 1 <= H - 1;
    const int64_t p00 = (level_start + (int64_t)y0 * W + x0) * MD;
    g.c[0] = (xa && ya) ? p00 : -1;
    g.c[1] = (xb && ya) ? p00 + MD : -1;
    g.c[2] = (xa && yb) ? p00 + (int64_t)W * MD : -1;
    g.c[3] = (xb && yb) ? p00 + (int64_t)W * MD + MD : -1;
    return g;
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                 const T *__restrict__ loc, const T *__restrict__ attn, T *__restrict__ out, Dims d, int D) {
    const int64_t total = (int64_t)d.N * d.Lq * d.M * D;
    const int pts = d.L * d.P, MD = d.M * D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const int64_t row = i / D;
        const int m = (int)(row % d.M);
        const int64_t n = row / ((int64_t)d.M * d.Lq);
        const T *frame = value + n * d.S * MD + m * D + c;
        T acc = 0;
        for (int l = 0; l < d.L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const int64_t ls = start[l];
            for (int p = 0; p < d.P; ++p) {
                const int64_t k = row * pts + l * d.P + p;
                const Cell<T> g = generic_cell<T>(loc[2 * k], loc[2 * k + 1], H, W, ls, MD);
                if (!g.in_range) continue;
                const T hx = (T)1 - g.lx, hy = (T)1 - g.ly;
                const T v0 = g.c[0] >= 0 ? frame[g.c[0]] : (T)0, v1 = g.c[1] >= 0 ? frame[g.c[1]] : (T)0;
                const T v2 = g.c[2] >= 0 ? frame[g.c[2]] : (T)0, v3 = g.c[3] >= 0 ? frame[g.c[3]] : (T)0;
                acc += attn[k] * (hy * hx * v0 + hy * g.lx * v1 + g.ly * hx * v2 + g.ly * g.lx * v3);
            }
        }
        out[i] = acc;
    }
}

// One CTA (128 threads) per (n, q, m) row; threads stride over channels, partial sums of the three
// point gradients are combined with a shuffle + shared-memory block reduction.
template <typename T>
__global__ void __launch_bounds__(128)
msda_bwd_generic(const T *__restrict__ grad_out, const T *__restrict__ value, const int64_t *__restrict__ shapes,
                 const int64_t *__restrict__ start, const T *__restrict__ loc, const T *__restrict__ attn,
                 T *__restrict__ grad_value, T *__restrict__ grad_loc, T *__restrict__ grad_attn, Dims d, int D) {
    __shared__ T red[3][4];
    const int64_t rows = (int64_t)d.N * d.Lq * d.M;
    const int pts = d.L * d.P, MD = d.M * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int m = (int)(row % d.M);
        const int64_t n = row / ((int64_t)d.M * d.Lq);
        const int64_t fbase = n * d.S * MD + m * D;
        for (int l = 0; l < d.L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
            const int64_t ls = start[l];
            for (int p = 0; p < d.P; ++p) {
                const int64_t k = row * pts + l * d.P + p;
                const Cell<T> g = generic_cell<T>(loc[2 * k], loc[2 * k + 1], H, W, ls, MD);
                const T a = attn[k];
                T sa = 0, sx = 0, sy = 0;
                if (g.in_range) {
                    const T hx = (T)1 - g.lx, hy = (T)1 - g.ly;
                    const T w0 = hy * hx, w1 = hy * g.lx, w2 = g.ly * hx, w3 = g.ly * g.lx;
                    for (int c = threadIdx.x; c < D; c += blockDim.x) {
                        const T go = grad_out[row * D + c];
                        const T ga = go * a;
                        T v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                        if (g.c[0] >= 0) { v0 = value[fbase + g.c[0] + c]; atomicAdd(grad_value + fbase + g.c[0] + c, w0 * ga); }
                        if (g.c[1] >= 0) { v1 = value[fbase + g.c[1] + c]; atomicAdd(grad_value + fbase + g.c[1] + c, w1 * ga); }
                        if (g.c[2] >= 0) { v2 = value[fbase + g.c[2] + c]; atomicAdd(grad_value + fbase + g.c[2] + c, w2 * ga); }
                        if (g.c[3] >= 0) { v3 = value[fbase + g.c[3] + c]; atomicAdd(grad_value + fbase + g.c[3] + c, w3 * ga); }
                        sa += go * (w0 * v0 + w1 * v1 + w2 * v2 + w3 * v3);
                        sx += ga * (hy * (v1 - v0) + g.ly * (v3 - v2));
                        sy += ga * (hx * (v2 - v0) + g.lx * (v3 - v1));
                    }
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, s);
                    sx += __shfl_xor_sync(0xffffffffu, sx, s);
                    sy += __shfl_xor_sync(0xffffffffu, sy, s);
                }
                if (lane == 0) { red[0][warp] = sa; red[1][warp] = sx; red[2][warp] = sy; }
                __syncthreads();
                if (threadIdx.x == 0) {
                    grad_attn[k] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
                    grad_loc[2 * k] = (T)W * (red[1][0] + red[1][1] + red[1][2] + red[1][3]);
                    grad_loc[2 * k + 1] = (T)H * (red[2][0] + red[2][1] + red[2][2] + red[2][3]);
                }
                __syncthreads();
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
Dims make_dims(int N, int S, int M, int L, int Lq, int P) {
    return Dims{N, S, M, L, Lq, P, (Lq == S && !g_force_linear.load()) ? 1 : 0, g_skip_scatter.load()};
}

bool tiled_ok(int channels, int L, int P) {
    return !g_force_generic.load() && channels == 32 && L <= kMaxLevels && L * P <= 32;
}

int check_dims(int N, int S, int M, int D, int L, int Lq, int P) {
    if (N < 0 || Lq < 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || P <= 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda: dimensions must be positive (batch and num_query may be 0)");
    if ((int64_t)Lq * M * L * P * 2 >= (int64_t)1 << 40 || (int64_t)S * M * D >= (int64_t)1 << 29)
        return fail(MSDA_ERR_UNSUPPORTED, "msda: a single frame exceeds 2^29 value elements (2 GiB of fp32)");
    return MSDA_OK;
}

bool misaligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

int grid_for(int ctas_per_sm_default, const std::atomic<int> &knob) {
    const int k = knob.load();
    return sm_count() * (k > 0 ? k : ctas_per_sm_default);
}

int after_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MSDA_OK : fail_cuda(e, what);
}

template <typename VT>
int launch_fwd_tiled(const VT *value, const int64_t *shapes, const int64_t *start, const float *loc, const float *attn,
                     VT *out, const Dims &d, cudaStream_t st) {
    const int rounds = (d.L * d.P + 7) / 8;
    const int grid = grid_for(2, g_fwd_ctas_per_sm);
    switch (rounds) {
        case 1: msda_fwd_tiled<VT, 1><<<grid, kWarps * 32, 0, st>>>(value, shapes, start, loc, attn, out, d); break;
        case 2: msda_fwd_tiled<VT, 2><<<grid, kWarps * 32, 0, st>>>(value, shapes, start, loc, attn, out, d); break;
        case 3: msda_fwd_tiled<VT, 3><<<grid, kWarps * 32, 0, st>>>(value, shapes, start, loc, attn, out, d); break;
        default: msda_fwd_tiled<VT, 4><<<grid, kWarps * 32, 0, st>>>(value, shapes, start, loc, attn, out, d); break;
    }
    return after_launch("msda_fwd_tiled");
}

template <typename VT>
int launch_bwd_tiled(const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                     const float *attn, float *gv, float *gl, float *ga, const Dims &d, cudaStream_t st) {
    const int rounds = (d.L * d.P + 7) / 8;
    const int grid = grid_for(1, g_bwd_ctas_per_sm);
    switch (rounds) {
        case 1: msda_bwd_tiled<VT, 1><<<grid, kWarps * 32, 0, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d); break;
        case 2: msda_bwd_tiled<VT, 2><<<grid, kWarps * 32, 0, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d); break;
        case 3: msda_bwd_tiled<VT, 3><<<grid, kWarps * 32, 0, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d); break;
        default: msda_bwd_tiled<VT, 4><<<grid, kWarps * 32, 0, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d); break;
    }
    return after_launch("msda_bwd_tiled");
}

template <typename T>
int forward_generic(const T *value, const int64_t *shapes, const int64_t *start, const T *loc, const T *attn, T *out,
                    const Dims &d, int D, cudaStream_t st) {
    const int64_t total = (int64_t)d.N * d.Lq * d.M * D;
    const int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 32 ? (total + 255) / 256 : (int64_t)sm_count() * 32);
    msda_fwd_generic<T><<<grid, 256, 0, st>>>(value, shapes, start, loc, attn, out, d, D);
    return after_launch("msda_fwd_generic");
}

template <typename T>
int backward_generic(const T *go, const T *value, const int64_t *shapes, const int64_t *start, const T *loc, const T *attn,
                     T *gv, T *gl, T *ga, const Dims &d, int D, cudaStream_t st) {
    const int64_t rows = (int64_t)d.N * d.Lq * d.M;
    const int grid = (int)(rows < (int64_t)sm_count() * 64 ? rows : (int64_t)sm_count() * 64);
    msda_bwd_generic<T><<<grid, 128, 0, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d, D);
    return after_launch("msda_bwd_generic");
}

template <typename T>
int forward_any(const T *value, const int64_t *shapes, const int64_t *start, const T *loc, const T *attn, int N, int S,
                int M, int D, int L, int Lq, int P, T *out, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!value || !shapes || !start || !loc || !attn || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward: null pointer");
    const Dims d = make_dims(N, S, M, L, Lq, P);
    return forward_generic<T>(value, shapes, start, loc, attn, out, d, D, (cudaStream_t)stream);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }
const char *msda_last_error(void) { return g_err; }
uint64_t msda_launch_count(void) { return g_launches.load(); }

int msda_kernel_plan(int elem_bytes, int num_heads, int channels, int num_levels, int num_point) {
    (void)num_heads;
    return (elem_bytes == 2 || elem_bytes == 4) && tiled_ok(channels, num_levels, num_point) ? 1 : 0;
}

int msda_set_option(const char *key, int value) {
    if (!key) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_set_option: null key");
    if (!strcmp(key, "fwd_ctas_per_sm")) { g_fwd_ctas_per_sm = value; return MSDA_OK; }
    if (!strcmp(key, "bwd_ctas_per_sm")) { g_bwd_ctas_per_sm = value; return MSDA_OK; }
    if (!strcmp(key, "force_generic")) { g_force_generic = value; return MSDA_OK; }
    if (!strcmp(key, "force_linear_walk")) { g_force_linear = value; return MSDA_OK; }
    if (!strcmp(key, "debug_skip_scatter")) { g_skip_scatter = value; return MSDA_OK; }
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_set_option: unknown key");
}

int msda_forward_f32(const float *value, const int64_t *shapes, const int64_t *start, const float *loc,
                     const float *attn, int N, int S, int M, int D, int L, int Lq, int P, float *out,
                     msda_stream_t stream) {
    if (!tiled_ok(D, L, P)) return forward_any<float>(value, shapes, start, loc, attn, N, S, M, D, L, Lq, P, out, stream);
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!value || !shapes || !start || !loc || !attn || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward_f32: null pointer");
    if (misaligned(value, 16) || misaligned(out, 16) || misaligned(loc, 8))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward_f32: value/output must be 16-byte aligned, sampling_loc 8-byte aligned");
    const Dims d = make_dims(N, S, M, L, Lq, P);
    return launch_fwd_tiled<float>(value, shapes, start, loc, attn, out, d, (cudaStream_t)stream);
}

int msda_backward_f32(const float *go, const float *value, const int64_t *shapes, const int64_t *start,
                      const float *loc, const float *attn, int N, int S, int M, int D, int L, int Lq, int P,
                      float *gv, float *gl, float *ga, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (!gv && (int64_t)N > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_f32: null grad_value");
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        const cudaError_t e = cudaMemsetAsync(gv, 0, sizeof(float) * (size_t)N * S * M * D, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward_f32: memset(grad_value)");
    }
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!go || !value || !shapes || !start || !loc || !attn || !gl || !ga)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_f32: null pointer");
    const Dims d = make_dims(N, S, M, L, Lq, P);
    if (!tiled_ok(D, L, P)) return backward_generic<float>(go, value, shapes, start, loc, attn, gv, gl, ga, d, D, st);
    if (misaligned(value, 16) || misaligned(go, 16) || misaligned(gv, 16) || misaligned(loc, 8) || misaligned(gl, 8))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_f32: value/grad_output/grad_value must be 16-byte aligned, loc tensors 8-byte aligned");
    return launch_bwd_tiled<float>(go, value, shapes, start, loc, attn, gv, gl, ga, d, st);
}

int msda_forward_f64(const double *value, const int64_t *shapes, const int64_t *start, const double *loc,
                     const double *attn, int N, int S, int M, int D, int L, int Lq, int P, double *out,
                     msda_stream_t stream) {
    return forward_any<double>(value, shapes, start, loc, attn, N, S, M, D, L, Lq, P, out, stream);
}

int msda_backward_f64(const double *go, const double *value, const int64_t *shapes, const int64_t *start,
                      const double *loc, const double *attn, int N, int S, int M, int D, int L, int Lq, int P,
                      double *gv, double *gl, double *ga, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (!gv && (int64_t)N > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_f64: null grad_value");
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        const cudaError_t e = cudaMemsetAsync(gv, 0, sizeof(double) * (size_t)N * S * M * D, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward_f64: memset(grad_value)");
    }
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!go || !value || !shapes || !start || !loc || !attn || !gl || !ga)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_f64: null pointer");
    const Dims d = make_dims(N, S, M, L, Lq, P);
    return backward_generic<double>(go, value, shapes, start, loc, attn, gv, gl, ga, d, D, st);
}

int msda_forward_bf16(const uint16_t *value, const int64_t *shapes, const int64_t *start, const float *loc,
                      const float *attn, int N, int S, int M, int D, int L, int Lq, int P, uint16_t *out,
                      msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (!tiled_ok(D, L, P) || g_force_generic.load())
        return fail(MSDA_ERR_UNSUPPORTED, "msda_forward_bf16: only channels == 32, num_levels <= 16, num_levels*num_point <= 32");
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!value || !shapes || !start || !loc || !attn || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward_bf16: null pointer");
    if (misaligned(value, 8) || misaligned(out, 8) || misaligned(loc, 8))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward_bf16: value/output/sampling_loc must be 8-byte aligned");
    const Dims d = make_dims(N, S, M, L, Lq, P);
    return launch_fwd_tiled<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16 *>(value), shapes, start, loc, attn,
                                           reinterpret_cast<__nv_bfloat16 *>(out), d, (cudaStream_t)stream);
}

int msda_backward_bf16(const uint16_t *go, const uint16_t *value, const int64_t *shapes, const int64_t *start,
                       const float *loc, const float *attn, int N, int S, int M, int D, int L, int Lq, int P,
                       float *gv32, uint16_t *gv16, float *gl, float *ga, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (!tiled_ok(D, L, P) || g_force_generic.load())
        return fail(MSDA_ERR_UNSUPPORTED, "msda_backward_bf16: only channels == 32, num_levels <= 16, num_levels*num_point <= 32");
    if (!gv32 && N > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: null grad_value_f32");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nval = (int64_t)N * S * M * D;
    if (N > 0) {
        const cudaError_t e = cudaMemsetAsync(gv32, 0, sizeof(float) * (size_t)nval, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward_bf16: memset(grad_value)");
    }
    if ((int64_t)N * Lq > 0) {
        if (!go || !value || !shapes || !start || !loc || !attn || !gl || !ga)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: null pointer");
        if (misaligned(value, 8) || misaligned(go, 8) || misaligned(gv32, 16) || misaligned(loc, 8) || misaligned(gl, 8))
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: misaligned pointer");
        const Dims d = make_dims(N, S, M, L, Lq, P);
        if (const int rc = launch_bwd_tiled<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16 *>(go),
                                                           reinterpret_cast<const __nv_bfloat16 *>(value), shapes, start,
                                                           loc, attn, gv32, gl, ga, d, st))
            return rc;
    }
    if (gv16 && nval > 0) {
        if (misaligned(gv16, 8)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: misaligned grad_value_bf16");
        const int64_t n4 = nval / 4;   // D == 32, so nval is a multiple of 4
        const int grid = (int)((n4 + 255) / 256 < (int64_t)sm_count() * 16 ? (n4 + 255) / 256 : (int64_t)sm_count() * 16);
        msda_f32_to_bf16<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(gv32), reinterpret_cast<uint2 *>(gv16), n4);
        return after_launch("msda_f32_to_bf16");
    }
    return MSDA_OK;
}

}  // extern "C"
