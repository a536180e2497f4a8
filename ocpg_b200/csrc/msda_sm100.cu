// msda_sm100.cu -- multi-scale deformable attention forward / backward for B200 (sm_100a).
//
// Written from scratch for Blackwell; the operator it implements is the reference's
// models/ops/src/cuda/ms_deform_im2col_cuda.cuh (fwd :237-299, bwd :301-403, bilinear helpers :33-159)
// behind the C ABI of include/msda_sm100.h.  See DESIGN.md for the data layout and the roofline
// of each kernel.
//
// Why it looks the way it does.  Per (query, head) the op gathers L*P*4 = 64 rows of 32 channels (128 B each in
// fp32) and, backward, scatters 64 rows: 8 KB + 8 KB of SM <-> L1/L2 traffic for 448 B of compulsory HBM traffic.
// Neither is served by HBM.  Measured on this machine (tools/ubench_gather.cu, profiles/r1_ubench_*.jsonl):
//   - a four-row LDG.128 costs an SM 4.1 cycles when all rows hit L1 and 8.1 cycles when one misses; rows
//     straight from L2 stream at ~2 cycles each (64 B/clk/SM);
//   - fp32 reds are bounded CHIP-wide by the L2 atomic units: 6.4 TB/s however they are issued (v4, v2, scalar, TMA bulk
//     reduce; 37 or 148 SMs), and far less when many CTAs hit the same few rows;
//   - LDS, LDG hits AND warp shuffles share one 128-byte-per-clock data path per SM (tools/ubench_pipes.cu): it is what
//     both kernels end up bound by.
// So:
//   1. few, wide memory instructions: 8 lanes x 128-bit cover one 32-channel row, the four 8-lane groups of a
//      warp work on four x-adjacent queries, so one LDG.128 / RED.128 moves four rows;
//   2. few issue slots: the per-point geometry (pixel coordinates, bilinear weights, clamped corner offsets) is
//      computed once per point by one lane, staged in shared memory, and read back with two broadcast LDS.128;
//      every load is unpredicated and in bounds (zero padding comes out of the weights);
//   3. locality: a CTA pass covers an 8x8 (or 8x4) spatial tile of queries of ONE head, whose rows overlap (L1);
//   4. the backward's reds are predicated off for zero-weight corners, and the persistent CTAs are spread over
//      all (frame, head) slices at once so that the coarse levels' few rows do not serialise in L2;
//   5. the backward is ROW-major (msda_bwd_sorted.cuh): a CTA tile's (query, point, corner) items are counting-sorted
//      by row in shared memory, every distinct row is loaded once and receives ONE vector red per lane; the
//      query-major msda_bwd_tiled below (8-lane butterfly shuffles for the per-point reductions) serves the
//      under-filled launches of the decoder.
//
// No tensor cores: the op is a gather / scatter with ~0.2 flop per byte.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <type_traits>

#include "msda_sm100.h"

namespace {

// ------------------------------------------------------------------------------------------------
// host-side state: last error (per thread), launch counter, tuning knobs, cached SM counts
// ------------------------------------------------------------------------------------------------
thread_local char g_err[256] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_fwd_ctas_per_sm{0};   // 0 = kernel default
std::atomic<int> g_bwd_ctas_per_sm{0};
std::atomic<int> g_fwd_warps{0};          // 0 = default; 8 or 16 warps per CTA
std::atomic<int> g_bwd_warps{0};
#ifdef MSDA_EXPERIMENTS                  // measurement-only switches: never in the product build (they return wrong gradients)
std::atomic<int> g_bwd_mode{0};          // see Dims::bwd_mode
std::atomic<int> g_skip_scatter{0};      // see Dims::debug_skip_scatter
#endif
std::atomic<int> g_unit{0};               // 0 = automatic; frames interleaved by the task walk
std::atomic<int> g_force_generic{0};
std::atomic<int> g_force_linear{0};     // experiments: never use the tiled query walk
std::atomic<int> g_bwd_algo{0};         // 0 = automatic (row-major msda_bwd_sorted for encoder shapes, Lq == S, unless under-filled), 1 = query-major msda_bwd_tiled, 2 = row-major for any filled launch
std::atomic<int> g_bwd_deep{0};         // 0 = automatic (under-filled launches), 1 = always, -1 = never: msda_bwd_tiled<.., DEEP>

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
// A CTA of WARPS warps covers, per pass, a TILE_W x TILE_H patch of queries (4 x-adjacent queries per warp):
//   WARPS = 8: 8 x 4,   16: 8 x 8.
template <int WARPS> struct Tile {
    static constexpr int W = WARPS == 32 ? 16 : 8;
    static constexpr int H = WARPS * 4 / W;
    static constexpr int kQueries = WARPS * 4;
    static constexpr int kWarpsPerRow = W / 4;
};

struct Dims {
    int N, S, M, L, Lq, P;       // D is 32 for the tiled kernels
    int tiled;                   // 1: Lq == S, walk queries as spatial tiles of their own level
    int fchunk;                  // frames whose passes are interleaved (TaskWalk)
#ifdef MSDA_EXPERIMENTS
    int debug_skip_scatter;      // msda_bwd_tiled omits the grad_value reds (wrong grad_value)
    int bwd_mode;                // see msda_bwd_tiled
#endif
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
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// base + unit * sizeof(T) with a 32-bit unit offset in ONE instruction (IMAD.WIDE.U32); plain pointer
// arithmetic makes nvcc rebuild the 64-bit index and scale it (4 instructions per gathered row).
template <typename T> __device__ __forceinline__ T *row_at(T *base, uint32_t unit) {
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(unit), "n"((int)sizeof(T)), "l"(base));
    return reinterpret_cast<T *>(r);
}

// Packed fp32 pairs: Blackwell's FFMA2 / FMUL2 do two lanes of fp32 math per issue slot.
struct Row {            // four channels as two pairs
    float2 lo, hi;
};
__device__ __forceinline__ float2 splat(float s) { return make_float2(s, s); }
__device__ __forceinline__ void fma_row(float s, const Row &v, Row &acc) {
    acc.lo = __ffma2_rn(splat(s), v.lo, acc.lo);
    acc.hi = __ffma2_rn(splat(s), v.hi, acc.hi);
}
__device__ __forceinline__ float dot_row(const Row &a, const Row &b) {
    const float2 t = __ffma2_rn(a.hi, b.hi, __fmul2_rn(a.lo, b.lo));
    return t.x + t.y;
}

// Four channels of one (pixel, head) row as stored: a 16-byte (fp32) or 8-byte (bf16) vector.  All row
// offsets inside the kernels are counted in these vectors ("units"): the same unit index addresses `value`
// (in its own dtype) and the fp32 `grad_value`.
template <typename VT> struct RowIO;
template <> struct RowIO<float> {
    using Vec = float4;
    static __device__ __forceinline__ Row load(const Vec *p) {
        const float4 v = __ldg(p);
        return Row{make_float2(v.x, v.y), make_float2(v.z, v.w)};
    }
    static __device__ __forceinline__ Row load_stream(const Vec *p) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return Row{make_float2(v.x, v.y), make_float2(v.z, v.w)};
    }
    static __device__ __forceinline__ void store(Vec *p, const Row &r) { *p = make_float4(r.lo.x, r.lo.y, r.hi.x, r.hi.y); }
};
template <> struct RowIO<__nv_bfloat16> {
    using Vec = uint2;
    static __device__ __forceinline__ Row unpack(uint2 raw) {
        return Row{make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u)),
                   make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u))};
    }
    static __device__ __forceinline__ Row load(const Vec *p) { return unpack(__ldg(p)); }
    static __device__ __forceinline__ Row load_stream(const Vec *p) {
        uint2 raw;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "l"(p));
        return unpack(raw);
    }
    static __device__ __forceinline__ void store(Vec *p, const Row &r) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(r.lo.x, r.lo.y), hi = __floats2bfloat162_rn(r.hi.x, r.hi.y);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t *>(&lo);
        raw.y = *reinterpret_cast<uint32_t *>(&hi);
        *p = raw;
    }
};

// grad_value row (4 channels of this lane) += s * g: one REDG.E.ADD.F32x4 (sm_90+), no return value.
// `on` predicates the instruction: a corner whose weight is exactly zero (outside the level, zero
// attention, out-of-range point) sends nothing -- the reds are what bounds the backward.
__device__ __forceinline__ void red_row(float4 *p, float s, const Row &g, bool on) {
    const float2 a = __fmul2_rn(splat(s), g.lo), b = __fmul2_rn(splat(s), g.hi);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                 "@q red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "r"((uint32_t)on) : "memory");
}

// One sample point as the lanes that gather need it.
//   o[4] : unit offsets (inside the frame+head slice of `value`) of the corners (y0,x0), (y0,x0+1), (y0+1,x0),
//          (y0+1,x0+1).  Rows / columns that fall outside the level are CLAMPED onto the nearest valid one and
//          their bilinear weight is zeroed instead: every load is in bounds and unpredicated, and the zero
//          padding of the reference (cuh:56-78) comes out of the weights.
//   w[4] : bilinear weights hy*hx, hy*lx, ly*hx, ly*lx; zero for invalid corners and out-of-range points.
struct PointGeo {
    uint32_t o00, o01, o10, o11;
    float w00, w01, w10, w11;
    float lx, ly;
    int valid;      // bit i: corner i contributes (00, 01, 10, 11)
};

// Pixel coordinates exactly as the reference's compiled kernel forms them: fma(loc, size, -0.5)
// (ms_deform_im2col_cuda.cuh:285-286; nvcc contracts the expression -- SASS of the reference op built for
// sm_100a: `FFMA R29, R12, R29, -0.5` -- so borderline points fall into the same bilinear cell as there),
// range test of :288, corner tests of :56/:62/:68/:74.
__device__ __forceinline__ PointGeo point_geometry(float loc_x, float loc_y, int H, int W, int level_start,
                                                   int pixel_units) {
    PointGeo g;
    const float fw = (float)W, fh = (float)H;
    const float x = fmaf(loc_x, fw, -0.5f);
    const float y = fmaf(loc_y, fh, -0.5f);
    const bool in_range = (y > -1.f) && (x > -1.f) && (y < fh) && (x < fw);   // false for NaN / inf
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;     // saturating conversion (NaN -> 0); clamped below
    g.lx = in_range ? x - xf : 0.f;
    g.ly = in_range ? y - yf : 0.f;
    const float hx = 1.f - g.lx, hy = 1.f - g.ly;
    const int xc0 = min(max(x0, 0), W - 1), yc0 = min(max(y0, 0), H - 1);
    const int xc1 = min(max(x0, -1) + 1, W - 1), yc1 = min(max(y0, -1) + 1, H - 1);   // max first: no overflow at INT_MAX
    const bool xa = in_range && x0 >= 0, xb = in_range && x0 < W - 1;
    const bool ya = y0 >= 0, yb = y0 < H - 1;
    const int r0 = level_start + yc0 * W, r1 = level_start + yc1 * W;
    g.o00 = (uint32_t)((r0 + xc0) * pixel_units);
    g.o01 = (uint32_t)((r0 + xc1) * pixel_units);
    g.o10 = (uint32_t)((r1 + xc0) * pixel_units);
    g.o11 = (uint32_t)((r1 + xc1) * pixel_units);
    g.w00 = (xa && ya) ? hy * hx : 0.f;
    g.w01 = (xb && ya) ? hy * g.lx : 0.f;
    g.w10 = (xa && yb) ? g.ly * hx : 0.f;
    g.w11 = (xb && yb) ? g.ly * g.lx : 0.f;
    g.valid = (int)(xa && ya) | ((int)(xb && ya) << 1) | ((int)(xa && yb) << 2) | ((int)(xb && yb) << 3);
    return g;
}

template <int WARPS>
__device__ __forceinline__ void load_level_table(LevelTable &lt, const int64_t *shapes, const int64_t *start, int L,
                                                 int Lq) {
    constexpr int kTileW = Tile<WARPS>::W, kTileH = Tile<WARPS>::H;
    if (threadIdx.x < L) {
        lt.H[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        lt.W[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        lt.start[threadIdx.x] = (int)start[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cum = 0, rows = 0, dense = 1;
        for (int l = 0; l < L; ++l) {
            const int tx = (lt.W[l] + kTileW - 1) / kTileW, ty = (lt.H[l] + kTileH - 1) / kTileH;
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

// The query a lane group works on in pass `tile` of a frame.  Tiled walk (encoder, Lq == S and the levels
// partition the queries): the tile is an 8 x 4 patch of one level's pixels, so the rows its 8 warps gather
// overlap and stay in L1.  Linear walk otherwise: 32 consecutive queries.
template <int WARPS>
__device__ __forceinline__ int select_query(bool tiled, const Dims &d, const LevelTable &lt, int tile, int warp,
                                            int grp) {
    constexpr int kTileW = Tile<WARPS>::W, kTileH = Tile<WARPS>::H, kPerRow = Tile<WARPS>::kWarpsPerRow;
    constexpr int kTaskQueries = Tile<WARPS>::kQueries;
    if (tiled) {
        int lv = 0;
        while (lv + 1 < d.L && tile >= lt.tile_cum[lv + 1]) ++lv;
        const int t = tile - lt.tile_cum[lv];
        const int ty = t / lt.tiles_x[lv], tx = t - ty * lt.tiles_x[lv];
        const int y = ty * kTileH + warp / kPerRow, x = tx * kTileW + ((warp % kPerRow) << 2) + grp;
        return (y < lt.H[lv] && x < lt.W[lv]) ? lt.start[lv] + y * lt.W[lv] + x : -1;
    }
    const int q = tile * kTaskQueries + warp * 4 + grp;
    return q < d.Lq ? q : -1;
}

// Persistent task walk shared by both kernels: worker `rank` of `workers` takes passes rank, rank + workers, ...
// (the grid is exactly SMs x resident CTAs per SM: one wave).  A pass is (frame n, head m, tile).  Passes are
// ordered so that the CTAs running at the same time work on DIFFERENT (frame, head) slices:
//     frames are taken `fchunk` at a time (as many as keep their `value` slices in L2 together); inside a chunk
//     the slice (frame, head) runs fastest, then the tile.
// Why: the backward's reds into a coarse pyramid level (60 .. 720 rows per slice, a quarter of all reds each) are
// serialised per address by the L2 atomic units -- 0.96 TB/s when every CTA hits the same 64 rows against
// 6.4 TB/s when they are spread over 32k rows (profiles/r1_ubench_gather_b200_part3_hotspots.jsonl).
struct TaskWalk {
    int n, m, tile;        // current pass
    uint32_t p, stride, total, tiles, M, N, fchunk;
    __device__ __forceinline__ TaskWalk(int N_, int M_, int tiles_, int fchunk_, int rank, int workers)
        : n(0), m(0), tile(0), p((uint32_t)rank), stride((uint32_t)workers), total((uint32_t)N_ * M_ * tiles_),
          tiles((uint32_t)tiles_), M((uint32_t)M_), N((uint32_t)N_), fchunk((uint32_t)fchunk_) {}
    // advance to the next pass; false when this worker is done
    __device__ __forceinline__ bool next() {
        if (p >= total) return false;
        const uint32_t per_chunk = fchunk * M * tiles;
        const uint32_t c = p / per_chunk, rem = p - c * per_chunk;
        const uint32_t f0 = c * fchunk, frames = min(fchunk, N - f0);      // the last chunk may be short
        const uint32_t slices = frames * M;
        const uint32_t t = rem / slices, sl = rem - t * slices;
        const uint32_t f = sl / M;
        tile = (int)t;
        n = (int)(f0 + f);
        m = (int)(sl - f * M);
        p += stride;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// Fused module path: the elementwise work MSDeformAttn.forward does around the
// operator -- softmax of the attention logits over the L*P points (ms_deform_attn.py:101-102) and
// sampling_locations = reference_points + offsets / (W_l, H_l) (2-d references, :104-107) or
// reference_xy + offsets / P * reference_wh * 0.5 (4-d boxes, :108-110) -- folded into the staging phase of
// the kernels, and its backward (softmax gradient, offset scaling) into the backward's last stage.  In this
// mode `loc` holds the raw offsets and `attn` the raw logits.  The arithmetic keeps torch's operation order
// (true divisions, no contraction) so the locations are bit-identical to the unfused module's.
// ------------------------------------------------------------------------------------------------
struct FusedArgs {
    const float *ref;        // [N][Lq][L][ref_dim] reference points
    int ref_dim;             // 2 or 4
    float *loc_out;          // forward, optional: sampling_locations [N][Lq][M][L][P][2]
    float *attn_out;         // forward, optional: attention_weights  [N][Lq][M][L][P]
    float *grad_loc_out;     // backward, optional: d/d sampling_locations (the caller sums it into grad_reference_points)
};

__device__ __forceinline__ float group_max(float v) {      // over the 8 lanes of a group
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
}
__device__ __forceinline__ float group_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

// In: xy[r] = raw offsets, a[r] = raw logits of this lane's points (8*r + cl) of query `nq`.  Out: xy[r] = sampling
// locations, a[r] = softmax probabilities (0 for lanes without a point).  Must be called by all 32 lanes.
template <int ROUNDS>
__device__ __forceinline__ void fused_softmax_and_locations(float2 (&xy)[ROUNDS], float (&a)[ROUNDS], bool query_on, int cl,
                                                            int pts, uint32_t lv, const LevelTable &lt, const Dims &d,
                                                            const FusedArgs &fa, int64_t nq, int64_t row) {
    float mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r)
        if (query_on && 8 * r + cl < pts) mx = fmaxf(mx, a[r]);
    mx = group_max(mx);
    float e[ROUNDS], sum = 0.f;
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        e[r] = (query_on && 8 * r + cl < pts) ? expf(a[r] - mx) : 0.f;
        sum += e[r];
    }
    sum = group_sum(sum);
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const bool on = query_on && 8 * r + cl < pts;
        a[r] = on ? __fdiv_rn(e[r], sum) : 0.f;
        if (!on) continue;
        const int l = (lv >> (8 * r)) & 0xff;
        if (fa.ref_dim == 2) {
            const float2 rp = __ldg(reinterpret_cast<const float2 *>(fa.ref) + (nq * d.L + l));
            xy[r].x = __fadd_rn(rp.x, __fdiv_rn(xy[r].x, (float)lt.W[l]));
            xy[r].y = __fadd_rn(rp.y, __fdiv_rn(xy[r].y, (float)lt.H[l]));
        } else {
            const float4 rp = __ldg(reinterpret_cast<const float4 *>(fa.ref) + (nq * d.L + l));
            const float fp = (float)d.P;
            xy[r].x = __fadd_rn(rp.x, __fmul_rn(__fmul_rn(__fdiv_rn(xy[r].x, fp), rp.z), 0.5f));
            xy[r].y = __fadd_rn(rp.y, __fmul_rn(__fmul_rn(__fdiv_rn(xy[r].y, fp), rp.w), 0.5f));
        }
        if (fa.loc_out) *reinterpret_cast<float2 *>(fa.loc_out + (row * pts + 8 * r + cl) * 2) = xy[r];
        if (fa.attn_out) fa.attn_out[row * pts + 8 * r + cl] = a[r];
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled forward, D = 32.  ROUNDS = ceil(L*P / 8): lane `cl` of a group stages points cl, 8+cl, ... of the
// group's query (slots past L*P carry zero weights), then every lane gathers all of them branch-free.
//   per point and warp: 2 broadcast LDS.128 (weights, offsets) + 4 LDG.128 (four rows each) + 8 FFMA2.
// ------------------------------------------------------------------------------------------------
template <int ROUNDS, int WARPS> struct FwdSmem {
    static constexpr int kSlots = ROUNDS * 8 + 1;   // +1: the four groups of a warp read four different banks
    LevelTable lt;
    float4 w[WARPS][4][kSlots];   // a*w00, a*w01, a*w10, a*w11
    uint4 o[WARPS][4][kSlots];    // unit offsets of the four corners
};
extern __shared__ __align__(16) unsigned char msda_smem[];

template <typename VT, int ROUNDS, int WARPS, bool FUSED>
#ifndef MSDA_FWD_WARPS_PER_SM
#define MSDA_FWD_WARPS_PER_SM 32      // resident forward warps per SM the register budget is set for (32 -> 64 registers)
#endif
__global__ void __launch_bounds__(WARPS * 32, MSDA_FWD_WARPS_PER_SM / WARPS)
msda_fwd_tiled(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
               const float *__restrict__ loc, const float *__restrict__ attn, VT *__restrict__ out, Dims d, FusedArgs fa) {
    using IO = RowIO<VT>;
    using Vec = typename IO::Vec;
    constexpr int kTaskQueries = Tile<WARPS>::kQueries;
    FwdSmem<ROUNDS, WARPS> &sm = *reinterpret_cast<FwdSmem<ROUNDS, WARPS> *>(msda_smem);
    LevelTable &lt = sm.lt;
    auto &s_w = sm.w;
    auto &s_o = sm.o;
    load_level_table<WARPS>(lt, shapes, start, d.L, d.Lq);
    const bool tiled = d.tiled && lt.dense;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int pts = d.L * d.P;
    const int pixel_units = d.M * 8;
    uint32_t lv = 0;     // level of this lane's point in each round, one byte per round
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) lv |= (uint32_t)min((8 * r + cl) / d.P, d.L - 1) << (8 * r);

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kTaskQueries - 1) / kTaskQueries;
    for (TaskWalk t(d.N, d.M, tiles, d.fchunk, blockIdx.x, gridDim.x); t.next();) {
        const int m = t.m;
        const int q = select_query<WARPS>(tiled, d, lt, t.tile, warp, grp);
        if (!__any_sync(0xffffffffu, q >= 0)) continue;
        const int64_t row = ((int64_t)t.n * d.Lq + max(q, 0)) * d.M + m;           // (n, q, m)
        // stage all points of the four queries
        float2 xy[ROUNDS];
        float a[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const bool on = q >= 0 && 8 * r + cl < pts;
            xy[r] = make_float2(-4.f, -4.f);        // out of range: zero weights, offsets clamped in bounds
            a[r] = 0.f;
            if (on) {
                xy[r] = ld_stream_f2(loc + (row * pts + 8 * r + cl) * 2);
                a[r] = ld_stream_f1(attn + row * pts + 8 * r + cl);
            }
        }
        if constexpr (FUSED)
            fused_softmax_and_locations<ROUNDS>(xy, a, q >= 0, cl, pts, lv, lt, d, fa, (int64_t)t.n * d.Lq + max(q, 0), row);
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int l = (lv >> (8 * r)) & 0xff;
            const PointGeo g = point_geometry(xy[r].x, xy[r].y, lt.H[l], lt.W[l], lt.start[l], pixel_units);
            const float aa = g.valid ? a[r] : 0.f;      // an out-of-range point ignores its weight (cuh:288)
            s_w[warp][grp][8 * r + cl] = make_float4(aa * g.w00, aa * g.w01, aa * g.w10, aa * g.w11);
            s_o[warp][grp][8 * r + cl] = make_uint4(g.o00, g.o01, g.o10, g.o11);
        }
        __syncwarp();
        // frame n, head m, this lane's four channels
        const Vec *vb = reinterpret_cast<const Vec *>(value) + (((int64_t)t.n * d.S * d.M + m) * 8 + cl);
        Row acc{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        // One LDG.128 per corner moves four rows (one per lane group).  What bounds this loop is the L1's request
        // rate: 4.1 cycles for a four-row request that hits, 8.1 as soon as ONE row misses (measured,
        // profiles/r1_ubench_gather_b200_part2.jsonl) -- deeper software pipelining / more registers per thread
        // change nothing (tried: 64 .. 255 registers, 1 .. 16 points of loads in flight per warp).
#pragma unroll
        for (int it = 0; it < ROUNDS * 8; ++it) {
            const float4 pw = s_w[warp][grp][it];
            const uint4 po = s_o[warp][grp][it];
            const Row v00 = IO::load(row_at(vb, po.x));
            const Row v01 = IO::load(row_at(vb, po.y));
            const Row v10 = IO::load(row_at(vb, po.z));
            const Row v11 = IO::load(row_at(vb, po.w));
            fma_row(pw.x, v00, acc);
            fma_row(pw.y, v01, acc);
            fma_row(pw.z, v10, acc);
            fma_row(pw.w, v11, acc);
        }
        __syncwarp();
        if (q >= 0) IO::store(reinterpret_cast<Vec *>(out) + (row * 8 + cl), acc);
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled backward, D = 32.  Same task walk as the forward.  grad_value (fp32) must be zero on entry.
//
// Per point and per lane (4 channels) the gather loop forms the four corner dot products
//     p_i = sum_c grad_out[c] * v_i[c]
// and sends the scatter rows (a * w_i) * grad_out as predicated vector reds.  The partial p_i of the 8
// lanes of a group are transposed through shared memory (one STS.128 per point, then the lane that
// owns the point adds the 8 partials), and that lane finishes with scalars (cuh:123-158 regrouped by corner):
//     grad_attn = sum_i w_i p_i
//     grad_x    = W * a * ( hy (p01 - p00) + ly (p11 - p10) )      (invalid corners dropped)
//     grad_y    = H * a * ( hx (p10 - p00) + lx (p11 - p01) )
// ------------------------------------------------------------------------------------------------
template <int WARPS> struct BwdSmem {
    LevelTable lt;
    float4 w[WARPS][4][9];      // a*w00, a*w01, a*w10, a*w11   (scatter weights), one round
    uint4 o[WARPS][4][9];       // unit offsets of the four corners
    float4 own[WARPS][32][2];   // per lane: its own point's bilinear weights; lx, ly, attention, corner validity
};

#ifndef MSDA_BWD_MINB8
#define MSDA_BWD_MINB8 3      // resident 8-warp backward CTAs per SM the register budget is set for (3 -> 85 registers)
#endif
#ifndef MSDA_BWD_MINB16
#define MSDA_BWD_MINB16 2
#endif
template <typename VT, int ROUNDS, int WARPS, bool FUSED, bool DEEP>
__global__ void __launch_bounds__(WARPS * 32, DEEP ? 1 : (WARPS == 8 ? MSDA_BWD_MINB8 : MSDA_BWD_MINB16))
msda_bwd_tiled(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
               const int64_t *__restrict__ start, const float *__restrict__ loc, const float *__restrict__ attn,
               float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn, Dims d,
               FusedArgs fa) {
    using IO = RowIO<VT>;
    using Vec = typename IO::Vec;
    constexpr int kTaskQueries = Tile<WARPS>::kQueries;
    BwdSmem<WARPS> &sm = *reinterpret_cast<BwdSmem<WARPS> *>(msda_smem);
    LevelTable &lt = sm.lt;
    auto &s_w = sm.w;
    auto &s_o = sm.o;
    auto &s_own = sm.own;
    load_level_table<WARPS>(lt, shapes, start, d.L, d.Lq);
    const bool tiled = d.tiled && lt.dense;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int pts = d.L * d.P;
    const int pixel_units = d.M * 8;
    // Dims::bwd_mode (measurement only): 1 = gather / reduce without the scatter, 2 = the scatter alone (no `value`
    // traffic at all).  Together they show what bounds the kernel: the scatter alone takes ~85 % of the full
    // backward -- the L2 atomic units, chip-wide.  0 = the real thing.
#ifdef MSDA_EXPERIMENTS
    const bool scatter = d.bwd_mode != 1 && !d.debug_skip_scatter, gather = d.bwd_mode != 2;
#else
    constexpr bool scatter = true, gather = true;
#endif
    const int rank = blockIdx.x, workers = gridDim.x;
    uint32_t lv = 0;     // level of this lane's point in each round, one byte per round
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) lv |= (uint32_t)min((8 * r + cl) / d.P, d.L - 1) << (8 * r);

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kTaskQueries - 1) / kTaskQueries;
    for (TaskWalk t(d.N, d.M, tiles, d.fchunk, rank, workers); t.next();) {
        const int m = t.m;
        const int q = select_query<WARPS>(tiled, d, lt, t.tile, warp, grp);
        if (!__any_sync(0xffffffffu, q >= 0)) continue;
        const int64_t row = ((int64_t)t.n * d.Lq + max(q, 0)) * d.M + m;
        const int64_t slice = ((int64_t)t.n * d.S * d.M + m) * 8 + cl;      // unit offset of frame n, head m, this lane
        const Vec *vb = reinterpret_cast<const Vec *>(value) + slice;
        float4 *gb = reinterpret_cast<float4 *>(grad_value) + slice;
        float2 xy[ROUNDS];
        float a[ROUNDS];
        Row go{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (q >= 0) go = IO::load_stream(reinterpret_cast<const Vec *>(grad_out) + (row * 8 + cl));
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const bool on = q >= 0 && 8 * r + cl < pts;
            xy[r] = make_float2(-4.f, -4.f);
            a[r] = 0.f;
            if (on) {
                xy[r] = ld_stream_f2(loc + (row * pts + 8 * r + cl) * 2);
                a[r] = ld_stream_f1(attn + row * pts + 8 * r + cl);
            }
        }
        const int64_t nq = (int64_t)t.n * d.Lq + max(q, 0);
        if constexpr (FUSED) {
            FusedArgs fwd_only = fa;       // the backward re-derives probabilities and locations; it emits neither
            fwd_only.loc_out = nullptr; fwd_only.attn_out = nullptr;
            fused_softmax_and_locations<ROUNDS>(xy, a, q >= 0, cl, pts, lv, lt, d, fwd_only, nq, row);
        }
        float g_prob[ROUNDS];      // fused mode: d/d probability of this lane's points, kept for the softmax gradient
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) g_prob[r] = 0.f;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int pt = 8 * r + cl;
            const int l = (lv >> (8 * r)) & 0xff;
            {
                const PointGeo g = point_geometry(xy[r].x, xy[r].y, lt.H[l], lt.W[l], lt.start[l], pixel_units);
                const float aa = g.valid ? a[r] : 0.f;      // an out-of-range point ignores its weight (cuh:365)
                s_w[warp][grp][cl] = make_float4(aa * g.w00, aa * g.w01, aa * g.w10, aa * g.w11);
                s_o[warp][grp][cl] = make_uint4(g.o00, g.o01, g.o10, g.o11);
                // what this lane needs again after the gather loop, parked in shared memory (registers are
                // what bounds the loads in flight)
                s_own[warp][lane][0] = make_float4(g.w00, g.w01, g.w10, g.w11);
                s_own[warp][lane][1] = make_float4(g.lx, g.ly, aa, __int_as_float(g.valid));
            }
            __syncwarp();
            if (!gather) {
                // scatter-only role: grad_value[corner] += (a * w_corner) * grad_out       (cuh:125,134,143,152)
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const float4 pw = s_w[warp][grp][it];
                    const uint4 po = s_o[warp][grp][it];
                    red_row(row_at(gb, po.x), pw.x, go, scatter && pw.x != 0.f);
                    red_row(row_at(gb, po.y), pw.y, go, scatter && pw.y != 0.f);
                    red_row(row_at(gb, po.z), pw.z, go, scatter && pw.z != 0.f);
                    red_row(row_at(gb, po.w), pw.w, go, scatter && pw.w != 0.f);
                }
                __syncwarp();
                continue;
            }
            // One step = one point: 4 row loads, 4 predicated reds, the lane's partial dot products (its 4 channels).
            auto step_rows = [&](int it, Row (&v)[4]) {
                const uint4 po = s_o[warp][grp][it];
                v[0] = IO::load(row_at(vb, po.x));
                v[1] = IO::load(row_at(vb, po.y));
                v[2] = IO::load(row_at(vb, po.z));
                v[3] = IO::load(row_at(vb, po.w));
            };
            auto step_finish = [&](int it, const Row (&v)[4]) -> float4 {
                const float4 pw = s_w[warp][grp][it];
                const uint4 po = s_o[warp][grp][it];
                // scatter: grad_value[corner] += (a * w_corner) * grad_out                  (cuh:125,134,143,152)
                red_row(row_at(gb, po.x), pw.x, go, scatter && pw.x != 0.f);
                red_row(row_at(gb, po.y), pw.y, go, scatter && pw.y != 0.f);
                red_row(row_at(gb, po.z), pw.z, go, scatter && pw.z != 0.f);
                red_row(row_at(gb, po.w), pw.w, go, scatter && pw.w != 0.f);
                return make_float4(dot_row(go, v[0]), dot_row(go, v[1]), dot_row(go, v[2]), dot_row(go, v[3]));
            };
            // Reduce-scatter of the partials over the 8 lanes of a group by butterfly shuffles: after the exchange
            // with mask 4, 2, 1 lane `cl` holds the sums of point `cl`.  No shared memory: the 32 KB transposition
            // buffer of the first version left the L1 at its 28 KB minimum (four 50 KB CTAs per SM) and the gather
            // at a 38 % hit rate.  exchange(): the half of the group with the mask bit clear keeps A, the other keeps B.
            auto exchange = [&](const float4 &A, const float4 &B, int mask) -> float4 {
                const bool up = (cl & mask) != 0;
                const float4 send = up ? A : B;
                float4 keep = up ? B : A;
                keep.x += __shfl_xor_sync(0xffffffffu, send.x, mask);
                keep.y += __shfl_xor_sync(0xffffffffu, send.y, mask);
                keep.z += __shfl_xor_sync(0xffffffffu, send.z, mask);
                keep.w += __shfl_xor_sync(0xffffffffu, send.w, mask);
                return keep;
            };
            float4 qs;
            if constexpr (DEEP) {
                // Under-filled launch (the decoder's handful of queries): one or two warps per SM, so nothing hides a
                // row load's latency except the warp's own loads -- issue all 32 row loads of the round before the first
                // use (128 registers; this instantiation runs one CTA per SM), then scatter and reduce.
                Row v[8][4];
#pragma unroll
                for (int it = 0; it < 8; ++it) step_rows(it, v[it]);
                float4 k1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) k1[j] = exchange(step_finish(j, v[j]), step_finish(j + 4, v[j + 4]), 4);
                qs = exchange(exchange(k1[0], k1[2], 2), exchange(k1[1], k1[3], 2), 1);
            } else {
                // points in the order (0,4) (2,6) | (1,5) (3,7): each pair is exchanged as soon as it is complete, so at
                // most two reduced float4 are live next to the point being gathered
                auto pair = [&](int j) -> float4 {
                    Row v[4];
                    step_rows(j, v);
                    const float4 A = step_finish(j, v);
                    step_rows(j + 4, v);
                    const float4 B = step_finish(j + 4, v);
                    return exchange(A, B, 4);
                };
                const float4 e0 = pair(0);
                const float4 k2a = exchange(e0, pair(2), 2);
                const float4 e1 = pair(1);
                qs = exchange(k2a, exchange(e1, pair(3), 2), 1);
            }
            if (q >= 0 && pt < pts) {
                const float4 ow = s_own[warp][lane][0], og = s_own[warp][lane][1];
                const float lx = og.x, ly = og.y, a_own = og.z;
                const int valid = __float_as_int(og.w);
                // clamped (invalid) corners carry someone else's row: drop them (zero padding, cuh:56-78)
                const float q00 = (valid & 1) ? qs.x : 0.f, q01 = (valid & 2) ? qs.y : 0.f;
                const float q10 = (valid & 4) ? qs.z : 0.f, q11 = (valid & 8) ? qs.w : 0.f;
                const float hx = 1.f - lx, hy = 1.f - ly;
                const float ga = ow.x * q00 + ow.y * q01 + ow.z * q10 + ow.w * q11;                   // :156
                const float gx = hy * (q01 - q00) + ly * (q11 - q10);                                 // :157
                const float gy = hx * (q10 - q00) + lx * (q11 - q01);                                 // :158
                const float2 gl = make_float2((float)lt.W[l] * a_own * gx, (float)lt.H[l] * a_own * gy);
                if constexpr (!FUSED) {
                    grad_attn[row * pts + pt] = ga;
                    *reinterpret_cast<float2 *>(grad_loc + (row * pts + pt) * 2) = gl;
                } else {
                    // d/d offsets through sampling_locations = ref + offsets / (W, H)            (ms_deform_attn.py:104-107)
                    //                          or           = ref_xy + offsets / P * ref_wh * 0.5   (:108-110)
                    g_prob[r] = ga;
                    if (fa.grad_loc_out) *reinterpret_cast<float2 *>(fa.grad_loc_out + (row * pts + pt) * 2) = gl;
                    float2 go_;
                    if (fa.ref_dim == 2) {
                        go_ = make_float2(__fdiv_rn(gl.x, (float)lt.W[l]), __fdiv_rn(gl.y, (float)lt.H[l]));
                    } else {
                        const float4 rp = __ldg(reinterpret_cast<const float4 *>(fa.ref) + (nq * d.L + l));
                        const float fp = (float)d.P;
                        go_ = make_float2(__fdiv_rn(__fmul_rn(__fmul_rn(gl.x, 0.5f), rp.z), fp),
                                          __fdiv_rn(__fmul_rn(__fmul_rn(gl.y, 0.5f), rp.w), fp));
                    }
                    *reinterpret_cast<float2 *>(grad_loc + (row * pts + pt) * 2) = go_;
                }
            }
            // the staging slots are rewritten by the next round: order this round's shared-memory reads before those
            // writes (the shuffles above give no memory ordering)
            __syncwarp();
        }
        if constexpr (FUSED) {
            // softmax gradient: d logit_i = p_i * (g_i - sum_j p_j g_j) over the L*P points of this (query, head)
            float dot = 0.f;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) dot = fmaf(a[r], g_prob[r], dot);
            dot = group_sum(dot);
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r)
                if (q >= 0 && 8 * r + cl < pts) grad_attn[row * pts + 8 * r + cl] = a[r] * (g_prob[r] - dot);
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

#include "msda_bwd_sorted.cuh"
#include "msda_epilogue.cuh"
#include "msda_decoder.cuh"
#include "msda_flatten.cuh"

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
// frames worked on at the same time: as many as keep value (+ grad_value) slices of ~48 MB in the 126 MB L2
int frame_chunk(int N, int S, int M, int D, int elem_bytes) {
    const int k = g_unit.load();
    if (k > 0) return k < N ? k : (N > 0 ? N : 1);
    const int64_t frame_bytes = (int64_t)S * M * D * elem_bytes;
    int64_t f = (48ll << 20) / (frame_bytes > 0 ? frame_bytes : 1);
    if (f < 1) f = 1;
    if (f > N) f = N;
    return (int)(f > 0 ? f : 1);
}

Dims make_dims(int N, int S, int M, int D, int L, int Lq, int P, int elem_bytes) {
    Dims d{};
    d.N = N; d.S = S; d.M = M; d.L = L; d.Lq = Lq; d.P = P;
    d.tiled = (Lq == S && !g_force_linear.load()) ? 1 : 0;
    d.fchunk = frame_chunk(N, S, M, D, elem_bytes);
#ifdef MSDA_EXPERIMENTS
    d.debug_skip_scatter = g_skip_scatter.load();
    d.bwd_mode = g_bwd_mode.load();
#endif
    return d;
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

// grid = (slots, M): enough CTAs for `ctas_per_sm` residents on every SM, split over the heads
constexpr int kDefaultFwdWarps = 16, kDefaultBwdWarps = 8;    // measured best on the A2D / YTVOS encoder shapes
int warps_for(const std::atomic<int> &knob, int dflt) {
    const int k = knob.load();
    return (k == 8 || k == 16) ? k : dflt;
}

// exactly one wave: SMs x resident CTAs per SM
int grid_for(int ctas_per_sm_default, const std::atomic<int> &knob) {
    const int k = knob.load();
    return sm_count() * (k > 0 ? k : ctas_per_sm_default);
}

int after_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MSDA_OK : fail_cuda(e, what);
}

template <typename K>
int configure(K kernel, size_t smem) {
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return MSDA_OK;
}

// `fa == nullptr`: the plain operator; otherwise the fused module path (FusedArgs).
template <typename VT, int ROUNDS, int WARPS>
int launch_fwd_one(const VT *value, const int64_t *shapes, const int64_t *start, const float *loc, const float *attn,
                   VT *out, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    const size_t smem = sizeof(FwdSmem<ROUNDS, WARPS>);
    const int grid = grid_for(MSDA_FWD_WARPS_PER_SM / WARPS, g_fwd_ctas_per_sm);
    if (fa) {
        if (const int rc = configure(msda_fwd_tiled<VT, ROUNDS, WARPS, true>, smem)) return rc;
        msda_fwd_tiled<VT, ROUNDS, WARPS, true><<<grid, WARPS * 32, smem, st>>>(value, shapes, start, loc, attn, out, d, *fa);
    } else {
        if (const int rc = configure(msda_fwd_tiled<VT, ROUNDS, WARPS, false>, smem)) return rc;
        msda_fwd_tiled<VT, ROUNDS, WARPS, false><<<grid, WARPS * 32, smem, st>>>(value, shapes, start, loc, attn, out, d, FusedArgs{});
    }
    return after_launch("msda_fwd_tiled");
}

template <typename VT, int ROUNDS>
int launch_fwd_rounds(const VT *value, const int64_t *shapes, const int64_t *start, const float *loc, const float *attn,
                      VT *out, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    switch (warps_for(g_fwd_warps, kDefaultFwdWarps)) {
        case 8: return launch_fwd_one<VT, ROUNDS, 8>(value, shapes, start, loc, attn, out, d, fa, st);
        default: return launch_fwd_one<VT, ROUNDS, 16>(value, shapes, start, loc, attn, out, d, fa, st);
    }
}

template <typename VT>
int launch_fwd_tiled(const VT *value, const int64_t *shapes, const int64_t *start, const float *loc, const float *attn,
                     VT *out, const Dims &d, cudaStream_t st, const FusedArgs *fa = nullptr) {
    switch ((d.L * d.P + 7) / 8) {
        case 1: return launch_fwd_rounds<VT, 1>(value, shapes, start, loc, attn, out, d, fa, st);
        case 2: return launch_fwd_rounds<VT, 2>(value, shapes, start, loc, attn, out, d, fa, st);
        case 3: return launch_fwd_rounds<VT, 3>(value, shapes, start, loc, attn, out, d, fa, st);
        default: return launch_fwd_rounds<VT, 4>(value, shapes, start, loc, attn, out, d, fa, st);
    }
}

template <typename VT, int ROUNDS, int WARPS, bool DEEP>
int launch_bwd_kernel(int grid, const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                      const float *attn, float *gv, float *gl, float *ga, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    const size_t smem = sizeof(BwdSmem<WARPS>);
    if (fa) {
        if (const int rc = configure(msda_bwd_tiled<VT, ROUNDS, WARPS, true, DEEP>, smem)) return rc;
        msda_bwd_tiled<VT, ROUNDS, WARPS, true, DEEP><<<grid, WARPS * 32, smem, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d, *fa);
    } else {
        if (const int rc = configure(msda_bwd_tiled<VT, ROUNDS, WARPS, false, DEEP>, smem)) return rc;
        msda_bwd_tiled<VT, ROUNDS, WARPS, false, DEEP><<<grid, WARPS * 32, smem, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d, FusedArgs{});
    }
    return after_launch("msda_bwd_tiled");
}

// The row-major backward (msda_bwd_sorted.cuh): one wave of as many CTAs as fit (registers / ~51 KB of shared memory each).
template <typename K>
int resident_ctas(K kernel, int threads, size_t smem, int fallback) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm <= 0) per_sm = fallback;
    return per_sm;
}
template <typename VT, int ROUNDS, bool GV16 = false>
int launch_bwd_sorted(const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                      const float *attn, void *gv, float *gl, float *ga, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    const size_t smem = sizeof(SortSmem<ROUNDS>);
    const int knob = g_bwd_ctas_per_sm.load();
    if (fa) {
        auto kernel = msda_bwd_sorted<VT, ROUNDS, true, GV16>;
        if (const int rc = configure(kernel, smem)) return rc;
        const int grid = sm_count() * (knob > 0 ? knob : resident_ctas(kernel, 256, smem, 1));
        kernel<<<grid, 256, smem, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d, *fa);
    } else {
        auto kernel = msda_bwd_sorted<VT, ROUNDS, false, GV16>;
        if (const int rc = configure(kernel, smem)) return rc;
        const int grid = sm_count() * (knob > 0 ? knob : resident_ctas(kernel, 256, smem, 1));
        kernel<<<grid, 256, smem, st>>>(go, value, shapes, start, loc, attn, gv, gl, ga, d, FusedArgs{});
    }
    return after_launch("msda_bwd_sorted");
}

// bf16 value with grad_value accumulated directly in bf16 (packed reds): the row-major kernel only
int launch_bwd_sorted_gv16(const __nv_bfloat16 *go, const __nv_bfloat16 *value, const int64_t *shapes, const int64_t *start,
                           const float *loc, const float *attn, uint16_t *gv16, float *gl, float *ga, const Dims &d,
                           const FusedArgs *fa, cudaStream_t st) {
    switch ((d.L * d.P + 7) / 8) {
        case 1: return launch_bwd_sorted<__nv_bfloat16, 1, true>(go, value, shapes, start, loc, attn, gv16, gl, ga, d, fa, st);
        case 2: return launch_bwd_sorted<__nv_bfloat16, 2, true>(go, value, shapes, start, loc, attn, gv16, gl, ga, d, fa, st);
        case 3: return launch_bwd_sorted<__nv_bfloat16, 3, true>(go, value, shapes, start, loc, attn, gv16, gl, ga, d, fa, st);
        default: return launch_bwd_sorted<__nv_bfloat16, 4, true>(go, value, shapes, start, loc, attn, gv16, gl, ga, d, fa, st);
    }
}

// The decoder's shapes (a handful of queries per frame) give fewer passes than one wave has CTAs: those launches use
// the DEEP instantiation (8 warps, one CTA per SM, every row load of a round in flight at once).
template <typename VT, int ROUNDS, int WARPS>
int launch_bwd_one(const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                   const float *attn, float *gv, float *gl, float *ga, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    const int grid = grid_for(WARPS == 8 ? MSDA_BWD_MINB8 : MSDA_BWD_MINB16, g_bwd_ctas_per_sm);
    if constexpr (WARPS == 8) {
        const int64_t passes = (int64_t)d.N * d.M * ((d.Lq + Tile<8>::kQueries - 1) / Tile<8>::kQueries);   // linear walk
        const int deep = g_bwd_deep.load();
        if (deep > 0 || (deep == 0 && passes <= 2 * (int64_t)sm_count()))
            return launch_bwd_kernel<VT, ROUNDS, 8, true>(sm_count(), go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
    }
    if constexpr (WARPS == 8) {
        // Row-major kernel for the encoder's self-attention (Lq == S: neighbouring queries sample neighbouring rows, which is
        // what its per-tile sort merges); object queries (Lq != S) have no such locality: query-major kernel.  bwd_algo = 2
        // forces the row-major kernel for any shape.
        const int algo = g_bwd_algo.load();
        if (algo != 1 && d.L <= kSortLevels && (d.tiled || algo == 2)) return launch_bwd_sorted<VT, ROUNDS>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
    }
    return launch_bwd_kernel<VT, ROUNDS, WARPS, false>(grid, go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
}

template <typename VT, int ROUNDS>
int launch_bwd_rounds(const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                      const float *attn, float *gv, float *gl, float *ga, const Dims &d, const FusedArgs *fa, cudaStream_t st) {
    switch (warps_for(g_bwd_warps, kDefaultBwdWarps)) {
        case 8: return launch_bwd_one<VT, ROUNDS, 8>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
        default: return launch_bwd_one<VT, ROUNDS, 16>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
    }
}

template <typename VT>
int launch_bwd_tiled(const VT *go, const VT *value, const int64_t *shapes, const int64_t *start, const float *loc,
                     const float *attn, float *gv, float *gl, float *ga, const Dims &d, cudaStream_t st,
                     const FusedArgs *fa = nullptr) {
    switch ((d.L * d.P + 7) / 8) {
        case 1: return launch_bwd_rounds<VT, 1>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
        case 2: return launch_bwd_rounds<VT, 2>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
        case 3: return launch_bwd_rounds<VT, 3>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
        default: return launch_bwd_rounds<VT, 4>(go, value, shapes, start, loc, attn, gv, gl, ga, d, fa, st);
    }
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
    const Dims d = make_dims(N, S, M, D, L, Lq, P, (int)sizeof(T));
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
    if (!strcmp(key, "frame_chunk")) { g_unit = value; return MSDA_OK; }
    if (!strcmp(key, "fwd_warps")) { g_fwd_warps = value; return MSDA_OK; }
    if (!strcmp(key, "bwd_warps")) { g_bwd_warps = value; return MSDA_OK; }
    if (!strcmp(key, "force_generic")) { g_force_generic = value; return MSDA_OK; }
    if (!strcmp(key, "force_linear_walk")) { g_force_linear = value; return MSDA_OK; }
#ifdef MSDA_EXPERIMENTS
    if (!strcmp(key, "bwd_mode")) { g_bwd_mode = value; return MSDA_OK; }
    if (!strcmp(key, "debug_skip_scatter")) { g_skip_scatter = value; return MSDA_OK; }
#else
    if (!strcmp(key, "bwd_mode") || !strcmp(key, "debug_skip_scatter"))
        return fail(MSDA_ERR_UNSUPPORTED, "msda_set_option: measurement-only switch, compiled in with -DMSDA_EXPERIMENTS only (tools/)");
#endif
    if (!strcmp(key, "bwd_deep")) { g_bwd_deep = value; return MSDA_OK; }
    if (!strcmp(key, "bwd_algo")) { g_bwd_algo = value; return MSDA_OK; }
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
    const Dims d = make_dims(N, S, M, D, L, Lq, P, 4);
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
    const Dims d = make_dims(N, S, M, D, L, Lq, P, 4);
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
    const Dims d = make_dims(N, S, M, D, L, Lq, P, 8);
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
    const Dims d = make_dims(N, S, M, D, L, Lq, P, 2);
    return launch_fwd_tiled<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16 *>(value), shapes, start, loc, attn,
                                           reinterpret_cast<__nv_bfloat16 *>(out), d, (cudaStream_t)stream);
}

int msda_backward_bf16(const uint16_t *go, const uint16_t *value, const int64_t *shapes, const int64_t *start,
                       const float *loc, const float *attn, int N, int S, int M, int D, int L, int Lq, int P,
                       float *gv32, uint16_t *gv16, float *gl, float *ga, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (!tiled_ok(D, L, P) || g_force_generic.load())
        return fail(MSDA_ERR_UNSUPPORTED, "msda_backward_bf16: only channels == 32, num_levels <= 16, num_levels*num_point <= 32");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nval = (int64_t)N * S * M * D;
    if (!gv32 && gv16) {
        // direct mode: grad_value accumulated in bf16 by packed reds (row-major kernel only), no fp32 buffer
        if (L > kSortLevels) return fail(MSDA_ERR_UNSUPPORTED, "msda_backward_bf16: direct bf16 accumulation needs num_levels <= 4 (pass grad_value_f32)");
        if (N > 0) {
            const cudaError_t e = cudaMemsetAsync(gv16, 0, sizeof(uint16_t) * (size_t)nval, st);
            if (e != cudaSuccess) return fail_cuda(e, "msda_backward_bf16: memset(grad_value)");
        }
        if ((int64_t)N * Lq == 0) return MSDA_OK;
        if (!go || !value || !shapes || !start || !loc || !attn || !gl || !ga)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: null pointer");
        if (misaligned(value, 8) || misaligned(go, 8) || misaligned(gv16, 8) || misaligned(loc, 8) || misaligned(gl, 8))
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: misaligned pointer");
        const Dims d = make_dims(N, S, M, D, L, Lq, P, 2);
        return launch_bwd_sorted_gv16(reinterpret_cast<const __nv_bfloat16 *>(go), reinterpret_cast<const __nv_bfloat16 *>(value),
                                      shapes, start, loc, attn, gv16, gl, ga, d, nullptr, st);
    }
    if (!gv32 && N > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: null grad_value_f32 and grad_value_bf16");
    if (N > 0) {
        const cudaError_t e = cudaMemsetAsync(gv32, 0, sizeof(float) * (size_t)nval, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_backward_bf16: memset(grad_value)");
    }
    if ((int64_t)N * Lq > 0) {
        if (!go || !value || !shapes || !start || !loc || !attn || !gl || !ga)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: null pointer");
        if (misaligned(value, 8) || misaligned(go, 8) || misaligned(gv32, 16) || misaligned(loc, 8) || misaligned(gl, 8))
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward_bf16: misaligned pointer");
        const Dims d = make_dims(N, S, M, D, L, Lq, P, 2);
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

// ---- fused module path: offsets / logits / reference points in, softmax + location arithmetic inside the kernels ----
namespace {
template <typename VT>
int fused_forward_any(const char *what, const VT *value, const int64_t *shapes, const int64_t *start, const float *offsets,
                      const float *logits, const float *ref, int ref_dim, int N, int S, int M, int D, int L, int Lq, int P,
                      VT *out, float *loc_out, float *attn_out, msda_stream_t stream) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_forward: ref_dim must be 2 or 4");
    if (!tiled_ok(D, L, P)) return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_forward: only channels == 32, num_levels <= 16, num_levels*num_point <= 32 (use the unfused operator)");
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!value || !shapes || !start || !offsets || !logits || !ref || !out) return fail(MSDA_ERR_INVALID_ARGUMENT, what);
    if (misaligned(value, 4 * sizeof(VT)) || misaligned(out, 4 * sizeof(VT)) || misaligned(offsets, 8) ||
        misaligned(ref, ref_dim == 2 ? 8 : 16) || (loc_out && misaligned(loc_out, 8)))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_forward: misaligned pointer");
    const Dims d = make_dims(N, S, M, D, L, Lq, P, (int)sizeof(VT));
    const FusedArgs fa{ref, ref_dim, loc_out, attn_out, nullptr};
    return launch_fwd_tiled<VT>(value, shapes, start, offsets, logits, out, d, (cudaStream_t)stream, &fa);
}

template <typename VT>
int fused_backward_any(const char *what, const VT *go, const VT *value, const int64_t *shapes, const int64_t *start,
                       const float *offsets, const float *logits, const float *ref, int ref_dim, int N, int S, int M, int D,
                       int L, int Lq, int P, float *gv32, float *g_off, float *g_logits, float *g_loc, cudaStream_t st) {
    if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
    if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward: ref_dim must be 2 or 4");
    if (!tiled_ok(D, L, P)) return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_backward: only channels == 32, num_levels <= 16, num_levels*num_point <= 32 (use the unfused operator)");
    if (!gv32 && N > 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward: null grad_value");
    if (N > 0) {
        const cudaError_t e = cudaMemsetAsync(gv32, 0, sizeof(float) * (size_t)N * S * M * D, st);
        if (e != cudaSuccess) return fail_cuda(e, "msda_fused_backward: memset(grad_value)");
    }
    if ((int64_t)N * Lq == 0) return MSDA_OK;
    if (!go || !value || !shapes || !start || !offsets || !logits || !ref || !g_off || !g_logits) return fail(MSDA_ERR_INVALID_ARGUMENT, what);
    if (misaligned(value, 4 * sizeof(VT)) || misaligned(go, 4 * sizeof(VT)) || misaligned(gv32, 16) || misaligned(offsets, 8) ||
        misaligned(g_off, 8) || misaligned(ref, ref_dim == 2 ? 8 : 16) || (g_loc && misaligned(g_loc, 8)))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward: misaligned pointer");
    const Dims d = make_dims(N, S, M, D, L, Lq, P, (int)sizeof(VT));
    const FusedArgs fa{ref, ref_dim, nullptr, nullptr, g_loc};
    return launch_bwd_tiled<VT>(go, value, shapes, start, offsets, logits, gv32, g_off, g_logits, d, st, &fa);
}
}  // namespace

extern "C" {

int msda_fused_forward_f32(const float *value, const int64_t *shapes, const int64_t *start, const float *offsets,
                           const float *logits, const float *ref, int ref_dim, int N, int S, int M, int D, int L, int Lq,
                           int P, float *out, float *loc_out, float *attn_out, msda_stream_t stream) {
    return fused_forward_any<float>("msda_fused_forward_f32: null pointer", value, shapes, start, offsets, logits, ref, ref_dim,
                                    N, S, M, D, L, Lq, P, out, loc_out, attn_out, stream);
}

int msda_fused_backward_f32(const float *go, const float *value, const int64_t *shapes, const int64_t *start,
                            const float *offsets, const float *logits, const float *ref, int ref_dim, int N, int S, int M,
                            int D, int L, int Lq, int P, float *gv, float *g_off, float *g_logits, float *g_loc,
                            msda_stream_t stream) {
    return fused_backward_any<float>("msda_fused_backward_f32: null pointer", go, value, shapes, start, offsets, logits, ref,
                                     ref_dim, N, S, M, D, L, Lq, P, gv, g_off, g_logits, g_loc, (cudaStream_t)stream);
}

int msda_fused_forward_bf16(const uint16_t *value, const int64_t *shapes, const int64_t *start, const float *offsets,
                            const float *logits, const float *ref, int ref_dim, int N, int S, int M, int D, int L, int Lq,
                            int P, uint16_t *out, float *loc_out, float *attn_out, msda_stream_t stream) {
    return fused_forward_any<__nv_bfloat16>("msda_fused_forward_bf16: null pointer", reinterpret_cast<const __nv_bfloat16 *>(value),
                                            shapes, start, offsets, logits, ref, ref_dim, N, S, M, D, L, Lq, P,
                                            reinterpret_cast<__nv_bfloat16 *>(out), loc_out, attn_out, stream);
}

int msda_fused_backward_bf16(const uint16_t *go, const uint16_t *value, const int64_t *shapes, const int64_t *start,
                             const float *offsets, const float *logits, const float *ref, int ref_dim, int N, int S, int M,
                             int D, int L, int Lq, int P, float *gv32, uint16_t *gv16, float *g_off, float *g_logits,
                             float *g_loc, msda_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!gv32 && gv16) {      // direct mode, see msda_backward_bf16
        if (const int rc = check_dims(N, S, M, D, L, Lq, P)) return rc;
        if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward: ref_dim must be 2 or 4");
        if (!tiled_ok(D, L, P) || L > kSortLevels)
            return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_backward_bf16: direct bf16 accumulation needs channels == 32, num_levels <= 4, num_levels*num_point <= 32");
        if (N > 0) {
            const cudaError_t e = cudaMemsetAsync(gv16, 0, sizeof(uint16_t) * (size_t)N * S * M * D, st);
            if (e != cudaSuccess) return fail_cuda(e, "msda_fused_backward_bf16: memset(grad_value)");
        }
        if ((int64_t)N * Lq == 0) return MSDA_OK;
        if (!go || !value || !shapes || !start || !offsets || !logits || !ref || !g_off || !g_logits)
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward_bf16: null pointer");
        if (misaligned(value, 8) || misaligned(go, 8) || misaligned(gv16, 8) || misaligned(offsets, 8) || misaligned(g_off, 8) ||
            misaligned(ref, ref_dim == 2 ? 8 : 16) || (g_loc && misaligned(g_loc, 8)))
            return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward_bf16: misaligned pointer");
        const Dims d = make_dims(N, S, M, D, L, Lq, P, 2);
        const FusedArgs fa{ref, ref_dim, nullptr, nullptr, g_loc};
        return launch_bwd_sorted_gv16(reinterpret_cast<const __nv_bfloat16 *>(go), reinterpret_cast<const __nv_bfloat16 *>(value),
                                      shapes, start, offsets, logits, gv16, g_off, g_logits, d, &fa, st);
    }
    if (const int rc = fused_backward_any<__nv_bfloat16>("msda_fused_backward_bf16: null pointer",
                                                         reinterpret_cast<const __nv_bfloat16 *>(go),
                                                         reinterpret_cast<const __nv_bfloat16 *>(value), shapes, start, offsets,
                                                         logits, ref, ref_dim, N, S, M, D, L, Lq, P, gv32, g_off, g_logits, g_loc, st))
        return rc;
    const int64_t nval = (int64_t)N * S * M * D;
    if (gv16 && nval > 0) {
        if (misaligned(gv16, 8)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward_bf16: misaligned grad_value_bf16");
        const int64_t n4 = nval / 4;
        const int grid = (int)((n4 + 255) / 256 < (int64_t)sm_count() * 16 ? (n4 + 255) / 256 : (int64_t)sm_count() * 16);
        msda_f32_to_bf16<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(gv32), reinterpret_cast<uint2 *>(gv16), n4);
        return after_launch("msda_f32_to_bf16");
    }
    return MSDA_OK;
}

}  // extern "C"

// ---- encoder layer epilogue: bias + residual + LayerNorm, Linear bias gradients ----
namespace {
template <int VEC>
int launch_ln_fwd(const float *x, const float *bias, const float *res, const float *gamma, const float *beta, float eps,
                  int64_t rows, float *z, float *y, float *mean, float *rstd, const DropoutArgs *da, cudaStream_t st) {
    const int64_t want = (rows + 7) / 8;
    const int grid = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    if (da) epilogue_ln_fwd<VEC, true><<<grid, 256, 0, st>>>(x, bias, res, gamma, beta, eps, rows, z, y, mean, rstd, *da);
    else epilogue_ln_fwd<VEC, false><<<grid, 256, 0, st>>>(x, bias, res, gamma, beta, eps, rows, z, y, mean, rstd, DropoutArgs{});
    return after_launch("epilogue_ln_fwd");
}
template <int VEC>
int launch_ln_bwd(const float *dy, const float *z, const float *mean, const float *rstd, const float *gamma, int64_t rows,
                  float *dz, float *dgamma, float *dbeta, float *dbias, float *dx, const DropoutArgs *da, cudaStream_t st) {
    const int64_t want = (rows + 7) / 8;
    const int grid = (int)(want < (int64_t)sm_count() * 2 ? want : (int64_t)sm_count() * 2);
    if (da) epilogue_ln_bwd<VEC, true><<<grid, 256, 0, st>>>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, dx, *da);
    else epilogue_ln_bwd<VEC, false><<<grid, 256, 0, st>>>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, nullptr, DropoutArgs{});
    return after_launch("epilogue_ln_bwd");
}
// p in [0, 1): keep <=> 32-bit word >= floor(p * 2^32)
int make_dropout(const void *rng, uint32_t salt, float p, DropoutArgs *out, const char *who) {
    if (!rng || misaligned(rng, 8)) return fail(MSDA_ERR_INVALID_ARGUMENT, "dropout: rng must point at two 8-byte aligned 64-bit words");
    if (!(p >= 0.f && p < 1.f)) return fail(MSDA_ERR_INVALID_ARGUMENT, "dropout: p must be in [0, 1)");
    (void)who;
    out->rng = static_cast<const uint64_t *>(rng);
    out->salt = salt;
    out->thresh = (uint32_t)((double)p * 4294967296.0);
    out->scale = 1.f / (1.f - p);
    return MSDA_OK;
}
int zero_fill(float *p, int64_t n, cudaStream_t st, const char *what) {
    if (!p || n <= 0) return MSDA_OK;
    const cudaError_t e = cudaMemsetAsync(p, 0, sizeof(float) * (size_t)n, st);
    return e == cudaSuccess ? MSDA_OK : fail_cuda(e, what);
}
int column_sum_any(bool relu, const float *x, const float *h, int64_t rows, int C, float *dpre, float *out, cudaStream_t st,
                   float scale = 1.f) {
    if (rows < 0 || C <= 0 || (C & 3)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_column_sum: channels must be a positive multiple of 4");
    if (!out) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_column_sum: null output");
    if (const int rc = zero_fill(out, C, st, "msda_column_sum: memset")) return rc;
    if (rows == 0) return MSDA_OK;
    if (!x || (relu && (!h || !dpre))) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_column_sum: null pointer");
    if (misaligned(x, 16) || misaligned(out, 16) || (relu && (misaligned(h, 16) || misaligned(dpre, 16))))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_column_sum: pointers must be 16-byte aligned");
    // few CTAs: each one ends with C/4 same-address reds, which the L2 serialises
    const int64_t want = (rows + 31) / 32, cap = (int64_t)sm_count() * (relu ? 4 : 2);
    const int grid = (int)(want < cap ? want : cap);
    const size_t smem = 256 * sizeof(float4);
    if (relu) column_sum_kernel<true><<<grid, 256, smem, st>>>(x, h, rows, C, dpre, out, scale);
    else column_sum_kernel<false><<<grid, 256, smem, st>>>(x, nullptr, rows, C, nullptr, out, 1.f);
    return after_launch("column_sum_kernel");
}
}  // namespace

extern "C" {

static int ln_forward_any(const float *x, const float *bias, const float *residual, const float *gamma, const float *beta, float eps,
                          int64_t rows, int channels, const DropoutArgs *da, float *z, float *y, float *mean, float *rstd,
                          msda_stream_t stream) {
    if (rows < 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_forward_f32: negative row count");
    if (rows == 0) return MSDA_OK;
    if (!x || !residual || !gamma || !beta || !z || !y || !mean || !rstd)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_forward_f32: null pointer");
    if (misaligned(x, 16) || misaligned(residual, 16) || misaligned(gamma, 16) || misaligned(beta, 16) || misaligned(z, 16) ||
        misaligned(y, 16) || (bias && misaligned(bias, 16)))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_forward_f32: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (channels) {
        case 128: return launch_ln_fwd<1>(x, bias, residual, gamma, beta, eps, rows, z, y, mean, rstd, da, st);
        case 256: return launch_ln_fwd<2>(x, bias, residual, gamma, beta, eps, rows, z, y, mean, rstd, da, st);
        case 512: return launch_ln_fwd<4>(x, bias, residual, gamma, beta, eps, rows, z, y, mean, rstd, da, st);
        case 1024: return launch_ln_fwd<8>(x, bias, residual, gamma, beta, eps, rows, z, y, mean, rstd, da, st);
        default: return fail(MSDA_ERR_UNSUPPORTED, "msda_epilogue_ln_forward_f32: channels must be 128, 256, 512 or 1024");
    }
}

static int ln_backward_any(const float *dy, const float *z, const float *mean, const float *rstd, const float *gamma,
                           int64_t rows, int channels, const DropoutArgs *da, float *dz, float *dx, float *dgamma, float *dbeta,
                           float *dbias, msda_stream_t stream) {
    if (rows < 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_backward_f32: negative row count");
    if (channels != 128 && channels != 256 && channels != 512 && channels != 1024)
        return fail(MSDA_ERR_UNSUPPORTED, "msda_epilogue_ln_backward_f32: channels must be 128, 256, 512 or 1024");
    if (!dgamma || !dbeta) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_backward_f32: null grad_gamma / grad_beta");
    cudaStream_t st = (cudaStream_t)stream;
    if (const int rc = zero_fill(dgamma, channels, st, "msda_epilogue_ln_backward_f32: memset")) return rc;
    if (const int rc = zero_fill(dbeta, channels, st, "msda_epilogue_ln_backward_f32: memset")) return rc;
    if (const int rc = zero_fill(dbias, channels, st, "msda_epilogue_ln_backward_f32: memset")) return rc;
    if (rows == 0) return MSDA_OK;
    if (!dy || !z || !mean || !rstd || !gamma || !dz || (da && !dx))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_backward_f32: null pointer");
    if (misaligned(dy, 16) || misaligned(z, 16) || misaligned(gamma, 16) || misaligned(dz, 16) || misaligned(dgamma, 16) ||
        misaligned(dbeta, 16) || (dbias && misaligned(dbias, 16)) || (da && misaligned(dx, 16)))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_epilogue_ln_backward_f32: pointers must be 16-byte aligned");
    switch (channels) {
        case 128: return launch_ln_bwd<1>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, dx, da, st);
        case 256: return launch_ln_bwd<2>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, dx, da, st);
        case 512: return launch_ln_bwd<4>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, dx, da, st);
        default: return launch_ln_bwd<8>(dy, z, mean, rstd, gamma, rows, dz, dgamma, dbeta, dbias, dx, da, st);
    }
}

int msda_epilogue_ln_forward_f32(const float *x, const float *bias, const float *residual, const float *gamma,
                                 const float *beta, float eps, int64_t rows, int channels, float *z, float *y, float *mean,
                                 float *rstd, msda_stream_t stream) {
    return ln_forward_any(x, bias, residual, gamma, beta, eps, rows, channels, nullptr, z, y, mean, rstd, stream);
}

int msda_epilogue_ln_backward_f32(const float *dy, const float *z, const float *mean, const float *rstd, const float *gamma,
                                  int64_t rows, int channels, float *dz, float *dgamma, float *dbeta, float *dbias,
                                  msda_stream_t stream) {
    return ln_backward_any(dy, z, mean, rstd, gamma, rows, channels, nullptr, dz, nullptr, dgamma, dbeta, dbias, stream);
}

int msda_epilogue_ln_dropout_forward_f32(const float *x, const float *bias, const float *residual, const float *gamma,
                                         const float *beta, float eps, int64_t rows, int channels, const void *rng,
                                         uint32_t salt, float p, float *z, float *y, float *mean, float *rstd,
                                         msda_stream_t stream) {
    DropoutArgs da;
    if (const int rc = make_dropout(rng, salt, p, &da, "msda_epilogue_ln_dropout_forward_f32")) return rc;
    return ln_forward_any(x, bias, residual, gamma, beta, eps, rows, channels, &da, z, y, mean, rstd, stream);
}

int msda_epilogue_ln_dropout_backward_f32(const float *dy, const float *z, const float *mean, const float *rstd,
                                          const float *gamma, int64_t rows, int channels, const void *rng, uint32_t salt,
                                          float p, float *dz, float *dx, float *dgamma, float *dbeta, float *dbias,
                                          msda_stream_t stream) {
    DropoutArgs da;
    if (const int rc = make_dropout(rng, salt, p, &da, "msda_epilogue_ln_dropout_backward_f32")) return rc;
    return ln_backward_any(dy, z, mean, rstd, gamma, rows, channels, &da, dz, dx, dgamma, dbeta, dbias, stream);
}

static int dropout_stream_grid(int64_t n4) {
    const int64_t want = (n4 + 255) / 256, cap = (int64_t)sm_count() * 8;
    return (int)(want < cap ? want : cap);
}

int msda_dropout_inplace_f32(float *h, int64_t n, const void *rng, uint32_t salt, float p, msda_stream_t stream) {
    if (n < 0 || (n & 3)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_inplace_f32: element count must be a non-negative multiple of 4");
    DropoutArgs da;
    if (const int rc = make_dropout(rng, salt, p, &da, "msda_dropout_inplace_f32")) return rc;
    if (n == 0) return MSDA_OK;
    if (!h || misaligned(h, 16)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_inplace_f32: h must be a 16-byte aligned pointer");
    dropout_inplace_kernel<<<dropout_stream_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(h), n / 4, da);
    return after_launch("dropout_inplace_kernel");
}

int msda_dropout_mask_u8(const void *rng, uint32_t salt, float p, int64_t n, uint8_t *keep, msda_stream_t stream) {
    if (n < 0 || (n & 3)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_mask_u8: element count must be a non-negative multiple of 4");
    DropoutArgs da;
    if (const int rc = make_dropout(rng, salt, p, &da, "msda_dropout_mask_u8")) return rc;
    if (n == 0) return MSDA_OK;
    if (!keep || misaligned(keep, 4)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_mask_u8: keep must be a 4-byte aligned pointer");
    dropout_mask_kernel<<<dropout_stream_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uchar4 *>(keep), n / 4, da);
    return after_launch("dropout_mask_kernel");
}

int msda_column_sum_f32(const float *x, int64_t rows, int channels, float *out, msda_stream_t stream) {
    return column_sum_any(false, x, nullptr, rows, channels, nullptr, out, (cudaStream_t)stream);
}

int msda_relu_backward_column_sum_f32(const float *dh, const float *h, int64_t rows, int channels, float *dpre, float *dbias,
                                      msda_stream_t stream) {
    return column_sum_any(true, dh, h, rows, channels, dpre, dbias, (cudaStream_t)stream);
}

int msda_relu_dropout_backward_column_sum_f32(const float *dh, const float *h_dropped, float p, int64_t rows, int channels,
                                              float *dpre, float *dbias, msda_stream_t stream) {
    if (!(p >= 0.f && p < 1.f)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_relu_dropout_backward_column_sum_f32: p must be in [0, 1)");
    return column_sum_any(true, dh, h_dropped, rows, channels, dpre, dbias, (cudaStream_t)stream, 1.f / (1.f - p));
}

// ---- decoder-side consumers ----
int msda_decoder_select_samples_f32(const float *sampling_loc, const float *attn_weight, const float *valid_ratios, int batch,
                                    int num_query, int num_heads, int num_levels, int num_point, int top,
                                    float *samples_keep, float *top_weights, int64_t *top_idx, msda_stream_t stream) {
    if (batch < 0 || num_query < 0 || num_heads <= 0 || num_levels <= 0 || num_point <= 0 || top <= 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_select_samples_f32: non-positive dimension");
    const int64_t K64 = (int64_t)num_heads * num_levels * num_point;
    if (top > 32 || K64 > 256 || top > K64)
        return fail(MSDA_ERR_UNSUPPORTED, "msda_decoder_select_samples_f32: needs top <= 32, top <= heads*levels*points <= 256");
    const int64_t rows = (int64_t)batch * num_query;
    if (rows == 0) return MSDA_OK;
    if (!sampling_loc || !attn_weight || !valid_ratios || !samples_keep)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_select_samples_f32: null pointer");
    if (misaligned(sampling_loc, 8) || misaligned(valid_ratios, 8) || misaligned(samples_keep, 8) || misaligned(attn_weight, 4) ||
        (top_idx && misaligned(top_idx, 8)))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_select_samples_f32: misaligned pointer");
    const int K = (int)K64;
    const int grid = (int)((rows + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
#define MSDA_SELECT(ITEMS) decoder_select_samples_kernel<ITEMS><<<grid, 128, 0, st>>>(sampling_loc, attn_weight, valid_ratios, rows, \
        num_query, K, num_levels, num_point, top, samples_keep, top_weights, top_idx)
    if (K <= 32) MSDA_SELECT(1);
    else if (K <= 64) MSDA_SELECT(2);
    else if (K <= 128) MSDA_SELECT(4);
    else MSDA_SELECT(8);
#undef MSDA_SELECT
    return after_launch("decoder_select_samples_kernel");
}

int msda_decoder_reference_points_f32(const float *reference_points, const float *valid_ratios, int batch, int num_query,
                                      int num_levels, int ref_dim, float *reference_points_input, msda_stream_t stream) {
    if (batch < 0 || num_query < 0 || num_levels <= 0)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_reference_points_f32: non-positive dimension");
    if (ref_dim != 2 && ref_dim != 4) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_reference_points_f32: ref_dim must be 2 or 4");
    const int64_t total = (int64_t)batch * num_query * num_levels;
    if (total == 0) return MSDA_OK;
    if (!reference_points || !valid_ratios || !reference_points_input)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_reference_points_f32: null pointer");
    const size_t al = ref_dim == 2 ? 8 : 16;
    if (misaligned(reference_points, al) || misaligned(reference_points_input, al) || misaligned(valid_ratios, 8))
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_decoder_reference_points_f32: misaligned pointer");
    const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
    decoder_reference_points_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
        reference_points, valid_ratios, total, num_query, num_levels, ref_dim, reference_points_input);
    return after_launch("decoder_reference_points_kernel");
}

}  // extern "C"

namespace {
template <typename K>
int flatten_launch(const FlattenArgs &a, K kernel, size_t smem, cudaStream_t st, const char *what) {
    // > 48 KB of dynamic shared memory needs the opt-in; set on every call (cheap, and correct on every device)
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "flatten: cudaFuncSetAttribute");
    }
    int per_sm = 0;                  // resident CTAs per SM of this kernel: one full wave of persistent CTAs
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm <= 0) per_sm = 1;
    const int64_t total = (int64_t)a.N * a.tiles_per_frame, cap = (int64_t)sm_count() * per_sm;
    kernel<<<(int)(total < cap ? total : cap), 256, smem, st>>>(a);
    return after_launch(what);
}
}  // namespace

extern "C" {

// ---- caller-side flattening ----
static int flatten_setup(FlattenArgs &a, const char *who, int num_levels, const int *heights, const int *widths, int batch,
                         int channels, int spatial_size, bool *vec4) {
    if (num_levels <= 0 || num_levels > kFlatMaxLevels) return fail(MSDA_ERR_UNSUPPORTED, "flatten: 1 <= num_levels <= 8");
    if (batch < 0 || channels <= 0 || !heights || !widths) return fail(MSDA_ERR_INVALID_ARGUMENT, "flatten: bad dimension / null shape array");
    (void)who;
    memset(&a, 0, sizeof(a));
    a.L = num_levels; a.N = batch; a.C = channels;
    // 16-byte kernels when every level's pixel count and the channel count are multiples of 4 (pointers: checked by the caller)
    bool v4 = channels % 4 == 0;
    for (int l = 0; l < num_levels; ++l) v4 = v4 && ((int64_t)heights[l] * widths[l]) % 4 == 0;
    *vec4 = v4;
    const int tile_hw = v4 ? kV4Tile : kFlatTileHW, tile_c = v4 ? kV4Tile : kFlatTileC;
    a.tiles_c = (channels + tile_c - 1) / tile_c;
    int64_t rows = 0, tiles = 0;
    for (int l = 0; l < num_levels; ++l) {
        if (heights[l] <= 0 || widths[l] <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "flatten: non-positive level shape");
        const int64_t hw = (int64_t)heights[l] * widths[l];
        if (rows + hw > INT32_MAX / 2) return fail(MSDA_ERR_UNSUPPORTED, "flatten: more than 2^30 pixels per frame");
        a.lv[l].hw = (int)hw;
        a.lv[l].start = (int)rows;
        a.lv[l].tiles_hw = (int)((hw + tile_hw - 1) / tile_hw);
        a.lv[l].tile_begin = (int)tiles;
        tiles += (int64_t)a.lv[l].tiles_hw * a.tiles_c;
        rows += hw;
    }
    if (tiles > INT32_MAX) return fail(MSDA_ERR_UNSUPPORTED, "flatten: too many tiles per frame");
    a.tiles_per_frame = (int)tiles;
    a.S = spatial_size > 0 ? spatial_size : (int)rows;
    if (a.S < rows) return fail(MSDA_ERR_INVALID_ARGUMENT, "flatten: spatial_size smaller than the levels' pixels");
    return MSDA_OK;
}
int msda_flatten_levels_f32(int num_levels, const float *const *src_levels, const float *const *pos_levels,
                            const float *level_embed, const int *heights, const int *widths, int batch, int channels,
                            float *src_flatten, float *pos_flatten, msda_stream_t stream) {
    FlattenArgs a;
    bool v4 = false;
    if (const int rc = flatten_setup(a, "msda_flatten_levels_f32", num_levels, heights, widths, batch, channels, 0, &v4)) return rc;
    if (batch == 0) return MSDA_OK;
    if (!src_levels || !src_flatten || (pos_levels && !pos_flatten)) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_flatten_levels_f32: null pointer");
    bool aligned = !misaligned(src_flatten, 16) && !(pos_flatten && misaligned(pos_flatten, 16)) && !(level_embed && misaligned(level_embed, 16));
    for (int l = 0; l < num_levels; ++l) {
        if (!src_levels[l] || (pos_levels && !pos_levels[l])) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_flatten_levels_f32: null level pointer");
        aligned = aligned && !misaligned(src_levels[l], 16) && !(pos_levels && misaligned(pos_levels[l], 16));
    }
    if (v4 && !aligned) {
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_flatten_levels_f32: pointers must be 16-byte aligned when every H*W and channels are multiples of 4");
    }
    for (int l = 0; l < num_levels; ++l) {
        a.lv[l].src = src_levels[l];
        a.lv[l].pos = pos_levels ? pos_levels[l] : nullptr;
    }
    a.level_embed = level_embed;
    a.src_flat = src_flatten;
    a.pos_flat = pos_flatten;
    cudaStream_t st = (cudaStream_t)stream;
    if (v4) {
        if (pos_levels) return flatten_launch(a, flatten_levels_v4_kernel<true>, 0, st, "flatten_levels_v4_kernel");
        return flatten_launch(a, flatten_levels_v4_kernel<false>, 0, st, "flatten_levels_v4_kernel");
    }
    const size_t tile = sizeof(float) * kFlatTileC * (kFlatTileHW + 1) * kFlatStages;
    if (pos_levels) return flatten_launch(a, flatten_levels_kernel<true>, 2 * tile, st, "flatten_levels_kernel");
    return flatten_launch(a, flatten_levels_kernel<false>, tile, st, "flatten_levels_kernel");
}

int msda_unflatten_levels_f32(int num_levels, const float *flat, const int *heights, const int *widths, int batch, int channels,
                              int spatial_size, float *const *maps, msda_stream_t stream) {
    FlattenArgs a;
    bool v4 = false;
    if (const int rc = flatten_setup(a, "msda_unflatten_levels_f32", num_levels, heights, widths, batch, channels, spatial_size, &v4)) return rc;
    if (batch == 0) return MSDA_OK;
    if (!flat || !maps) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_unflatten_levels_f32: null pointer");
    bool aligned = !misaligned(flat, 16);
    for (int l = 0; l < num_levels; ++l) {
        if (!maps[l]) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_unflatten_levels_f32: null level pointer");
        aligned = aligned && !misaligned(maps[l], 16);
        a.lv[l].map_out = maps[l];
    }
    if (v4 && !aligned)
        return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_unflatten_levels_f32: pointers must be 16-byte aligned when every H*W and channels are multiples of 4");
    a.flat_in = flat;
    if (v4) return flatten_launch(a, unflatten_levels_v4_kernel, 0, (cudaStream_t)stream, "unflatten_levels_v4_kernel");
    return flatten_launch(a, unflatten_levels_kernel, sizeof(float) * kFlatTileHW * (kFlatTileC + 1) * kFlatStages,
                          (cudaStream_t)stream, "unflatten_levels_kernel");
}

}  // extern "C"
