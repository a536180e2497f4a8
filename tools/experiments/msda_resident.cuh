// msda_resident.cuh -- the "resident" kernels: the coarse pyramid levels of one (frame, head) slice live in
// shared memory for as long as a CTA works on that slice.  Included by msda_sm100.cu inside its anonymous
// namespace (it uses the helpers defined there: LevelTable, Row, RowIO, point geometry, select_query).
//
// Why.  The tiled kernels are bound by what an SM can move between L1/L2 and its registers, not by HBM
// (DESIGN.md section 3): a four-row LDG.128 costs 4 cycles when every row hits L1 and 8 as soon as one misses,
// and backward every scattered row is a 128-byte atomic that the L2 serialises (6.4 TB/s chip-wide).  The
// pyramid makes most of that traffic go to very few rows: of the L*P*4 = 64 rows a (query, head) touches, 48
// lie in levels 1..3, which for a 360x640 input are 1220 rows = 152 KB per (frame, head) -- they fit in the
// 227 KB of shared memory of one SM.  So:
//   * gather (forward, and the backward's dot products): rows of the resident levels are read with LDS.128 --
//     1.0 cycle per row, never a miss; only the finest level(s) still go through L1/L2;
//   * scatter (grad_value): rows of the resident levels are accumulated in shared memory with plain
//     load-add-store by warps that OWN a (level, channel-slice) -- no atomics at all -- and flushed to global
//     memory once per slice; only the finest level(s) still send reds to L2.
// Which levels are resident is decided on the device from the int64 shape tensors (the host never reads
// them, like the reference: ms_deform_attn_cuda.cu:67-68): the longest suffix of levels that is contiguous in
// `value` and fits the capacity.  With nothing resident the kernels degrade to plain global gathers / reds.
//
// Work decomposition: one persistent CTA per SM.  Frames are taken in chunks that keep their slices in L2
// together; inside a chunk the (slice, tile) space is cut into one contiguous range per CTA, so a CTA stages
// at most two slices per chunk.

struct ResTable {
    int l0;      // first resident level (== L: none)
    int s0;      // first resident row of the slice
    int rows;    // resident rows: [s0, S)
};

__device__ __forceinline__ void resident_plan(ResTable &rt, const LevelTable &lt, int L, int S, int cap_rows) {
    int l0 = L, s0 = S, end = S;
    for (int l = L - 1; l >= 0; --l) {
        if (lt.H[l] <= 0 || lt.W[l] <= 0 || lt.start[l] < 0) break;
        if (lt.start[l] + lt.H[l] * lt.W[l] != end) break;        // not contiguous with the levels above it
        if (S - lt.start[l] > cap_rows) break;
        l0 = l; s0 = lt.start[l]; end = lt.start[l];
    }
    rt.l0 = l0; rt.s0 = s0; rt.rows = S - s0;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Copy rows [s0, S) of head m of frame n into shared memory, row-major, 32 channels per row.  Ends with a
// CTA barrier; the caller puts one in front (the previous slice may still be in use).
template <typename VT>
__device__ __forceinline__ void stage_resident_rows(unsigned char *s_rows, const VT *value, int n, int m, const Dims &d,
                                                    const ResTable &rt) {
    constexpr int kRowB = 32 * (int)sizeof(VT), kChunks = kRowB / 16;
    const unsigned char *src = reinterpret_cast<const unsigned char *>(value) +
                               (((int64_t)n * d.S + rt.s0) * d.M + m) * (int64_t)kRowB;
    const int64_t stride = (int64_t)d.M * kRowB;
    const uint32_t dst = smem_u32(s_rows);
    const int total = rt.rows * kChunks;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int r = i / kChunks, c = i - r * kChunks;
        cp_async16(dst + (uint32_t)i * 16u, src + r * stride + c * 16);
    }
    cp_async_wait_all();
    __syncthreads();
}

// The four corner pixels of a sample point as ROW numbers of the slice (level_start + y*W + x), clamped into
// the level with the weight zeroed -- same arithmetic as point_geometry() in msda_sm100.cu.
struct PointPx {
    int s00, s01, s10, s11;
    float w00, w01, w10, w11;
    float lx, ly;
    int valid;
};
__device__ __forceinline__ PointPx point_pixels(float loc_x, float loc_y, int H, int W, int level_start) {
    PointPx g;
    const float fw = (float)W, fh = (float)H;
    const float x = fmaf(loc_x, fw, -0.5f);
    const float y = fmaf(loc_y, fh, -0.5f);
    const bool in_range = (y > -1.f) && (x > -1.f) && (y < fh) && (x < fw);
    const float xf = floorf(x), yf = floorf(y);
    const int x0 = (int)xf, y0 = (int)yf;
    g.lx = in_range ? x - xf : 0.f;
    g.ly = in_range ? y - yf : 0.f;
    const float hx = 1.f - g.lx, hy = 1.f - g.ly;
    const int xc0 = min(max(x0, 0), W - 1), yc0 = min(max(y0, 0), H - 1);
    const int xc1 = min(max(x0, -1) + 1, W - 1), yc1 = min(max(y0, -1) + 1, H - 1);
    const bool xa = in_range && x0 >= 0, xb = in_range && x0 < W - 1;
    const bool ya = y0 >= 0, yb = y0 < H - 1;
    const int r0 = level_start + yc0 * W, r1 = level_start + yc1 * W;
    g.s00 = r0 + xc0; g.s01 = r0 + xc1; g.s10 = r1 + xc0; g.s11 = r1 + xc1;
    g.w00 = (xa && ya) ? hy * hx : 0.f;
    g.w01 = (xb && ya) ? hy * g.lx : 0.f;
    g.w10 = (xa && yb) ? g.ly * hx : 0.f;
    g.w11 = (xb && yb) ? g.ly * g.lx : 0.f;
    g.valid = (int)(xa && ya) | ((int)(xb && ya) << 1) | ((int)(xa && yb) << 2) | ((int)(xb && yb) << 3);
    return g;
}

// Contiguous cut of a chunk's (slice, tile) space: CTA `rank` of `workers` takes [lo, hi).
struct ChunkRange {
    int64_t lo, hi;
    __device__ __forceinline__ ChunkRange(int frames, int M, int tiles, int rank, int workers) {
        const int64_t work = (int64_t)frames * M * tiles;
        lo = work * rank / workers;
        hi = work * (rank + 1) / workers;
    }
};

// ------------------------------------------------------------------------------------------------
// Resident forward.  Same warp layout as msda_fwd_tiled: a warp = 4 x-adjacent queries of one head, 8 lanes
// x 16 B per row; per point 2 broadcast LDS.128 (weights, byte offsets) + 4 row loads (LDS.128 from the
// resident levels, LDG.128 otherwise) + 8 FFMA2.
// ------------------------------------------------------------------------------------------------
template <int WARPS> struct ResGatherSmem {
    LevelTable lt;
    ResTable rt;
    float4 w[WARPS][4][9];   // a*w00, a*w01, a*w10, a*w11 of the 8 points of one round (+1: the 4 groups read 4 banks)
    uint4 o[WARPS][4][9];    // BYTE offsets of the four corner rows: into the resident copy or the global slice
};
template <typename SM> struct SmemHeader { static constexpr size_t value = (sizeof(SM) + 127) & ~(size_t)127; };

template <typename VT, int ROUNDS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1)
msda_fwd_resident(const VT *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ start,
                  const float *__restrict__ loc, const float *__restrict__ attn, VT *__restrict__ out, Dims d,
                  int cap_rows) {
    using IO = RowIO<VT>;
    using Vec = typename IO::Vec;
    using SM = ResGatherSmem<WARPS>;
    constexpr int kTaskQueries = Tile<WARPS>::kQueries;
    constexpr uint32_t kRowB = 32 * sizeof(VT);
    constexpr int kBlock = WARPS >= 32 ? 2 : 4;      // points per straight-line block: row loads in flight vs registers
    SM &sm = *reinterpret_cast<SM *>(msda_smem);
    unsigned char *s_rows = msda_smem + SmemHeader<SM>::value;
    LevelTable &lt = sm.lt;
    load_level_table<WARPS>(lt, shapes, start, d.L, d.Lq);
    if (threadIdx.x == 0) resident_plan(sm.rt, lt, d.L, d.S, cap_rows);
    __syncthreads();
    const ResTable rt = sm.rt;
    const bool tiled = d.tiled && lt.dense;
    const int nglob = rt.l0 * d.P;          // points [0, nglob) gather from global memory, the rest from shared

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int pts = d.L * d.P;
    uint32_t lv = 0;     // level of this lane's point in each round, one byte per round
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) lv |= (uint32_t)min((8 * r + cl) / d.P, d.L - 1) << (8 * r);
    const unsigned char *s_lane = s_rows + cl * sizeof(Vec);
    const float4 *pw_ = sm.w[warp][grp];
    const uint4 *po_ = sm.o[warp][grp];

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kTaskQueries - 1) / kTaskQueries;
    for (int f0 = 0; f0 < d.N; f0 += d.fchunk) {
        const ChunkRange range(min(d.fchunk, d.N - f0), d.M, tiles, blockIdx.x, gridDim.x);
        int cur = -1;
        for (int64_t wk = range.lo; wk < range.hi; ++wk) {
            const int slice = (int)(wk / tiles), tile = (int)(wk - (int64_t)slice * tiles);
            const int n = f0 + slice / d.M, m = slice - (slice / d.M) * d.M;
            if (slice != cur) {
                cur = slice;
                __syncthreads();
                stage_resident_rows<VT>(s_rows, value, n, m, d, rt);
            }
            const int q = select_query<WARPS>(tiled, d, lt, tile, warp, grp);
            if (!__any_sync(0xffffffffu, q >= 0)) continue;
            const int64_t row = ((int64_t)n * d.Lq + max(q, 0)) * d.M + m;           // (n, q, m)
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
            const unsigned char *vb = reinterpret_cast<const unsigned char *>(value) +
                                      (((int64_t)n * d.S * d.M + m) * (int64_t)kRowB + cl * sizeof(Vec));
            Row acc{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            auto step_global = [&](int it) {
                const float4 pw = pw_[it];
                const uint4 po = po_[it];
                const Row v00 = IO::load(reinterpret_cast<const Vec *>(row_at(vb, po.x)));
                const Row v01 = IO::load(reinterpret_cast<const Vec *>(row_at(vb, po.y)));
                const Row v10 = IO::load(reinterpret_cast<const Vec *>(row_at(vb, po.z)));
                const Row v11 = IO::load(reinterpret_cast<const Vec *>(row_at(vb, po.w)));
                fma_row(pw.x, v00, acc);
                fma_row(pw.y, v01, acc);
                fma_row(pw.z, v10, acc);
                fma_row(pw.w, v11, acc);
            };
            auto step_shared = [&](int it) {
                const float4 pw = pw_[it];
                const uint4 po = po_[it];
                const Row v00 = IO::cvt(*reinterpret_cast<const Vec *>(s_lane + po.x));
                const Row v01 = IO::cvt(*reinterpret_cast<const Vec *>(s_lane + po.y));
                const Row v10 = IO::cvt(*reinterpret_cast<const Vec *>(s_lane + po.z));
                const Row v11 = IO::cvt(*reinterpret_cast<const Vec *>(s_lane + po.w));
                fma_row(pw.x, v00, acc);
                fma_row(pw.y, v01, acc);
                fma_row(pw.z, v10, acc);
                fma_row(pw.w, v11, acc);
            };
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                {
                    const int l = (lv >> (8 * r)) & 0xff;
                    const PointPx g = point_pixels(xy[r].x, xy[r].y, lt.H[l], lt.W[l], lt.start[l]);
                    const float aa = g.valid ? a[r] : 0.f;      // an out-of-range point ignores its weight (cuh:288)
                    const bool res = l >= rt.l0;
                    const int sub = res ? rt.s0 : 0;
                    const uint32_t mul = res ? kRowB : kRowB * (uint32_t)d.M;
                    sm.w[warp][grp][cl] = make_float4(aa * g.w00, aa * g.w01, aa * g.w10, aa * g.w11);
                    sm.o[warp][grp][cl] = make_uint4((uint32_t)(g.s00 - sub) * mul, (uint32_t)(g.s01 - sub) * mul,
                                                     (uint32_t)(g.s10 - sub) * mul, (uint32_t)(g.s11 - sub) * mul);
                }
                __syncwarp();
                // points [0, nglob): rows from global memory (L1/L2); the rest: rows from shared memory.  Two loops of
                // straight-line blocks of kBlock points, no per-point branch.
                const int g_end = min(max(nglob - 8 * r, 0), 8);
                int it = 0;
                for (; it + kBlock <= g_end; it += kBlock) {
#pragma unroll
                    for (int k = 0; k < kBlock; ++k) step_global(it + k);
                }
                for (; it < g_end; ++it) step_global(it);
                for (; it + kBlock <= 8; it += kBlock) {
#pragma unroll
                    for (int k = 0; k < kBlock; ++k) step_shared(it + k);
                }
                for (; it < 8; ++it) step_shared(it);
                __syncwarp();
            }
            if (q >= 0) IO::store(reinterpret_cast<Vec *>(out) + (row * 8 + cl), acc);
        }
    }
}
