// msda_bwd_sorted.cuh -- the aggregating backward (D = 32): grad_value contributions are summed per DISTINCT row
// inside a CTA tile before they are sent to L2 as vector reds.  Included by msda_sm100.cu.
//
// Implements ms_deform_im2col_cuda.cuh:301-403 (col2im) + :87-159 (bilinear gradients) of the reference, replacing its
// four scalar atomics per thread and point (cuh:125-152) by one `red.global.add.v4.f32` per lane and RUN of up to four
// items of the same row.
//
// Why.  Per (query, head) the backward touches 64 rows of `value` twice: it reads the row (for d/d location, d/d
// attention) and adds a multiple of the query's grad_out row into the same row of grad_value.  In the encoder the 32
// queries of an 8 x 4 tile of one head sample the same few hundred rows over and over: 77 % of the tile's valid
// (query, point, corner) items hit a row that another item of the tile also hits (tools/experiments/duplicate_rows.py).
// The query-major kernel (msda_bwd_tiled) pays one L1 row load and one L2 red per ITEM -- 39.9 M red sectors for a 24.7 MB
// tensor on the A2D shape, 80 % of the chip's L2 atomic throughput.  This kernel turns the tile ROW-major:
//
//   stage   every thread owns (query, point) pairs exactly as in msda_bwd_tiled; it derives the four corner items
//           {row, query, a*w_corner} of each;
//   sort    the tile's <= 2048 items are counting-sorted by row in shared memory (integer shared-memory atomics are
//           native; float ones are a CAS loop).  The histogram covers a WINDOW of 24 x 16 cells per pyramid level,
//           centred on the mean sampled pixel of the tile at that level.  A cell with c items becomes c/4 full RUNS of
//           four item slots, one more run padded with a null item (weight 0) if three items remain, and otherwise
//           c%4 LOOSE items; items outside their window are loose too, so any distribution of sampling locations is
//           handled, just without merging;
//   walk    each 8-lane group of a warp takes a run: it loads the row of `value` ONCE, then for each of the four items
//           reads the query's grad_out row from shared memory (one LDS.128: the same data-path cost as the L1 load it
//           replaces), accumulates  acc += (a*w) * grad_out  in registers and forms the corner dot product
//           p = <grad_out, value_row>  (8-lane butterfly over the four items); then ONE red.v4 per lane.  Fixed trip
//           count, no predicates, no divergence.  Loose items take the same path one at a time.  The dot products go
//           back to the owners through shared memory;
//   finish  the owner of a (query, point) combines its four p's exactly like msda_bwd_tiled (cuh:123-158 regrouped).
//
// Shared-memory ordering is by the CTA barriers S1..S6 alone (no warp-level assumptions).  Per array, writer -> readers:
//   go     stage (before S1)            -> walk (S5..S6); next writer: next tile's stage, after S6
//   stat   stage atomics (before S1)    -> histogram (S1..S2); reset in the scan (after S2), next atomics after S6
//   hist   histogram atomics (S1..S2)   -> scan (S2..S4; it writes cell_pos) -> place (S4..S5); cleared in the walk (after S5),
//                                          next atomics after the next S1
//   n_overflow  histogram (S1..S2)      -> place (S4..S5, every thread reads it); reset in the walk (after S5)
//   n_runs, n_win_loose, warp_tot       scan (S2..S4) -> scan / place (..S5); next writer after the next S2
//   items, rows  place (S4..S5)         -> walk (S5..S6); next writer after the next S4
//   p      walk (S5..S6)                -> finish (after S6); next writer after the next S5
//
// Row loads and reds drop 2.4x on the A2D / YTVOS encoder shapes in the init regime (642 per tile instead of 1532); what
// is left per item is one 128-byte shared-memory read and ~10 instructions.
#pragma once

// Window of the counting sort: kWinX x kWinY cells per level, kSortLevels levels (more levels: msda_bwd_tiled).
// 24 x 16 = +-3 sigma of the init regime's 2-pixel offsets around an 8 x 4 query tile.
constexpr int kWinX = 24, kWinY = 16, kWinCells = kWinX * kWinY, kSortLevels = 4;
constexpr int kSortCells = kSortLevels * kWinCells;

template <int ROUNDS> struct SortSmem {
    static constexpr int kQueries = 32;                          // 8 warps x 4 queries
    static constexpr int kPoints = ROUNDS * 8;                   // point slots per (query, head)
    static constexpr int kItems = kQueries * kPoints * 4;        // (query, point, corner)
    // Item slots.  Runs grow from the front (a cell of c items takes at most c + 1 <= 4c/3 run slots: full runs of
    // four, and a fourth, null, slot when three items remain), loose items from the back, so 4/3 kItems slots always
    // suffice; four more hold the null run the tail of the walk points at.
    static constexpr int kSlots = (kItems * 4 / 3 + 15) / 16 * 16;
    static constexpr uint32_t kDummySlot = kItems;               // p slot of the null items
    LevelTable lt;
    int stat[kSortLevels][4];                                    // sum x0, sum y0, points in range (this tile)
    uint32_t warp_tot[8];
    uint32_t n_runs, n_win_loose, n_overflow, pad_;
    alignas(16) uint32_t hist[kSortCells];                       // items per window cell (a plain u32 array: the atomics use all 32 banks)
    // per window cell, written by the scan: runs before | loose items before << 16 -- or, when the tile has at most 2048 items
    // (kPacked), everything the place phase needs in one word: runs before | loose before << 10 | (items - 1) << 21
    static constexpr bool kPacked = kItems <= 2048;
    alignas(16) uint32_t cell_pos[kSortCells];
    uint32_t rows[kItems];                                       // row (unit offset) of run r at [r], of loose item i at [kItems-1-i]
    alignas(16) uint2 items[kSlots + 4];                                     // {grad_out row offset | p slot << 16, a * w_corner}
    float4 go[kQueries][8];                                      // grad_out rows of the tile's queries, fp32
    float4 p[kQueries * kPoints + 1];                            // per (query, point): the four corner dot products (+ dummy)
};

// -DMSDA_CHECKED (tools/ builds only): every shared-memory index the sort derives from data is range-checked and the
// kernel traps on a violation -- compute-sanitizer is closed on the B200 pool (profiles/r2_compute_sanitizer_closed.log).
#ifdef MSDA_CHECKED
#define SORT_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define SORT_CHECK(cond) do { } while (0)
#endif

#ifndef MSDA_SORT_MINB
#define MSDA_SORT_MINB 4      // resident CTAs per SM the register budget is set for (4 -> 64 registers; ~55 KB of shared memory each)
#endif

// One sample point as its owner keeps it across the phases.
struct SortPoint {
    int x0, y0;          // top-left corner pixel, in [-1, W-1] x [-1, H-1] when the point is in range
    float lx, ly;
    float aa;            // attention weight (0 when the point is out of range, cuh:288 / :365)
    int valid;           // bit i: corner i lies inside the level
    int level;
};

__device__ __forceinline__ SortPoint sort_point(float loc_x, float loc_y, float a, int H, int W, int level) {
    SortPoint g;
    const float fw = (float)W, fh = (float)H;
    const float x = fmaf(loc_x, fw, -0.5f);          // cuh:285-286 as compiled (see point_geometry)
    const float y = fmaf(loc_y, fh, -0.5f);
    const bool in_range = (y > -1.f) && (x > -1.f) && (y < fh) && (x < fw);      // cuh:288; false for NaN / inf
    const float xf = floorf(x), yf = floorf(y);
    g.x0 = in_range ? (int)xf : 0;
    g.y0 = in_range ? (int)yf : 0;
    g.lx = in_range ? x - xf : 0.f;
    g.ly = in_range ? y - yf : 0.f;
    const bool xa = in_range && g.x0 >= 0, xb = in_range && g.x0 < W - 1;
    const bool ya = g.y0 >= 0, yb = g.y0 < H - 1;
    g.valid = (int)(xa && ya) | ((int)(xb && ya) << 1) | ((int)(xa && yb) << 2) | ((int)(xb && yb) << 3);
    g.aa = g.valid ? a : 0.f;
    g.level = level;
    return g;
}

// window origin along one axis: centred on the mean, clamped into the level
__device__ __forceinline__ int window_origin(int sum, int cnt, int win, int size) {
    const int mean = cnt ? __float2int_rn(__fdividef((float)sum, (float)cnt)) : 0;
    return max(min(mean - win / 2 + 1, size - win), 0);
}

// grad_value row (this lane's four channels) += r: one vector red.  GV16 = false: the fp32 tensor / accumulation buffer
// (REDG.E.ADD.F32x4); GV16 = true: straight into a bf16 grad_value as packed pairs (REDG.E.ADD.BF16x4: half the L2 atomic
// sectors, no fp32 buffer, no conversion pass -- but every red rounds the running sum in memory to bf16).
template <bool GV16>
__device__ __forceinline__ void red_run(void *slice_base, uint32_t unit, const Row &r, bool on) {
    if constexpr (!GV16) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                     "@q red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                     :: "l"(row_at(static_cast<float4 *>(slice_base), unit)), "f"(r.lo.x), "f"(r.lo.y), "f"(r.hi.x), "f"(r.hi.y),
                        "r"((uint32_t)on) : "memory");
    } else {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(r.lo.x, r.lo.y), hi = __floats2bfloat162_rn(r.hi.x, r.hi.y);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t"
                     "@q red.relaxed.gpu.global.add.noftz.v2.bf16x2 [%0], {%1,%2};\n\t}"
                     :: "l"(row_at(static_cast<uint2 *>(slice_base), unit)), "r"(*reinterpret_cast<const uint32_t *>(&lo)),
                        "r"(*reinterpret_cast<const uint32_t *>(&hi)), "r"((uint32_t)on) : "memory");
    }
}

template <typename VT, int ROUNDS, bool FUSED, bool GV16 = false>
__global__ void __launch_bounds__(256, MSDA_SORT_MINB)
msda_bwd_sorted(const VT *__restrict__ grad_out, const VT *__restrict__ value, const int64_t *__restrict__ shapes,
                const int64_t *__restrict__ start, const float *__restrict__ loc, const float *__restrict__ attn,
                void *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn, Dims d,
                FusedArgs fa) {
    using IO = RowIO<VT>;
    using Vec = typename IO::Vec;
    using SM = SortSmem<ROUNDS>;
    using GVec = typename std::conditional<GV16, uint2, float4>::type;      // one lane's four channels of a grad_value row
    constexpr int WARPS = 8;
    constexpr int kTaskQueries = Tile<WARPS>::kQueries;
    constexpr uint32_t kFull = 0xffffffffu;
    static_assert(kSortCells == 256 * 6 && kWinX == 4 * 6, "the scan gives every thread six x-adjacent cells");
    SM &sm = *reinterpret_cast<SM *>(msda_smem);
    LevelTable &lt = sm.lt;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 3, cl = lane & 7;
    const int qloc = warp * 4 + grp;                 // this lane group's query inside the tile
    const int pts = d.L * d.P;
    const int pixel_units = d.M * 8;

    load_level_table<WARPS>(lt, shapes, start, d.L, d.Lq);
    if (tid == 0) sm.n_overflow = 0;
    if (tid < 4) sm.items[SM::kSlots + tid] = make_uint2(SM::kDummySlot << 16, 0u);       // the null run
    for (int i = tid; i < kSortCells; i += 256) sm.hist[i] = 0u;
    if (tid < kSortLevels * 4) (&sm.stat[0][0])[tid] = 0;
    __syncthreads();
    const bool tiled = d.tiled && lt.dense;

    uint32_t lv = 0;     // level of this lane's point in each round, one byte per round
    int l_lo[ROUNDS], l_hi[ROUNDS];      // levels present in round r (the same for every lane)
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        lv |= (uint32_t)min((8 * r + cl) / d.P, d.L - 1) << (8 * r);
        l_lo[r] = min((8 * r) / d.P, d.L - 1);
        l_hi[r] = min((8 * r + 7) / d.P, d.L - 1);
    }

    const int tiles = tiled ? lt.tile_cum[d.L] : (d.Lq + kTaskQueries - 1) / kTaskQueries;
    for (TaskWalk t(d.N, d.M, tiles, d.fchunk, blockIdx.x, gridDim.x); t.next();) {
        const int m = t.m;
        const int q = select_query<WARPS>(tiled, d, lt, t.tile, warp, grp);
        const int64_t row = ((int64_t)t.n * d.Lq + max(q, 0)) * d.M + m;
        const int64_t slice = ((int64_t)t.n * d.S * d.M + m) * 8 + cl;      // unit offset of frame n, head m, this lane
        const Vec *vb = reinterpret_cast<const Vec *>(value) + slice;
        GVec *gb = static_cast<GVec *>(grad_value) + slice;

        // ---- stage: grad_out row -> shared memory, this lane's points -> registers, window statistics ----
        {
            Row go{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            if (q >= 0) go = IO::load_stream(reinterpret_cast<const Vec *>(grad_out) + (row * 8 + cl));
            sm.go[qloc][cl] = make_float4(go.lo.x, go.lo.y, go.hi.x, go.hi.y);
        }
        float2 xy[ROUNDS];
        float a[ROUNDS];
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
        SortPoint sp[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int l = (lv >> (8 * r)) & 0xff;
            sp[r] = sort_point(xy[r].x, xy[r].y, a[r], lt.H[l], lt.W[l], l);
            // mean top-left pixel of the tile's in-range points, per level
            for (int l2 = l_lo[r]; l2 <= l_hi[r]; ++l2) {
                const bool mine = sp[r].valid && l == l2;
                const int sx = __reduce_add_sync(kFull, mine ? sp[r].x0 : 0);
                const int sy = __reduce_add_sync(kFull, mine ? sp[r].y0 : 0);
                const int sc = __reduce_add_sync(kFull, mine ? 1 : 0);
                if (lane == 0 && sc) {
                    atomicAdd(&sm.stat[l2][0], sx);
                    atomicAdd(&sm.stat[l2][1], sy);
                    atomicAdd(&sm.stat[l2][2], sc);
                }
            }
        }
        __syncthreads();                                                                        // S1: statistics complete

        // ---- histogram: window origin per level, cell of each valid corner, rank inside the cell ----
        // Kept per round for the later phases: cell00 = window cell of the top-left corner (the others are +1, +kWinX,
        // +kWinX+1), mask = corners inside the window (bits 0-3) | valid corners outside it (bits 4-7), four 16-bit ranks.
        int cell00[ROUNDS];
        uint32_t mask[ROUNDS], rank[ROUNDS][2];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int l = sp[r].level;
            const int4 st = *reinterpret_cast<const int4 *>(sm.stat[l]);
            const int ux = sp[r].x0 - window_origin(st.x, st.z, kWinX, lt.W[l]);
            const int uy = sp[r].y0 - window_origin(st.y, st.z, kWinY, lt.H[l]);
            cell00[r] = l * kWinCells + uy * kWinX + ux;
            const uint32_t in_x = ((unsigned)ux < (unsigned)kWinX ? 5u : 0u) | ((unsigned)(ux + 1) < (unsigned)kWinX ? 10u : 0u);
            const uint32_t in_y = ((unsigned)uy < (unsigned)kWinY ? 3u : 0u) | ((unsigned)(uy + 1) < (unsigned)kWinY ? 12u : 0u);
            const uint32_t inw = in_x & in_y & (uint32_t)sp[r].valid;
            mask[r] = inw | (((uint32_t)sp[r].valid & ~inw) << 4);
        }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            // items outside their window: one ticket per thread and round (rare in the encoder's regimes)
            const uint32_t n_out = __popc(mask[r] >> 4);
            uint32_t out_rank = 0;
            if (n_out) out_rank = atomicAdd(&sm.n_overflow, n_out);
            uint32_t rk[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                rk[c] = out_rank;
                if ((mask[r] >> (4 + c)) & 1u) ++out_rank;
                if ((mask[r] >> c) & 1u) {
                    SORT_CHECK((unsigned)(cell00[r] + (c & 1) + (c >> 1) * kWinX) < (unsigned)kSortCells);
                    rk[c] = atomicAdd(&sm.hist[cell00[r] + (c & 1) + (c >> 1) * kWinX], 1u);
                }
            }
            rank[r][0] = rk[0] | (rk[1] << 16);
            rank[r][1] = rk[2] | (rk[3] << 16);
        }
        __syncthreads();                                                                        // S2: histogram complete

        // ---- scan: a cell of c items = c/4 full runs, one more (null-padded) run if three items remain, else c%4 loose items ----
        {
            const int c0 = tid * 6;      // six x-adjacent cells of one window row
            uint32_t cnt[6], tsum = 0;
#pragma unroll
            for (int k = 0; k < 6; k += 2) {
                const uint2 v = *reinterpret_cast<const uint2 *>(&sm.hist[c0 + k]);
                cnt[k] = v.x; cnt[k + 1] = v.y;
            }
            uint32_t pre[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {      // low half: runs, high half: loose items
                pre[k] = tsum;
                const uint32_t r = cnt[k] & 3u;
                tsum += (cnt[k] >> 2) + (r == 3u ? 1u : r << 16);
            }
            uint32_t incl = tsum;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t o = __shfl_up_sync(kFull, incl, s);
                if (lane >= s) incl += o;
            }
            if (lane == 31) sm.warp_tot[warp] = incl;
            if (tid < kSortLevels * 4) (&sm.stat[0][0])[tid] = 0;       // statistics are dead after S2: reset for the next tile
            __syncthreads();                                                                    // S3
            uint32_t excl = incl - tsum;
            for (int w = 0; w < warp; ++w) excl += sm.warp_tot[w];
            if (tsum) {
#pragma unroll
                for (int k = 0; k < 6; k += 2) {
                    uint32_t a0 = excl + pre[k], a1 = excl + pre[k + 1];
                    if constexpr (SM::kPacked) {
                        a0 = (a0 & 0xffffu) | ((a0 >> 16) << 10) | ((cnt[k] - 1u) << 21);
                        a1 = (a1 & 0xffffu) | ((a1 >> 16) << 10) | ((cnt[k + 1] - 1u) << 21);
                    }
                    *reinterpret_cast<uint2 *>(&sm.cell_pos[c0 + k]) = make_uint2(a0, a1);
                }
            }
            if (tid == 255) {
                sm.n_runs = (excl + tsum) & 0xffffu;
                sm.n_win_loose = (excl + tsum) >> 16;
            }
        }
        __syncthreads();                                                                        // S4: offsets complete

        // ---- place: every item to its slot; the first item of a run also records the run's row ----
        const int n_runs = (int)sm.n_runs;
        const uint32_t n_win_loose = sm.n_win_loose;
        const int n_loose = (int)(n_win_loose + sm.n_overflow);      // complete since S2; reset after S5
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            const int l = sp[r].level;
            const float hx = 1.f - sp[r].lx, hy = 1.f - sp[r].ly;
            const float aw4[4] = {sp[r].aa * (hy * hx), sp[r].aa * (hy * sp[r].lx), sp[r].aa * (sp[r].ly * hx), sp[r].aa * (sp[r].ly * sp[r].lx)};
            const uint32_t tag0 = (uint32_t)(qloc * 128) | ((uint32_t)((qloc * SM::kPoints + 8 * r + cl) * 4) << 16);
            const int row_step = lt.W[l] * pixel_units;
            const uint32_t unit00 = (uint32_t)((lt.start[l] + sp[r].y0 * lt.W[l] + sp[r].x0) * pixel_units);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool inw = (mask[r] >> c) & 1u, out = (mask[r] >> (4 + c)) & 1u;
                const uint32_t rk = (rank[r][c >> 1] >> (16 * (c & 1))) & 0xffffu;
                const uint32_t unit = unit00 + (uint32_t)((c & 1) * pixel_units + (c >> 1) * row_step);
                uint2 h = make_uint2(0u, 0u);      // {items of the cell, runs before | loose before << 16}
                if (inw) {
                    const int cell = cell00[r] + (c & 1) + (c >> 1) * kWinX;
                    if constexpr (SM::kPacked) {
                        const uint32_t w = sm.cell_pos[cell];
                        h = make_uint2((w >> 21) + 1u, (w & 0x3ffu) | (((w >> 10) & 0x7ffu) << 16));
                    } else {
                        h = make_uint2(sm.hist[cell], sm.cell_pos[cell]);
                    }
                }
                const uint32_t rem = h.x & 3u;
                const uint32_t run_items = rem == 3u ? h.x : h.x - rem;      // the first run_items ranks of the cell go to runs
                const bool in_run = rk < run_items;                           // false outside the window (run_items = 0)
                const uint32_t li = inw ? (h.y >> 16) + rk - run_items : n_win_loose + rk;
                const uint32_t pos = in_run ? 4u * (h.y & 0xffffu) + rk : (uint32_t)(SM::kSlots - 1) - li;
                const uint32_t row_at_ = in_run ? (h.y & 0xffffu) + (rk >> 2) : (uint32_t)(SM::kItems - 1) - li;
                if (inw || out) {
                    SORT_CHECK(pos < (uint32_t)SM::kSlots && row_at_ < (uint32_t)SM::kItems && (!inw || rk < h.x));
                    SORT_CHECK(!in_run || (pos < 4u * (uint32_t)n_runs && row_at_ < (uint32_t)n_runs));
                    SORT_CHECK(in_run || (li < (uint32_t)n_loose && 4u * (uint32_t)n_runs + li < (uint32_t)SM::kSlots));
                    sm.items[pos] = make_uint2(tag0 + ((uint32_t)c << 16), __float_as_uint(aw4[c]));
                    if (!in_run || !(rk & 3u)) sm.rows[row_at_] = unit;
                    if (in_run && rem == 3u && rk + 1 == h.x) sm.items[pos + 1] = make_uint2(SM::kDummySlot << 16, 0u);      // null item pads the run
                }
            }
        }
        __syncthreads();                                                                        // S5: items complete

        // ---- walk ----
        {
            if (tid == 0) sm.n_overflow = 0;          // every thread read it before S5; its next writer comes after S1
            // the histogram is dead: clear it for the next tile (its next writer comes after S6 / S1)
            for (int i = tid; i < kSortCells / 4; i += 256) reinterpret_cast<uint4 *>(sm.hist)[i] = make_uint4(0u, 0u, 0u, 0u);
            const char *go_lane = reinterpret_cast<const char *>(&sm.go[0][cl]);
            float *p_flat = reinterpret_cast<float *>(sm.p);
            auto go_row = [&](uint32_t tag) -> Row {
                const float4 g4 = *reinterpret_cast<const float4 *>(go_lane + (tag & 0xffffu));
                return Row{make_float2(g4.x, g4.y), make_float2(g4.z, g4.w)};
            };
            // runs: one per lane group, four items each, all of the same row
            {
                int run = warp * 4 + grp;
                bool has = run < n_runs;
                uint32_t unit = has ? sm.rows[run] : 0u;
                const uint4 *ip = reinterpret_cast<const uint4 *>(&sm.items[has ? 4 * run : SM::kSlots]);
                Row v = IO::load_stream(row_at(vb, unit));
                const bool up4 = (cl & 4) != 0, up2 = (cl & 2) != 0, even = !(cl & 1);
#pragma unroll 2
                for (int base = warp * 4; base < n_runs; base += 32) {
                    run += 32;
                    const bool has_next = run < n_runs;
                    const uint32_t unit_next = has_next ? sm.rows[run] : 0u;
                    const uint4 *ip_next = reinterpret_cast<const uint4 *>(&sm.items[has_next ? 4 * run : SM::kSlots]);
                    const Row v_next = IO::load_stream(row_at(vb, unit_next));      // next row in flight during this run
                    const uint4 i01 = ip[0], i23 = ip[1];                            // {tag, a*w} of items 0,1 and 2,3
                    const uint32_t tag[4] = {i01.x, i01.z, i23.x, i23.z};
                    SORT_CHECK((tag[0] & 0xffffu) < 4096u && (tag[3] & 0xffffu) < 4096u && (tag[0] >> 16) <= SM::kDummySlot && (tag[3] >> 16) <= SM::kDummySlot);
                    const float aw[4] = {__uint_as_float(i01.y), __uint_as_float(i01.w), __uint_as_float(i23.y), __uint_as_float(i23.w)};
                    Row acc{make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
                    float ds[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const Row g = go_row(tag[k]);
                        fma_row(aw[k], g, acc);          // grad_value[row] += (a * w_corner) * grad_out   (cuh:125,134,143,152)
                        ds[k] = dot_row(g, v);           // this lane's four channels of <grad_out, value row>
                    }
                    // reduce over the 8 lanes of the group: lanes 2j, 2j+1 end with the sum of item j
                    const float e0 = (up4 ? ds[2] : ds[0]) + __shfl_xor_sync(kFull, up4 ? ds[0] : ds[2], 4);
                    const float e1 = (up4 ? ds[3] : ds[1]) + __shfl_xor_sync(kFull, up4 ? ds[1] : ds[3], 4);
                    float tot = (up2 ? e1 : e0) + __shfl_xor_sync(kFull, up2 ? e0 : e1, 2);
                    tot += __shfl_xor_sync(kFull, tot, 1);
                    const uint32_t my_tag = up4 ? (up2 ? tag[3] : tag[2]) : (up2 ? tag[1] : tag[0]);
                    if (even) p_flat[my_tag >> 16] = tot;      // null items: the dummy slot
                    red_run<GV16>(gb, unit, acc, has);
                    has = has_next; unit = unit_next; ip = ip_next; v = v_next;
                }
            }
            // loose items: four per lane group and step, each with its own row load and its own red; their dot products share
            // one butterfly -- the same summation tree as the runs' (lane pairs 4 apart first): whether an item lands in a run or
            // here depends on the order of the histogram atomics, and d/d location, d/d attention stay bitwise reproducible
            for (int base = warp * 4; base < n_loose; base += 128) {
                uint2 it[4];
                uint32_t unit[4];
                Row v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int li = base + 32 * u + grp;
                    const bool has = li < n_loose;
                    it[u] = sm.items[has ? SM::kSlots - 1 - li : SM::kSlots];      // none: the null item (weight 0, dummy p slot)
                    unit[u] = has ? sm.rows[SM::kItems - 1 - li] : 0xffffffffu;
                    v[u] = IO::load_stream(row_at(vb, has ? unit[u] : 0u));
                }
                float ds[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const Row g = go_row(it[u].x);
                    const float aw = __uint_as_float(it[u].y);
                    ds[u] = dot_row(g, v[u]);
                    red_run<GV16>(gb, unit[u], Row{__fmul2_rn(splat(aw), g.lo), __fmul2_rn(splat(aw), g.hi)}, unit[u] != 0xffffffffu);
                }
                const bool up4 = (cl & 4) != 0, up2 = (cl & 2) != 0;
                const float e0 = (up4 ? ds[2] : ds[0]) + __shfl_xor_sync(kFull, up4 ? ds[0] : ds[2], 4);
                const float e1 = (up4 ? ds[3] : ds[1]) + __shfl_xor_sync(kFull, up4 ? ds[1] : ds[3], 4);
                float tot = (up2 ? e1 : e0) + __shfl_xor_sync(kFull, up2 ? e0 : e1, 2);
                tot += __shfl_xor_sync(kFull, tot, 1);
                const uint32_t my_tag = up4 ? (up2 ? it[3].x : it[2].x) : (up2 ? it[1].x : it[0].x);
                if (!(cl & 1)) p_flat[my_tag >> 16] = tot;
            }
        }
        __syncthreads();                                                                        // S6: dot products complete

        // ---- finish: the owner of a point combines its four corner dot products (cuh:123-158 regrouped by corner) ----
        float g_prob[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            g_prob[r] = 0.f;
            const int pt = 8 * r + cl;
            if (!(q >= 0 && pt < pts)) continue;
            const int l = sp[r].level;
            const float4 ps = sm.p[qloc * SM::kPoints + pt];
            const int valid = sp[r].valid;
            const float lx = sp[r].lx, ly = sp[r].ly, a_own = sp[r].aa;
            const float hx = 1.f - lx, hy = 1.f - ly;
            // corners outside the level were never items: zero padding (cuh:56-78)
            const float q00 = (valid & 1) ? ps.x : 0.f, q01 = (valid & 2) ? ps.y : 0.f;
            const float q10 = (valid & 4) ? ps.z : 0.f, q11 = (valid & 8) ? ps.w : 0.f;
            const float ga = (hy * hx) * q00 + (hy * lx) * q01 + (ly * hx) * q10 + (ly * lx) * q11;       // :156
            const float gx = hy * (q01 - q00) + ly * (q11 - q10);                                         // :157
            const float gy = hx * (q10 - q00) + lx * (q11 - q01);                                         // :158
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
        // no barrier here: the next tile's first shared-memory writes (go rows, statistics) touch nothing this phase
        // reads, and S1 of the next tile orders everything else
    }
}
