// msda_flatten.cuh -- the layout traffic either side of the encoder, included by
// msda_sm100.cu inside its anonymous namespace.
//
// The reference's DeformableTransformer.forward turns the L feature maps [N][C][H_l][W_l] into the operator's layout
// [N][S][C] with flatten(2).transpose(1, 2) per level + torch.cat, does the same for the positional embeddings after adding
// the level embedding (models/deformable_transformer.py:149-169), and converts the encoder's output back into maps with
// reshape + permute + contiguous per level (:205-212): 4 adds, 2 strided cats and 3 strided copies.  These are
// transpositions of [C][H*W] blocks, i.e. pure HBM streams:
//   flatten_levels_kernel     all levels of src (and pos + level_embed) in ONE launch: reads 1 (2), writes 1 (2) matrices
//   unflatten_levels_kernel   [N][S][C] -> per-level [N][C][H*W]; also the backward of the former (and vice versa)
// Tiles of HT pixels x CT channels through shared memory (padded: conflict-free both ways); a CTA of 256 threads walks
// tiles of every (level, frame) in one flattened index space.  What matters is the length of the contiguous runs a tile
// touches in DRAM on BOTH sides (HT*4 bytes on the map side, CT*4 bytes on the flattened side): see DESIGN.md.

constexpr int kFlatMaxLevels = 8;
#ifndef MSDA_FLAT_HT
#define MSDA_FLAT_HT 64
#endif
#ifndef MSDA_FLAT_CT
#define MSDA_FLAT_CT 64
#endif
constexpr int kFlatTileHW = MSDA_FLAT_HT;    // pixels per tile   (multiples of 32; runs of HT*4 bytes on the map side)
constexpr int kFlatTileC = MSDA_FLAT_CT;     // channels per tile (runs of CT*4 bytes on the flattened side)
constexpr int kFlatPerThread = kFlatTileHW * kFlatTileC / 256;
struct FlattenLevel {
    const float *src;       // [N][C][hw]
    const float *pos;       // [N][C][hw] or null
    float *map_out;         // unflatten: [N][C][hw]
    int hw, start;          // pixels of the level, first row in the flattened layout
    int tiles_hw;           // ceil(hw / kFlatTileHW)
    int tile_begin;         // first tile index of the level (tiles are counted per frame)
};
struct FlattenArgs {
    FlattenLevel lv[kFlatMaxLevels];
    int L, N, C, S;
    int tiles_c;            // ceil(C / kFlatTileC)
    int tiles_per_frame;    // over all levels
    const float *level_embed;      // [L][C] or null
    float *src_flat, *pos_flat;    // [N][S][C]
    const float *flat_in;          // unflatten: [N][S][C]
};

__device__ __forceinline__ void flatten_tile_coords(const FlattenArgs &a, int64_t t, int &n, int &l, int &hw0, int &c0) {
    n = (int)(t / a.tiles_per_frame);
    int r = (int)(t - (int64_t)n * a.tiles_per_frame);
    l = 0;
    while (l + 1 < a.L && r >= a.lv[l + 1].tile_begin) ++l;
    r -= a.lv[l].tile_begin;
    const int tc = r / a.lv[l].tiles_hw;
    hw0 = (r - tc * a.lv[l].tiles_hw) * kFlatTileHW;
    c0 = tc * kFlatTileC;
}

// A tile is moved as 128-byte segments, one per warp instruction: on the map side segment g = (channel g / (HT/32),
// pixels 32 * (g % (HT/32)) ...), on the flattened side g = (pixel g / (CT/32), channels 32 * (g % (CT/32)) ...); warp w
// takes segments w, w + 8, ...
//
// Loads are cp.async (LDGSTS, 4 bytes per lane: map rows are only 4-byte aligned in general) straight into a ring of
// kFlatStages shared-memory tiles, so the loads of the next tiles are in flight while the current one is stored; the
// register-staged first version (load -> barrier -> store, nothing in flight during the store phase) reached 28-59 % of
// the HBM peak whatever the tile shape: it was bound by latency, not by DRAM run lengths.
constexpr int kFlatStages = 3;

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc, bool on) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int bytes = on ? 4 : 0;          // 0: nothing is read, the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <bool WITH_POS>
__global__ void __launch_bounds__(256)
flatten_levels_kernel(const __grid_constant__ FlattenArgs a) {
    constexpr int kHS = kFlatTileHW / 32, kCS = kFlatTileC / 32, kT = WITH_POS ? 2 : 1;
    using Tile = float[kFlatTileC][kFlatTileHW + 1];
    Tile *ring = reinterpret_cast<Tile *>(msda_smem);                 // [kFlatStages][kT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;

    auto issue = [&](int64_t t, int stage) {                           // all loads of tile t -> ring[stage]
        if (t < total) {
            int n, l, hw0, c0;
            flatten_tile_coords(a, t, n, l, hw0, c0);
            const FlattenLevel &lv = a.lv[l];
#pragma unroll
            for (int k = 0; k < kFlatPerThread; ++k) {
                const int g = warp + 8 * k, cl = g / kHS, hl = (g % kHS) * 32 + lane;
                const int c = c0 + cl, hw = hw0 + hl;
                const bool on = c < a.C && hw < lv.hw;
                const int64_t i = on ? ((int64_t)n * a.C + c) * lv.hw + hw : 0;
                cp_async_f32(&ring[stage * kT][cl][hl], lv.src + i, on);
                if constexpr (WITH_POS) cp_async_f32(&ring[stage * kT + 1][cl][hl], lv.pos + i, on);
            }
        }
        cp_async_commit();
    };

    int64_t t = blockIdx.x;
#pragma unroll
    for (int s = 0; s < kFlatStages - 1; ++s) issue(t + (int64_t)s * gridDim.x, s);
    for (int it = 0; t < total; t += gridDim.x, ++it) {
        cp_async_wait<kFlatStages - 2>();                              // tile `it` has landed (for this thread)
        __syncthreads();                                               // ... for all threads; and tile it-1's buffer is free
        issue(t + (int64_t)(kFlatStages - 1) * gridDim.x, (it + kFlatStages - 1) % kFlatStages);
        const int stage = it % kFlatStages;
        int n, l, hw0, c0;
        flatten_tile_coords(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
        float le[kCS];
#pragma unroll
        for (int j = 0; j < kCS; ++j) {
            const int c = c0 + j * 32 + lane;
            le[j] = (WITH_POS && a.level_embed && c < a.C) ? __ldg(a.level_embed + l * a.C + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kFlatPerThread; ++k) {
            const int g = warp + 8 * k, pl = g / kCS, cl = (g % kCS) * 32 + lane;
            const int p = hw0 + pl, c = c0 + cl;
            if (c < a.C && p < lv.hw) {
                const int64_t o = ((int64_t)n * a.S + lv.start + p) * a.C + c;
                a.src_flat[o] = ring[stage * kT][cl][pl];
                if constexpr (WITH_POS) a.pos_flat[o] = ring[stage * kT + 1][cl][pl] + le[(warp + 8 * k) % kCS];
            }
        }
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(256)
unflatten_levels_kernel(const __grid_constant__ FlattenArgs a) {
    constexpr int kHS = kFlatTileHW / 32, kCS = kFlatTileC / 32;
    using Tile = float[kFlatTileHW][kFlatTileC + 1];
    Tile *ring = reinterpret_cast<Tile *>(msda_smem);                 // [kFlatStages]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;

    auto issue = [&](int64_t t, int stage) {
        if (t < total) {
            int n, l, hw0, c0;
            flatten_tile_coords(a, t, n, l, hw0, c0);
            const FlattenLevel &lv = a.lv[l];
#pragma unroll
            for (int k = 0; k < kFlatPerThread; ++k) {
                const int g = warp + 8 * k, pl = g / kCS, cl = (g % kCS) * 32 + lane;
                const int p = hw0 + pl, c = c0 + cl;
                const bool on = c < a.C && p < lv.hw;
                cp_async_f32(&ring[stage][pl][cl], a.flat_in + (on ? ((int64_t)n * a.S + lv.start + p) * a.C + c : 0), on);
            }
        }
        cp_async_commit();
    };

    int64_t t = blockIdx.x;
#pragma unroll
    for (int s = 0; s < kFlatStages - 1; ++s) issue(t + (int64_t)s * gridDim.x, s);
    for (int it = 0; t < total; t += gridDim.x, ++it) {
        cp_async_wait<kFlatStages - 2>();
        __syncthreads();
        issue(t + (int64_t)(kFlatStages - 1) * gridDim.x, (it + kFlatStages - 1) % kFlatStages);
        const int stage = it % kFlatStages;
        int n, l, hw0, c0;
        flatten_tile_coords(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
#pragma unroll
        for (int k = 0; k < kFlatPerThread; ++k) {
            const int g = warp + 8 * k, cl = g / kHS, hl = (g % kHS) * 32 + lane;
            const int c = c0 + cl, hw = hw0 + hl;
            if (c < a.C && hw < lv.hw) lv.map_out[((int64_t)n * a.C + c) * lv.hw + hw] = ring[stage][hl][cl];
        }
    }
    cp_async_wait<0>();
}

// ---- 16-byte path: every level has H*W % 4 == 0, C % 4 == 0, 16-byte aligned pointers (the production shapes) ----
// ncu on the scalar kernels above: 61-69 % of the issue slots busy at 26-41 % DRAM throughput -- with one 4-byte element
// per load / cp.async / LDS / store instruction they are bound by instruction issue, not by memory.  Here a thread owns a
// 4 x 4 block (4 channels x 4 pixels): 4 LDG.128 along the pixels, the transposition is a renaming of registers, 4
// STS.128 into a [pixel][channel] tile whose 16-byte slots are XOR-swizzled by the pixel group (conflict-free for both
// the column-wise writes and the row-wise reads), then LDS.128 + STG.128 along the channels: 16 memory instructions per
// 16 elements instead of 64, index arithmetic amortised 4x.  Tile = 64 pixels x 64 channels, one block per thread.
constexpr int kV4Tile = 64;

__device__ __forceinline__ void flatten_tile_coords_v4(const FlattenArgs &a, int64_t t, int &n, int &l, int &hw0, int &c0) {
    n = (int)(t / a.tiles_per_frame);
    int r = (int)(t - (int64_t)n * a.tiles_per_frame);
    l = 0;
    while (l + 1 < a.L && r >= a.lv[l + 1].tile_begin) ++l;
    r -= a.lv[l].tile_begin;
    const int tc = r / a.lv[l].tiles_hw;
    hw0 = (r - tc * a.lv[l].tiles_hw) * kV4Tile;
    c0 = tc * kV4Tile;
}

template <bool WITH_POS>
__global__ void __launch_bounds__(256)
flatten_levels_v4_kernel(const __grid_constant__ FlattenArgs a) {
    __shared__ float4 tile[WITH_POS ? 2 : 1][kV4Tile][kV4Tile / 4];      // [pixel][16-byte channel slot, swizzled]
    const int hg = threadIdx.x & 15, cg = threadIdx.x >> 4;              // this thread's block: pixels 4hg.., channels 4cg..
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        int n, l, hw0, c0;
        flatten_tile_coords_v4(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
        const int hw = hw0 + 4 * hg, c = c0 + 4 * cg;
        if (hw < lv.hw && c < a.C) {
            const int64_t i = ((int64_t)n * a.C + c) * lv.hw + hw;
            const int slot = cg ^ (hg & 7);
            float4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) r[j] = __ldg(reinterpret_cast<const float4 *>(lv.src + i + (int64_t)j * lv.hw));
            tile[0][4 * hg + 0][slot] = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
            tile[0][4 * hg + 1][slot] = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
            tile[0][4 * hg + 2][slot] = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
            tile[0][4 * hg + 3][slot] = make_float4(r[0].w, r[1].w, r[2].w, r[3].w);
            if constexpr (WITH_POS) {
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = __ldg(reinterpret_cast<const float4 *>(lv.pos + i + (int64_t)j * lv.hw));
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.level_embed) e = __ldg(reinterpret_cast<const float4 *>(a.level_embed + l * a.C + c));
                tile[1][4 * hg + 0][slot] = make_float4(r[0].x + e.x, r[1].x + e.y, r[2].x + e.z, r[3].x + e.w);
                tile[1][4 * hg + 1][slot] = make_float4(r[0].y + e.x, r[1].y + e.y, r[2].y + e.z, r[3].y + e.w);
                tile[1][4 * hg + 2][slot] = make_float4(r[0].z + e.x, r[1].z + e.y, r[2].z + e.z, r[3].z + e.w);
                tile[1][4 * hg + 3][slot] = make_float4(r[0].w + e.x, r[1].w + e.y, r[2].w + e.z, r[3].w + e.w);
            }
        }
        __syncthreads();
        const int c4 = threadIdx.x & 15;                                   // 16-byte channel slot of the flattened row
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pl = (threadIdx.x >> 4) + 16 * k, p = hw0 + pl, cc = c0 + 4 * c4;
            if (p < lv.hw && cc < a.C) {
                const int64_t o = ((int64_t)n * a.S + lv.start + p) * a.C + cc;
                const int slot = c4 ^ ((pl >> 2) & 7);
                *reinterpret_cast<float4 *>(a.src_flat + o) = tile[0][pl][slot];
                if constexpr (WITH_POS) *reinterpret_cast<float4 *>(a.pos_flat + o) = tile[1][pl][slot];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
unflatten_levels_v4_kernel(const __grid_constant__ FlattenArgs a) {
    __shared__ float4 tile[kV4Tile][kV4Tile / 4];
    const int hg = threadIdx.x & 15, cg = threadIdx.x >> 4;
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        int n, l, hw0, c0;
        flatten_tile_coords_v4(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
        const int c4 = threadIdx.x & 15;
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pl = (threadIdx.x >> 4) + 16 * k, p = hw0 + pl, cc = c0 + 4 * c4;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < lv.hw && cc < a.C) v[k] = __ldg(reinterpret_cast<const float4 *>(a.flat_in + ((int64_t)n * a.S + lv.start + p) * a.C + cc));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pl = (threadIdx.x >> 4) + 16 * k;
            tile[pl][c4 ^ ((pl >> 2) & 7)] = v[k];
        }
        __syncthreads();
        const int hw = hw0 + 4 * hg, c = c0 + 4 * cg;
        if (hw < lv.hw && c < a.C) {
            const int slot = cg ^ (hg & 7);
            const float4 p0 = tile[4 * hg + 0][slot], p1 = tile[4 * hg + 1][slot], p2 = tile[4 * hg + 2][slot], p3 = tile[4 * hg + 3][slot];
            float *o = lv.map_out + ((int64_t)n * a.C + c) * lv.hw + hw;
            *reinterpret_cast<float4 *>(o) = make_float4(p0.x, p1.x, p2.x, p3.x);
            *reinterpret_cast<float4 *>(o + lv.hw) = make_float4(p0.y, p1.y, p2.y, p3.y);
            *reinterpret_cast<float4 *>(o + 2 * (int64_t)lv.hw) = make_float4(p0.z, p1.z, p2.z, p3.z);
            *reinterpret_cast<float4 *>(o + 3 * (int64_t)lv.hw) = make_float4(p0.w, p1.w, p2.w, p3.w);
        }
        __syncthreads();
    }
}
