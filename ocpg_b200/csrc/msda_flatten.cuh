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
// 32 x 32 tiles through shared memory (padded: conflict-free both ways), 128-byte coalesced rows on both sides; a CTA of
// 256 threads walks tiles of every (level, frame) in one flattened index space.

constexpr int kFlatMaxLevels = 8;
struct FlattenLevel {
    const float *src;       // [N][C][hw]
    const float *pos;       // [N][C][hw] or null
    float *map_out;         // unflatten: [N][C][hw]
    int hw, start;          // pixels of the level, first row in the flattened layout
    int tiles_hw;           // ceil(hw / 32)
    int tile_begin;         // first tile index of the level (tiles are counted per frame)
};
struct FlattenArgs {
    FlattenLevel lv[kFlatMaxLevels];
    int L, N, C, S;
    int tiles_c;            // ceil(C / 32)
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
    hw0 = (r - tc * a.lv[l].tiles_hw) * 32;
    c0 = tc * 32;
}

template <bool WITH_POS>
__global__ void __launch_bounds__(256)
flatten_levels_kernel(const __grid_constant__ FlattenArgs a) {
    __shared__ float tile[WITH_POS ? 2 : 1][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        int n, l, hw0, c0;
        flatten_tile_coords(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
        const int hw = hw0 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                 // rows c0 + ty + 8k of [C][hw]: coalesced along hw
            const int c = c0 + ty + 8 * k;
            if (c < a.C && hw < lv.hw) {
                const int64_t i = ((int64_t)n * a.C + c) * lv.hw + hw;
                tile[0][ty + 8 * k][tx] = __ldg(lv.src + i);
                if constexpr (WITH_POS) tile[1][ty + 8 * k][tx] = __ldg(lv.pos + i) + (a.level_embed ? __ldg(a.level_embed + l * a.C + c) : 0.f);
            }
        }
        __syncthreads();
        const int c = c0 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                 // rows hw0 + ty + 8k of [S][C]: coalesced along c
            const int p = hw0 + ty + 8 * k;
            if (c < a.C && p < lv.hw) {
                const int64_t o = ((int64_t)n * a.S + lv.start + p) * a.C + c;
                a.src_flat[o] = tile[0][tx][ty + 8 * k];
                if constexpr (WITH_POS) a.pos_flat[o] = tile[1][tx][ty + 8 * k];
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
unflatten_levels_kernel(const __grid_constant__ FlattenArgs a) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t total = (int64_t)a.N * a.tiles_per_frame;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        int n, l, hw0, c0;
        flatten_tile_coords(a, t, n, l, hw0, c0);
        const FlattenLevel &lv = a.lv[l];
        const int c = c0 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int p = hw0 + ty + 8 * k;
            if (c < a.C && p < lv.hw) tile[ty + 8 * k][tx] = __ldg(a.flat_in + ((int64_t)n * a.S + lv.start + p) * a.C + c);
        }
        __syncthreads();
        const int hw = hw0 + tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cc = c0 + ty + 8 * k;
            if (cc < a.C && hw < lv.hw) lv.map_out[((int64_t)n * a.C + cc) * lv.hw + hw] = tile[tx][ty + 8 * k];
        }
        __syncthreads();
    }
}
