// msda_decoder.cuh -- the decoder-side consumers of the hot path, included by
// msda_sm100.cu inside its anonymous namespace.
//
// Around every cross-attention the reference's DeformableTransformerDecoder (models/deformable_transformer.py:353-375)
//   (1) scales the reference points by the per-level valid ratios          (:358-363)   1-2 elementwise launches
//   (2) divides the returned sampling locations by the valid ratios        (:368)       1 launch over N*Lq*M*L*P*2
//   (3) takes the 30 largest of the M*L*P = 128 attention weights per query (:372)      topk: sort-based, several launches
//   (4) gathers the sampling locations of those 30                          (:375)      repeat + gather
// on tensors of a few KB (config 4: 25 queries): pure launch latency.  Here (2)-(4) are ONE launch -- a warp per query
// keeps the 128 weights in registers (4 per lane), extracts the top-k by repeated warp arg-max (two redux.sync per element),
// and only the k selected locations are ever divided -- and (1) is one launch.
//
// Order of the selection = torch.topk(sorted=True): descending weight; equal weights in ascending index order (torch
// leaves the order of ties unspecified); NaN counts as the largest value, as in torch.

// Weights as order-preserving 32-bit keys (sign flip for positives, complement for negatives: NaN sorts last = largest,
// like torch.topk), so that one redux.sync (a single warp-wide max instruction) finds the winner's weight and a second one
// (min over the tied lanes' indices) its position: 2 warp reductions per extracted element instead of 5 x 2 dependent
// shuffle steps -- the selection is a chain of `top` dependent iterations, so their latency is the kernel's run time.
__device__ __forceinline__ uint32_t order_key(float w) {
    const uint32_t b = __float_as_uint(w);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ITEMS = ceil(K / 32) weights per lane; top <= 32 (lane t keeps the t-th winner).
template <int ITEMS>
__global__ void __launch_bounds__(128)
decoder_select_samples_kernel(const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ valid_ratios,
                              int64_t rows, int Lq, int K, int L, int P, int top, float *__restrict__ samples_out,
                              float *__restrict__ weights_out, int64_t *__restrict__ idx_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (n, q)
    if (row >= rows) return;
    uint32_t key[ITEMS];                  // 0 = taken / past the end (every real key is > 0 except one NaN pattern)
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int k = lane + 32 * j;
        key[j] = k < K ? order_key(__ldg(attn + row * K + k)) : 0u;
    }
    uint32_t my_key = 0;
    int my_k = 0;
    for (int t = 0; t < top; ++t) {
        uint32_t best = key[0];
#pragma unroll
        for (int j = 1; j < ITEMS; ++j) best = max(best, key[j]);
        const uint32_t m = __reduce_max_sync(0xffffffffu, best);
        // among the lanes that hold the maximum: the smallest index (ascending j = ascending k within a lane)
        int cand = 0x7fffffff;
#pragma unroll
        for (int j = ITEMS - 1; j >= 0; --j)
            if (key[j] == m) cand = lane + 32 * j;
        const int win = (int)__reduce_min_sync(0xffffffffu, (uint32_t)cand);
        if ((win & 31) == lane) {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j)
                if (j == (win >> 5)) key[j] = 0u;
        }
        if (lane == t) { my_key = m; my_k = win; }
    }
    if (lane < top) {
        const int n = (int)(row / Lq);
        const int level = (my_k / P) % L;                                  // k = (m * L + l) * P + p
        const float2 s = __ldg(reinterpret_cast<const float2 *>(loc) + row * K + my_k);
        const float2 vr = __ldg(reinterpret_cast<const float2 *>(valid_ratios) + (int64_t)n * L + level);
        reinterpret_cast<float2 *>(samples_out)[row * top + lane] = make_float2(__fdiv_rn(s.x, vr.x), __fdiv_rn(s.y, vr.y));   // :368
        if (weights_out) weights_out[row * top + lane] = key_value(my_key);
        if (idx_out) idx_out[row * top + lane] = my_k;
    }
}

// reference_points_input[n, q, l, :] = reference_points[n, q, :] * valid_ratios[n, l, :] (2-d), or
//                                     = reference_points[n, q, :] * (vr, vr)[n, l, :]      (4-d boxes)        (:358-363)
__global__ void __launch_bounds__(256)
decoder_reference_points_kernel(const float *__restrict__ ref, const float *__restrict__ valid_ratios, int64_t total, int Lq, int L,
                                int ref_dim, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(i % L);
        const int64_t nq = i / L;
        const int n = (int)(nq / Lq);
        const float2 vr = __ldg(reinterpret_cast<const float2 *>(valid_ratios) + (int64_t)n * L + l);
        if (ref_dim == 2) {
            const float2 r = __ldg(reinterpret_cast<const float2 *>(ref) + nq);
            reinterpret_cast<float2 *>(out)[i] = make_float2(r.x * vr.x, r.y * vr.y);
        } else {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(ref) + nq);
            reinterpret_cast<float4 *>(out)[i] = make_float4(r.x * vr.x, r.y * vr.y, r.z * vr.x, r.w * vr.y);
        }
    }
}
