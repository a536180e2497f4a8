// msda_decoder.cuh -- the decoder-side consumers of the hot path, included by
// msda_sm100.cu inside its anonymous namespace.
//
// Around every cross-attention the reference's DeformableTransformerDecoder (models/deformable_transformer.py:353-375)
//   (1) scales the reference points by the per-level valid ratios          (:358-363)   1-2 elementwise launches
//   (2) divides the returned sampling locations by the valid ratios        (:368)       1 launch over N*Lq*M*L*P*2
//   (3) takes the 30 largest of the M*L*P = 128 attention weights per query (:372)      topk: sort-based, several launches
//   (4) gathers the sampling locations of those 30                          (:375)      repeat + gather
// on tensors of a few KB (config 4: 25 queries): pure launch latency.  Here (2)-(4) are ONE launch -- a warp per query
// keeps the 128 weights in registers (4 per lane), extracts the top-k by repeated warp arg-max (k x 5 shuffle steps),
// and only the k selected locations are ever divided -- and (1) is one launch.
//
// Order of the selection = torch.topk(sorted=True): descending weight; equal weights in ascending index order (torch
// leaves the order of ties unspecified).  NaN weights are not ordered specially (softmax outputs have none).

// ITEMS = ceil(K / 32) weights per lane; top <= 32 (lane t keeps the t-th winner).
template <int ITEMS>
__global__ void __launch_bounds__(128)
decoder_select_samples_kernel(const float *__restrict__ loc, const float *__restrict__ attn, const float *__restrict__ valid_ratios,
                              int64_t rows, int Lq, int K, int L, int P, int top, float *__restrict__ samples_out,
                              float *__restrict__ weights_out, int64_t *__restrict__ idx_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (n, q)
    if (row >= rows) return;
    float w[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
        const int k = lane + 32 * j;
        w[j] = k < K ? __ldg(attn + row * K + k) : -INFINITY;
    }
    uint32_t taken = 0;                   // bit j: this lane's item j was selected
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
        if (lane + 32 * j >= K) taken |= 1u << j;
    float my_w = 0.f;
    int my_k = 0;
    for (int t = 0; t < top; ++t) {
        // this lane's best remaining item (lowest index among equals: ascending j = ascending k)
        float bw = -INFINITY;
        int bk = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const bool free_ = !((taken >> j) & 1u);
            if (free_ && (bk == 0x7fffffff || w[j] > bw)) { bw = w[j]; bk = lane + 32 * j; }
        }
        // warp arg-max on (weight desc, index asc)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float ow = __shfl_xor_sync(0xffffffffu, bw, s);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, s);
            const bool better = ok != 0x7fffffff && (bk == 0x7fffffff || ow > bw || (ow == bw && ok < bk));
            if (better) { bw = ow; bk = ok; }
        }
        if ((bk & 31) == lane && bk != 0x7fffffff) taken |= 1u << (bk >> 5);
        if (lane == t) { my_w = bw; my_k = bk; }
    }
    if (lane < top) {
        const int n = (int)(row / Lq);
        const int level = (my_k / P) % L;                                  // k = (m * L + l) * P + p
        const float2 s = __ldg(reinterpret_cast<const float2 *>(loc) + row * K + my_k);
        const float2 vr = __ldg(reinterpret_cast<const float2 *>(valid_ratios) + (int64_t)n * L + level);
        reinterpret_cast<float2 *>(samples_out)[row * top + lane] = make_float2(__fdiv_rn(s.x, vr.x), __fdiv_rn(s.y, vr.y));   // :368
        if (weights_out) weights_out[row * top + lane] = my_w;
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
