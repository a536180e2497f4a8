// msda_epilogue.cuh -- the encoder layer's epilogue around the hot path, included by
// msda_sm100.cu inside its anonymous namespace.
//
// The reference's DeformableTransformerEncoderLayer (models/deformable_transformer.py:243-260) follows the attention
// and the FFN with `src = norm(src + dropout(proj(x) + bias))`.  In eager PyTorch that is a bias add, a residual add and
// a LayerNorm forward; backward it is LayerNorm's grad-input kernel, its gamma/beta kernel
// (GammaBetaBackwardCUDAKernelTemplate: 191 us per call on 24 100 x 256 -- 19 % of a TF32 encoder step) and a separate
// column-sum kernel per Linear bias (13 %).  These four kernels are pure HBM streams, so unlike the gather they can run
// at the HBM roofline:
//   epilogue_ln_fwd      z = x + bias + residual;  y = LayerNorm(z) * gamma + beta      reads 2, writes 2 row-matrices
//   epilogue_ln_bwd      dz (= d x = d residual), d gamma, d beta, d bias                reads 2, writes 1
//   column_sum           d bias of a Linear: sum over rows                               reads 1
//   relu_bwd_column_sum  d pre-activation = d h * (h > 0) and its column sum             reads 2, writes 1
// One warp per row (C / 32 channels per lane as float4 chunks), row statistics by xor-shuffles, the per-column sums kept
// in registers across a persistent CTA's rows and flushed once per CTA (shared-memory reduce over its warps, then one
// red per column).  Grids are SMs x 4 CTAs of 8 warps.

template <int VEC>          // float4 chunks per lane: C = 128 * VEC
struct LaneRow {
    float4 v[VEC];
};

__device__ __forceinline__ void red_add_f32x4(float *p, const float4 &v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

template <int VEC>
__device__ __forceinline__ LaneRow<VEC> load_row(const float *p, int lane) {      // lane owns chunks lane, lane + 32, ...
    LaneRow<VEC> r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = __ldg(reinterpret_cast<const float4 *>(p) + lane + 32 * k);
    return r;
}
template <int VEC>
__device__ __forceinline__ LaneRow<VEC> load_row_stream(const float *p, int lane) {
    LaneRow<VEC> r;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(reinterpret_cast<const float4 *>(p) + lane + 32 * k));
        r.v[k] = v;
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void store_row(float *p, int lane, const LaneRow<VEC> &r) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) reinterpret_cast<float4 *>(p)[lane + 32 * k] = r.v[k];
}

// ---- dropout (deformable_transformer.py:226-235: dropout1 / dropout2 / dropout3, p = 0.1 in training) ----
// The keep mask is never stored: it is a pure function of (rng[0], rng[1], salt, element index) -- Philox4x32-7, one
// call per float4 chunk -- so the backward regenerates it.  `rng` points at two 64-bit words in device memory (drawn by
// the caller from torch's CUDA generator: reproducible under torch.manual_seed, and a captured CUDA graph sees fresh
// words on every replay); `salt` tells the call sites of one layer apart.  keep <=> word >= p * 2^32.
struct DropoutArgs {
    const uint64_t *rng;
    uint32_t salt;
    uint32_t thresh;     // floor(p * 2^32)
    float scale;         // 1 / (1 - p)
};
struct DropoutKey {
    uint2 key;
    uint32_t c2, c3;
};
__device__ __forceinline__ DropoutKey dropout_key(const DropoutArgs &da) {
    const uint64_t a = __ldg(da.rng), b = __ldg(da.rng + 1);
    return DropoutKey{make_uint2((uint32_t)a, (uint32_t)(a >> 32)), (uint32_t)b ^ da.salt, (uint32_t)(b >> 32)};
}
__device__ __forceinline__ uint4 philox4x32_7(uint2 key, uint4 c) {
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return c;
}
// keep * scale for the four elements of float4 chunk `chunk` (a global chunk index)
__device__ __forceinline__ float4 dropout_factor(const DropoutKey &k, const DropoutArgs &da, int64_t chunk) {
    const uint4 r = philox4x32_7(k.key, make_uint4((uint32_t)chunk, (uint32_t)((uint64_t)chunk >> 32), k.c2, k.c3));
    return make_float4(r.x >= da.thresh ? da.scale : 0.f, r.y >= da.thresh ? da.scale : 0.f,
                       r.z >= da.thresh ? da.scale : 0.f, r.w >= da.thresh ? da.scale : 0.f);
}

// y = LayerNorm(dropout(x + bias) + residual) * gamma + beta; also writes z (the pre-norm sum) and the row statistics.
template <int VEC, bool DROP>
__global__ void __launch_bounds__(256)
epilogue_ln_fwd(const float *__restrict__ x, const float *__restrict__ bias, const float *__restrict__ residual,
                const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int64_t rows,
                float *__restrict__ z_out, float *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                DropoutArgs da) {
    constexpr int C = 128 * VEC;
    DropoutKey dk{};
    if constexpr (DROP) dk = dropout_key(da);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const LaneRow<VEC> g = load_row<VEC>(gamma, lane), b = load_row<VEC>(beta, lane);
    LaneRow<VEC> bi;
#pragma unroll
    for (int k = 0; k < VEC; ++k) bi.v[k] = bias ? __ldg(reinterpret_cast<const float4 *>(bias) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = warp; r < rows; r += nwarps) {
        LaneRow<VEC> z = load_row_stream<VEC>(x + r * C, lane);
        const LaneRow<VEC> res = load_row_stream<VEC>(residual + r * C, lane);
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            if constexpr (DROP) {                                   // residual + dropout(x + bias)
                const float4 f = dropout_factor(dk, da, r * (C / 4) + lane + 32 * k);
                z.v[k].x = __fmul_rn(z.v[k].x + bi.v[k].x, f.x) + res.v[k].x;
                z.v[k].y = __fmul_rn(z.v[k].y + bi.v[k].y, f.y) + res.v[k].y;
                z.v[k].z = __fmul_rn(z.v[k].z + bi.v[k].z, f.z) + res.v[k].z;
                z.v[k].w = __fmul_rn(z.v[k].w + bi.v[k].w, f.w) + res.v[k].w;
            } else {
                z.v[k].x = (z.v[k].x + bi.v[k].x) + res.v[k].x;       // torch's order: (x + bias) + residual
                z.v[k].y = (z.v[k].y + bi.v[k].y) + res.v[k].y;
                z.v[k].z = (z.v[k].z + bi.v[k].z) + res.v[k].z;
                z.v[k].w = (z.v[k].w + bi.v[k].w) + res.v[k].w;
            }
            s += (z.v[k].x + z.v[k].y) + (z.v[k].z + z.v[k].w);
        }
        const float mean = warp_sum(s) * (1.f / C);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float dx = z.v[k].x - mean, dy = z.v[k].y - mean, dz = z.v[k].z - mean, dw = z.v[k].w - mean;
            q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
        LaneRow<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            o.v[k].x = (z.v[k].x - mean) * rstd * g.v[k].x + b.v[k].x;
            o.v[k].y = (z.v[k].y - mean) * rstd * g.v[k].y + b.v[k].y;
            o.v[k].z = (z.v[k].z - mean) * rstd * g.v[k].z + b.v[k].z;
            o.v[k].w = (z.v[k].w - mean) * rstd * g.v[k].w + b.v[k].w;
        }
        store_row<VEC>(z_out + r * C, lane, z);
        store_row<VEC>(y + r * C, lane, o);
        if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    }
}

// Flush a CTA's per-lane column partials: warps add into shared memory (one after the other), then one red per column.
template <int VEC, int NACC>
__device__ __forceinline__ void flush_column_sums(const LaneRow<VEC> (&acc)[NACC], float *const (&out)[NACC], float *smem) {
    constexpr int C = 128 * VEC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < NACC * C; i += blockDim.x) smem[i] = 0.f;
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
        if (w == warp) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float4 *p = reinterpret_cast<float4 *>(smem + a * C) + lane + 32 * k;
                    float4 t = *p;
                    t.x += acc[a].v[k].x; t.y += acc[a].v[k].y; t.z += acc[a].v[k].z; t.w += acc[a].v[k].w;
                    *p = t;
                }
        }
        __syncthreads();
    }
    // one 16-byte red per column chunk: every CTA of the grid adds into the same C addresses, and the L2 serialises
    // same-address atomics, so their number (not their bytes) is what this flush costs
    for (int i = threadIdx.x; i < NACC * (C / 4); i += blockDim.x) {
        float *dst = out[i / (C / 4)];
        if (dst) red_add_f32x4(dst + 4 * (i % (C / 4)), reinterpret_cast<const float4 *>(smem)[i]);
    }
}

// dz = rstd * (g - mean(g) - zh * mean(g * zh)),  g = dy * gamma, zh = (z - mean) * rstd;  d gamma += dy * zh,
// d beta += dy, d bias += dz (outputs zero-filled by the launcher).
// With DROP: dz is the residual's gradient, dx_out = dz * keep / (1 - p) the Linear output's, and d bias sums dx.
template <int VEC, bool DROP>
__global__ void __launch_bounds__(256)
epilogue_ln_bwd(const float *__restrict__ dy, const float *__restrict__ z, const float *__restrict__ mean_in,
                const float *__restrict__ rstd_in, const float *__restrict__ gamma, int64_t rows, float *__restrict__ dz_out,
                float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ dbias, float *__restrict__ dx_out,
                DropoutArgs da) {
    constexpr int C = 128 * VEC;
    DropoutKey dk{};
    if constexpr (DROP) dk = dropout_key(da);
    __shared__ __align__(16) float s_cols[3 * C];
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const LaneRow<VEC> g = load_row<VEC>(gamma, lane);
    LaneRow<VEC> acc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[a].v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = warp; r < rows; r += nwarps) {
        const LaneRow<VEC> d = load_row_stream<VEC>(dy + r * C, lane);
        LaneRow<VEC> zh = load_row_stream<VEC>(z + r * C, lane);
        const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
        float s1 = 0.f, s2 = 0.f;
        LaneRow<VEC> gg;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            zh.v[k].x = (zh.v[k].x - mean) * rstd; zh.v[k].y = (zh.v[k].y - mean) * rstd;
            zh.v[k].z = (zh.v[k].z - mean) * rstd; zh.v[k].w = (zh.v[k].w - mean) * rstd;
            gg.v[k].x = d.v[k].x * g.v[k].x; gg.v[k].y = d.v[k].y * g.v[k].y;
            gg.v[k].z = d.v[k].z * g.v[k].z; gg.v[k].w = d.v[k].w * g.v[k].w;
            s1 += (gg.v[k].x + gg.v[k].y) + (gg.v[k].z + gg.v[k].w);
            s2 += (gg.v[k].x * zh.v[k].x + gg.v[k].y * zh.v[k].y) + (gg.v[k].z * zh.v[k].z + gg.v[k].w * zh.v[k].w);
        }
        const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
        LaneRow<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            o.v[k].x = rstd * (gg.v[k].x - m1 - zh.v[k].x * m2); o.v[k].y = rstd * (gg.v[k].y - m1 - zh.v[k].y * m2);
            o.v[k].z = rstd * (gg.v[k].z - m1 - zh.v[k].z * m2); o.v[k].w = rstd * (gg.v[k].w - m1 - zh.v[k].w * m2);
            acc[0].v[k].x += d.v[k].x * zh.v[k].x; acc[0].v[k].y += d.v[k].y * zh.v[k].y;
            acc[0].v[k].z += d.v[k].z * zh.v[k].z; acc[0].v[k].w += d.v[k].w * zh.v[k].w;
            acc[1].v[k].x += d.v[k].x; acc[1].v[k].y += d.v[k].y; acc[1].v[k].z += d.v[k].z; acc[1].v[k].w += d.v[k].w;
            if constexpr (DROP) {
                const float4 f = dropout_factor(dk, da, r * (C / 4) + lane + 32 * k);
                const float4 dx = make_float4(o.v[k].x * f.x, o.v[k].y * f.y, o.v[k].z * f.z, o.v[k].w * f.w);
                reinterpret_cast<float4 *>(dx_out + r * C)[lane + 32 * k] = dx;
                acc[2].v[k].x += dx.x; acc[2].v[k].y += dx.y; acc[2].v[k].z += dx.z; acc[2].v[k].w += dx.w;
            } else {
                acc[2].v[k].x += o.v[k].x; acc[2].v[k].y += o.v[k].y; acc[2].v[k].z += o.v[k].z; acc[2].v[k].w += o.v[k].w;
            }
        }
        store_row<VEC>(dz_out + r * C, lane, o);
    }
    float *const outs[3] = {dgamma, dbeta, dbias};
    flush_column_sums<VEC, 3>(acc, outs, s_cols);
}

// out[c] += sum_r x[r][c]; with `mask_src` (the ReLU output) also dpre[r][c] = scale * x[r][c] * (mask_src[r][c] > 0), summed.
// Any C that is a multiple of 4: a CTA's threads tile (rows x column chunks); out is zero-filled by the launcher.
template <bool RELU>
__global__ void __launch_bounds__(256)
column_sum_kernel(const float *__restrict__ x, const float *__restrict__ mask_src, int64_t rows, int C, float *__restrict__ dpre,
                  float *__restrict__ out, float scale) {
    float4 *s_part = reinterpret_cast<float4 *>(msda_smem);          // [rows_per_pass][cols_per_pass]
    const int C4 = C >> 2;
    const int cols = C4 < 256 ? C4 : 256;                             // column chunks a CTA covers at a time
    const int rpp = 256 / cols;                                       // rows per pass (>= 1)
    const int tx = threadIdx.x % cols, ty = threadIdx.x / cols;
    const bool active = ty < rpp;
    // CTA b takes the contiguous row block [r0, r1)
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    for (int c0 = 0; c0 < C4; c0 += cols) {
        const int c = c0 + tx;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && c < C4) {
            auto take = [&](int64_t r) -> float4 {
                float4 v;
                const float4 *src = reinterpret_cast<const float4 *>(x) + r * C4 + c;
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src));
                if constexpr (RELU) {
                    float4 h;
                    const float4 *hs = reinterpret_cast<const float4 *>(mask_src) + r * C4 + c;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(h.x), "=f"(h.y), "=f"(h.z), "=f"(h.w) : "l"(hs));
                    // `scale` = 1 / (1 - p) when h is the DROPPED ReLU output (h > 0 <=> positive and kept), else 1
                    v.x = h.x > 0.f ? v.x * scale : 0.f; v.y = h.y > 0.f ? v.y * scale : 0.f;
                    v.z = h.z > 0.f ? v.z * scale : 0.f; v.w = h.w > 0.f ? v.w * scale : 0.f;
                    reinterpret_cast<float4 *>(dpre)[r * C4 + c] = v;
                }
                return v;
            };
            // kU rows in flight per thread: a single 16-byte load per iteration leaves the kernel latency-bound
            constexpr int kU = RELU ? 4 : 8;
            int64_t r = r0 + ty;
            for (; r + (kU - 1) * (int64_t)rpp < r1; r += kU * (int64_t)rpp) {
                float4 a[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) a[u] = take(r + u * (int64_t)rpp);
#pragma unroll
                for (int u = 0; u < kU; ++u) { acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w; }
            }
            for (; r < r1; r += rpp) {
                const float4 a0 = take(r);
                acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
            }
        }
        if (active) s_part[ty * cols + tx] = acc;
        __syncthreads();
        if (ty == 0 && c < C4) {
            for (int k = 1; k < rpp; ++k) {
                const float4 p = s_part[k * cols + tx];
                acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
            }
            if (r0 < r1) red_add_f32x4(out + 4 * c, acc);
        }
        __syncthreads();
    }
}

// h *= keep / (1 - p) in place (dropout2, after the FFN's ReLU: deformable_transformer.py:244).  Because the result is
// positive exactly where the ReLU output was positive AND kept, the backward needs no mask: relu_bwd with `scale`.
__global__ void __launch_bounds__(256)
dropout_inplace_kernel(float4 *__restrict__ h, int64_t n4, DropoutArgs da) {
    const DropoutKey dk = dropout_key(da);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = h[i];
        const float4 f = dropout_factor(dk, da, i);
        v.x *= f.x; v.y *= f.y; v.z *= f.z; v.w *= f.w;
        h[i] = v;
    }
}
// the keep mask itself, one byte per element (tests: the kernels never materialise it)
__global__ void __launch_bounds__(256)
dropout_mask_kernel(uchar4 *__restrict__ out, int64_t n4, DropoutArgs da) {
    const DropoutKey dk = dropout_key(da);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 f = dropout_factor(dk, da, i);
        out[i] = make_uchar4(f.x != 0.f, f.y != 0.f, f.z != 0.f, f.w != 0.f);
    }
}
