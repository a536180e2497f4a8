/*
 * msda_oracle.c -- CPU restatement of the reference MSDeformAttn arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ocpg_b200/ (the product) may call, link or import
 * this file; it is used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference leg as the checker / reported baseline.
 *
 * What it follows (all paths relative to /root/reference/models/ops/src/cuda/):
 *   - bilinear sample with zero padding ........ ms_deform_im2col_cuda.cuh:33-84
 *   - gradients of one sample point ............ ms_deform_im2col_cuda.cuh:87-159
 *   - forward indexing / point loop ............ ms_deform_im2col_cuda.cuh:237-299
 *   - backward indexing / reductions ........... ms_deform_im2col_cuda.cuh:301-403
 *   - tensor shape conventions ................. ms_deform_attn_cuda.cu:40-60
 *
 * Parity pin: tests/test_oracle.py checks this file against tests/golden/ (npz files), which were produced
 * by importing the reference's own ms_deform_attn_core_pytorch (functions/ms_deform_attn_func.py:41-61)
 * in the build container (generator: tests/golden/make_golden.py), on the reference's own test
 * geometry (test.py:21-28) and on out-of-range / ragged geometries the reference never tests.
 *
 * Layouts (row-major contiguous):
 *   value  [N][S][M][D]      shapes [L][2] = (H_l, W_l) int64     start [L] int64
 *   loc    [N][Lq][M][L][P][2] = (x, y) normalised to [0,1]       attn [N][Lq][M][L][P]
 *   out / grad_out [N][Lq][M*D]
 *
 * The per-point arithmetic is kept in the reference's evaluation order and in the tensor's own
 * scalar type (so the f32 instantiation is a model of the reference CUDA kernel's rounding), the f64
 * instantiation is the ground truth used for tolerances.  The pixel coordinate loc*size - 0.5 is a
 * fused multiply-add: that is what nvcc makes of cuh:285-286 (SASS of the reference op built for
 * sm_100a: `FFMA R29, R12, R29, -0.5`), so the cell a borderline point falls into is the reference's.
 * Build: see oracle/Makefile (-ffp-contract=off keeps the f32 instantiation deterministic).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MSDA_ORACLE_DEFINE(SUFFIX, T, FLOOR, FMA)                                                       \
                                                                                                    \
/* One sample point: the four corner indices (or -1), and the four bilinear weights.             */ \
/* cuh:38-53 for the cell decomposition, cuh:56/62/68/74 for the per-corner validity tests.      */ \
static inline void cell_##SUFFIX(T y, T x, int H, int W, long corner[4], T wgt[4], T frac[4])       \
{                                                                                                   \
    const int y0 = (int)FLOOR(y), x0 = (int)FLOOR(x);                                               \
    const int y1 = y0 + 1, x1 = x0 + 1;                                                             \
    const T ly = y - (T)y0, lx = x - (T)x0;                                                         \
    const T hy = (T)1 - ly, hx = (T)1 - lx;                                                         \
    corner[0] = (y0 >= 0 && x0 >= 0) ? (long)y0 * W + x0 : -1;                                      \
    corner[1] = (y0 >= 0 && x1 <= W - 1) ? (long)y0 * W + x1 : -1;                                  \
    corner[2] = (y1 <= H - 1 && x0 >= 0) ? (long)y1 * W + x0 : -1;                                  \
    corner[3] = (y1 <= H - 1 && x1 <= W - 1) ? (long)y1 * W + x1 : -1;                              \
    wgt[0] = hy * hx; wgt[1] = hy * lx; wgt[2] = ly * hx; wgt[3] = ly * lx;                         \
    frac[0] = ly; frac[1] = lx; frac[2] = hy; frac[3] = hx;                                         \
}                                                                                                   \
                                                                                                    \
void msda_oracle_forward_##SUFFIX(const T *value, const int64_t *shapes, const int64_t *start,      \
                                  const T *loc, const T *attn, int N, int S, int M, int D, int L,   \
                                  int Lq, int P, T *out)                                            \
{                                                                                                   \
    const long rows = (long)N * Lq * M;                                                             \
    _Pragma("omp parallel for schedule(static)")                                                    \
    for (long r = 0; r < rows; ++r) {                                                               \
        const int m = (int)(r % M);                                                                 \
        const long n = r / ((long)Lq * M);                                                          \
        const T *pl = loc + r * L * P * 2;                                                          \
        const T *pa = attn + r * L * P;                                                             \
        T *po = out + r * D;                                                                        \
        for (int c = 0; c < D; ++c) po[c] = (T)0;                                                   \
        for (int l = 0; l < L; ++l) {                                                               \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                           \
            const T *vl = value + ((n * S + start[l]) * M + m) * D;   /* cuh:269,278 */             \
            for (int p = 0; p < P; ++p, pl += 2, ++pa) {                                            \
                const T y = FMA(pl[1], (T)H, (T)-0.5);                /* cuh:285 */                 \
                const T x = FMA(pl[0], (T)W, (T)-0.5);                /* cuh:286 */                 \
                if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) continue; /* cuh:288 */      \
                long cr[4]; T w[4], fr[4];                                                          \
                cell_##SUFFIX(y, x, H, W, cr, w, fr);                                               \
                for (int c = 0; c < D; ++c) {                                                       \
                    T v[4];                                                                         \
                    for (int i = 0; i < 4; ++i) v[i] = cr[i] >= 0 ? vl[cr[i] * M * D + c] : (T)0;   \
                    const T b = w[0] * v[0] + w[1] * v[1] + w[2] * v[2] + w[3] * v[3]; /* :80-82 */ \
                    po[c] += b * pa[0];                                               /* :290 */    \
                }                                                                                   \
            }                                                                                       \
        }                                                                                           \
    }                                                                                               \
}                                                                                                   \
                                                                                                    \
/* grad_value is accumulated serially per frame (frames are independent, cuh:353-355), so the   */  \
/* oracle's grad_value is deterministic, unlike the reference's atomicAdd order.                */  \
void msda_oracle_backward_##SUFFIX(const T *value, const int64_t *shapes, const int64_t *start,     \
                                   const T *loc, const T *attn, const T *grad_out, int N, int S,    \
                                   int M, int D, int L, int Lq, int P, T *grad_value, T *grad_loc,  \
                                   T *grad_attn)                                                    \
{                                                                                                   \
    memset(grad_value, 0, sizeof(T) * (size_t)N * S * M * D);         /* cu:121 */                  \
    memset(grad_loc, 0, sizeof(T) * (size_t)N * Lq * M * L * P * 2);  /* cu:122 */                  \
    memset(grad_attn, 0, sizeof(T) * (size_t)N * Lq * M * L * P);     /* cu:123 */                  \
    _Pragma("omp parallel for schedule(static)")                                                    \
    for (long nm = 0; nm < (long)N * M; ++nm) {                                                     \
        const long n = nm / M;                                                                      \
        const int m = (int)(nm % M);                                                                \
        for (int q = 0; q < Lq; ++q) {                                                              \
            const long r = (n * Lq + q) * M + m;                                                    \
            const T *pl = loc + r * L * P * 2;                                                      \
            const T *pa = attn + r * L * P;                                                         \
            const T *pg = grad_out + r * D;                                                         \
            T *gl = grad_loc + r * L * P * 2;                                                       \
            T *ga = grad_attn + r * L * P;                                                          \
            for (int l = 0; l < L; ++l) {                                                           \
                const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                       \
                const long base = ((n * S + start[l]) * M + m) * D;                                 \
                for (int p = 0; p < P; ++p, pl += 2, ++pa, gl += 2, ++ga) {                         \
                    const T y = FMA(pl[1], (T)H, (T)-0.5);                                          \
                    const T x = FMA(pl[0], (T)W, (T)-0.5);                                          \
                    if (!(y > (T)-1 && x > (T)-1 && y < (T)H && x < (T)W)) continue; /* :365 */     \
                    long cr[4]; T w[4], fr[4];                                                      \
                    cell_##SUFFIX(y, x, H, W, cr, w, fr);                                           \
                    const T ly = fr[0], lx = fr[1], hy = fr[2], hx = fr[3];                         \
                    T sum_a = (T)0, sum_x = (T)0, sum_y = (T)0;                                     \
                    for (int c = 0; c < D; ++c) {                                                   \
                        const T g = pg[c];                                                          \
                        const T ga_c = g * pa[0];                     /* top_grad_value, :115 */    \
                        T v[4];                                                                     \
                        for (int i = 0; i < 4; ++i) {                                               \
                            v[i] = (T)0;                                                            \
                            if (cr[i] >= 0) {                                                       \
                                const long o = base + cr[i] * M * D + c;                            \
                                v[i] = value[o];                                                    \
                                grad_value[o] += w[i] * ga_c;         /* :125,134,143,152 */        \
                            }                                                                       \
                        }                                                                           \
                        T gh = (T)0, gw = (T)0;                       /* :116, :123-151 */          \
                        gh -= hx * v[0]; gw -= hy * v[0];                                           \
                        gh -= lx * v[1]; gw += hy * v[1];                                           \
                        gh += hx * v[2]; gw -= ly * v[2];                                           \
                        gh += lx * v[3]; gw += ly * v[3];                                           \
                        const T b = w[0] * v[0] + w[1] * v[1] + w[2] * v[2] + w[3] * v[3];          \
                        sum_a += g * b;                               /* :156 */                    \
                        sum_x += (T)W * gw * ga_c;                    /* :157 */                    \
                        sum_y += (T)H * gh * ga_c;                    /* :158 */                    \
                    }                                                                               \
                    gl[0] = sum_x; gl[1] = sum_y; ga[0] = sum_a;      /* :377-393 */                \
                }                                                                                   \
            }                                                                                       \
        }                                                                                           \
    }                                                                                               \
}

MSDA_ORACLE_DEFINE(f32, float, floorf, fmaf)
MSDA_ORACLE_DEFINE(f64, double, floor, fma)

int msda_oracle_abi_version(void) { return 1; }
