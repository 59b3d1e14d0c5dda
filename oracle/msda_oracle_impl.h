/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle body, included once per precision by msda_oracle.c.
 *
 * Restates, for the host CPU, the arithmetic of the reference multi-scale deformable attention op.
 * All citations are relative to
 *   /root/reference/projects/UNINEXT/uninext/models/deformable_detr/ops/src/cuda/ms_deform_im2col_cuda.cuh
 * ("cuh") unless stated otherwise.
 *
 * Expected macros: REAL (float|double), FN(name) (symbol suffixing).
 *
 * Memory layout (ms_deform_attn_cuda.cu:40-48,57-60):
 *   value   [N, S, M, D]          spatial_shapes [L, 2] = (H_l, W_l)   level_start_index [L]
 *   loc     [N, Lq, M, L, P, 2]   last dim (x, y), normalised to [0,1] of the level map
 *   attn    [N, Lq, M, L, P]      out [N, Lq, M*D]
 */

/* One bilinear tap. Mirrors the per-corner predicates of cuh:47-78 (forward) and cuh:114-158 (backward).
 * v[k] receives the address offset (in rows of D) of corner k or -1 when that corner is outside the map. */
static inline void FN(corners)(REAL h_im, REAL w_im, int H, int W,
                               long long rows[4], REAL cw[4], REAL *lh_o, REAL *lw_o)
{
    const int h0 = (int)floor((double)h_im);   /* cuh:38-41 */
    const int w0 = (int)floor((double)w_im);
    const int h1 = h0 + 1, w1 = w0 + 1;
    const REAL lh = h_im - (REAL)h0, lw = w_im - (REAL)w0;   /* cuh:43-44 */
    const REAL hh = (REAL)1 - lh, hw = (REAL)1 - lw;         /* cuh:45 */
    rows[0] = (h0 >= 0 && w0 >= 0)         ? (long long)h0 * W + w0 : -1;   /* cuh:56-60 */
    rows[1] = (h0 >= 0 && w1 <= W - 1)     ? (long long)h0 * W + w1 : -1;   /* cuh:62-66 */
    rows[2] = (h1 <= H - 1 && w0 >= 0)     ? (long long)h1 * W + w0 : -1;   /* cuh:68-72 */
    rows[3] = (h1 <= H - 1 && w1 <= W - 1) ? (long long)h1 * W + w1 : -1;   /* cuh:74-78 */
    cw[0] = hh * hw; cw[1] = hh * lw; cw[2] = lh * hw; cw[3] = lh * lw;     /* cuh:80 */
    *lh_o = lh; *lw_o = lw;
}

/* Forward: restates ms_deformable_im2col_gpu_kernel (cuh:237-299). */
void FN(msda_oracle_forward)(const REAL *value, const int64_t *shapes, const int64_t *lsi,
                             const REAL *loc, const REAL *attn,
                             int N, int S, int M, int D, int L, int Lq, int P, REAL *out)
{
    const long long rowstride = (long long)M * D;            /* cuh:47 w_stride */
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < N; ++b) {
        for (int q = 0; q < Lq; ++q) {
            for (int m = 0; m < M; ++m) {
                const long long qm = ((long long)b * Lq + q) * M + m;      /* cuh:258 sampling_index */
                REAL *o = out + qm * D;
                for (int c = 0; c < D; ++c) o[c] = (REAL)0;
                for (int l = 0; l < L; ++l) {
                    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];   /* cuh:274-277 */
                    const REAL *vl = value + ((long long)b * S + lsi[l]) * rowstride + (long long)m * D;
                    for (int p = 0; p < P; ++p) {
                        const long long sidx = (qm * L + l) * P + p;
                        const REAL x = loc[2 * sidx], y = loc[2 * sidx + 1], a = attn[sidx];
                        const REAL h_im = y * (REAL)H - (REAL)0.5;               /* cuh:285 */
                        const REAL w_im = x * (REAL)W - (REAL)0.5;               /* cuh:286 */
                        if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;   /* cuh:288 */
                        long long rows[4]; REAL cw[4], lh, lw;
                        FN(corners)(h_im, w_im, H, W, rows, cw, &lh, &lw);
                        for (int c = 0; c < D; ++c) {
                            REAL v[4];
                            for (int k = 0; k < 4; ++k) v[k] = rows[k] >= 0 ? vl[rows[k] * rowstride + c] : (REAL)0;
                            const REAL val = cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3]; /* cuh:82 */
                            o[c] += val * a;                                     /* cuh:290 */
                        }
                    }
                }
            }
        }
    }
}

/* Backward: restates ms_deform_attn_col2im_bilinear (cuh:87-159) and the channel reduction of the col2im
 * kernels (cuh:301-403 and variants). grad_value must NOT be pre-zeroed by the caller: it is zero-filled here,
 * as the reference host wrapper does with at::zeros_like (ms_deform_attn_cuda.cu:121-123).
 * Parallel over (b, m): each (b, m) pair owns a disjoint channel slice of grad_value, so no atomics are needed
 * and the summation order over (q, l, p) is fixed (deterministic, unlike the reference's atomicAdd order). */
void FN(msda_oracle_backward)(const REAL *grad_out, const REAL *value, const int64_t *shapes, const int64_t *lsi,
                              const REAL *loc, const REAL *attn,
                              int N, int S, int M, int D, int L, int Lq, int P,
                              REAL *grad_value, REAL *grad_loc, REAL *grad_attn)
{
    const long long rowstride = (long long)M * D;
    memset(grad_value, 0, sizeof(REAL) * (size_t)N * S * M * D);
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < N; ++b) {
        for (int m = 0; m < M; ++m) {
            for (int q = 0; q < Lq; ++q) {
                const long long qm = ((long long)b * Lq + q) * M + m;
                const REAL *g = grad_out + qm * D;
                for (int l = 0; l < L; ++l) {
                    const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
                    const long long base = ((long long)b * S + lsi[l]) * rowstride + (long long)m * D;
                    const REAL *vl = value + base;
                    REAL *gvl = grad_value + base;
                    for (int p = 0; p < P; ++p) {
                        const long long sidx = (qm * L + l) * P + p;
                        const REAL x = loc[2 * sidx], y = loc[2 * sidx + 1], a = attn[sidx];
                        const REAL h_im = y * (REAL)H - (REAL)0.5;
                        const REAL w_im = x * (REAL)W - (REAL)0.5;
                        REAL ga = 0, gx = 0, gy = 0;
                        if (h_im > -1 && w_im > -1 && h_im < H && w_im < W) {        /* cuh:369 */
                            long long rows[4]; REAL cw[4], lh, lw;
                            FN(corners)(h_im, w_im, H, W, rows, cw, &lh, &lw);
                            const REAL hh = (REAL)1 - lh, hw = (REAL)1 - lw;
                            for (int c = 0; c < D; ++c) {
                                const REAL tg = g[c] * a;                             /* cuh:112 top_grad_value */
                                REAL v[4];
                                for (int k = 0; k < 4; ++k) {
                                    v[k] = (REAL)0;
                                    if (rows[k] >= 0) {
                                        v[k] = vl[rows[k] * rowstride + c];
                                        gvl[rows[k] * rowstride + c] += cw[k] * tg;   /* cuh:125,134,143,152 */
                                    }
                                }
                                /* cuh:123-124,132-133,141-142,150-151 */
                                const REAL gh = -hw * v[0] - lw * v[1] + hw * v[2] + lw * v[3];
                                const REAL gw = -hh * v[0] + hh * v[1] - lh * v[2] + lh * v[3];
                                const REAL val = cw[0] * v[0] + cw[1] * v[1] + cw[2] * v[2] + cw[3] * v[3];
                                ga += g[c] * val;                                     /* cuh:156 */
                                gx += (REAL)W * gw * tg;                              /* cuh:157 */
                                gy += (REAL)H * gh * tg;                              /* cuh:158 */
                            }
                        }
                        grad_attn[sidx] = ga;
                        grad_loc[2 * sidx] = gx;
                        grad_loc[2 * sidx + 1] = gy;
                    }
                }
            }
        }
    }
}
