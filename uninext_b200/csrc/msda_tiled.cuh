// msda_tiled.cuh -- the sm_100a fast path: row-vectorised gather kernels (forward and backward).
//
// Thread mapping (both kernels).  A value row of one head is D contiguous elements; LPR = D / kElems lanes cover it
// with one 16-byte access each (fp32 D=32: 8 lanes, bf16 D=32: 4 lanes).  A warp therefore carries GPW = 32 / LPR
// "groups"; each group owns one (batch, query, head) pair at a time.  Work inside a group is split two ways:
//   stage 1 (by tap):     lane `sub` resolves taps sub, sub+LPR, ... : reads (x, y, a), resolves the bilinear
//                         geometry once (tap_geometry), keeps it in registers.
//   stage 2 (by channel): for every tap the owning lane broadcasts 4 masked corner weights + 2 row indices with
//                         group-wide shuffles; every lane then issues four 16-byte row loads for ITS channel slice.
// This removes the reference forward kernel's 32x-redundant per-channel index arithmetic and scalar loads
// (cuh:272-296) and the reference backward kernel's per-tap __syncthreads + serial shared-memory reductions
// (cuh:347-401): the backward channel reduction is a shuffle reduce-scatter that leaves tap j's sums on lane j,
// i.e. on the lane that already holds that tap's geometry.
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kTiledThreads = 256;

template <int LPR>
__device__ __forceinline__ float group_bcast(float v, int src) { return __shfl_sync(kFullMask, v, src, LPR); }
template <int LPR>
__device__ __forceinline__ int group_bcast(int v, int src) { return __shfl_sync(kFullMask, v, src, LPR); }

struct LevelSmem {
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
};

__device__ __forceinline__ void load_levels(LevelSmem &lv, const int64_t *shapes, const int64_t *lsi, int L) {
    if (threadIdx.x < L) {
        lv.H[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        lv.W[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        lv.start[threadIdx.x] = (int)lsi[threadIdx.x];
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// forward:  out[b,q,m,:] = sum_taps a * bilinear(value_l[b,:,m,:], x, y)            (reference cuh:237-299)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int D, int LP_MAX>
__global__ void __launch_bounds__(kTiledThreads)
msda_fwd_tiled(const T *__restrict__ value, const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ loc, const float *__restrict__ attn,
               int S, int M, int L, int Lq, int P, long long npairs, int pairs_per_cta, T *__restrict__ out)
{
    constexpr int VEC = RowVec<T>::kElems;
    constexpr int LPR = D / VEC;            // lanes per row
    constexpr int GPW = 32 / LPR;           // (b,q,m) pairs in flight per warp
    constexpr int NSL = LP_MAX / LPR;       // taps resolved per lane
    static_assert(D % VEC == 0 && (LPR & (LPR - 1)) == 0 && LPR <= 32 && LP_MAX % LPR == 0, "bad tiling");

    __shared__ LevelSmem lv;
    load_levels(lv, shapes, lsi, L);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane % LPR, grp = lane / LPR;
    const int LP = L * P;
    const size_t row_elems = (size_t)M * D;
    const long long cta_begin = (long long)blockIdx.x * pairs_per_cta;
    const long long cta_end = min(npairs, cta_begin + (long long)pairs_per_cta);

    for (long long p0 = cta_begin + (long long)warp * GPW; p0 < cta_end; p0 += (long long)nwarps * GPW) {
        const long long pair_raw = p0 + grp;
        const bool active = pair_raw < cta_end;
        const long long pair = active ? pair_raw : cta_end - 1;          // keep idle groups on legal addresses
        const int m = (int)(pair % M);
        const int b = (int)((pair / M) / Lq);

        // ---- stage 1: this lane resolves its taps ----
        float tw[NSL][4];
        int tr0[NSL], tr1[NSL];
#pragma unroll
        for (int k = 0; k < NSL; ++k) {
            const int s = sub + k * LPR;
            tw[k][0] = tw[k][1] = tw[k][2] = tw[k][3] = 0.f;
            tr0[k] = 0; tr1[k] = 0;
            if (s < LP && active) {
                const long long t = pair * LP + s;
                const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + t);
                const float a = __ldg(attn + t);
                const int l = s / P;
                const TapGeom g = tap_geometry(xy.x, xy.y, lv.H[l], lv.W[l], lv.start[l]);
                const float hh = 1.f - g.lh, hw = 1.f - g.lw;
                tw[k][0] = (g.mask & 1u) ? hh * hw * a : 0.f;
                tw[k][1] = (g.mask & 2u) ? hh * g.lw * a : 0.f;
                tw[k][2] = (g.mask & 4u) ? g.lh * hw * a : 0.f;
                tw[k][3] = (g.mask & 8u) ? g.lh * g.lw * a : 0.f;
                tr0[k] = g.r0;
                tr1[k] = g.r1 | (g.dw << 31);
            }
        }

        // ---- stage 2: gather rows for this lane's channel slice ----
        const T *base = value + (size_t)b * S * row_elems + (size_t)m * D + (size_t)sub * VEC;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int k = 0; k < NSL; ++k) {
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const float w00 = group_bcast<LPR>(tw[k][0], j);
                const float w01 = group_bcast<LPR>(tw[k][1], j);
                const float w10 = group_bcast<LPR>(tw[k][2], j);
                const float w11 = group_bcast<LPR>(tw[k][3], j);
                const int r0 = group_bcast<LPR>(tr0[k], j);
                const int r1x = group_bcast<LPR>(tr1[k], j);
                const size_t dwo = (r1x < 0) ? row_elems : 0;
                const T *p00 = base + (size_t)r0 * row_elems;
                const T *p10 = base + (size_t)(r1x & 0x7fffffff) * row_elems;
                float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
                RowVec<T>::load(p00, v00);
                RowVec<T>::load(p00 + dwo, v01);
                RowVec<T>::load(p10, v10);
                RowVec<T>::load(p10 + dwo, v11);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    acc[e] = fmaf(w00, v00[e], acc[e]);
                    acc[e] = fmaf(w01, v01[e], acc[e]);
                    acc[e] = fmaf(w10, v10[e], acc[e]);
                    acc[e] = fmaf(w11, v11[e], acc[e]);
                }
            }
        }
        if (active) RowVec<T>::store(out + (size_t)pair * D + (size_t)sub * VEC, acc);
    }
}

// Shuffle reduce-scatter inside a group of LPR lanes: on entry part[j][c] is this lane's partial sum for tap j,
// corner c; on exit res[c] is the full group sum for tap `sub`.  LPR-1 rounds of 4*LPR/2^r shuffles.
template <int LPR>
__device__ __forceinline__ void group_reduce_scatter(float (&part)[LPR][4], int sub, float (&res)[4]) {
#pragma unroll
    for (int d = LPR / 2; d >= 1; d /= 2) {
        const bool upper = (sub & d) != 0;
#pragma unroll
        for (int j = 0; j < d; ++j) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float lo = part[j][c], hi = part[j + d][c];
                const float send = upper ? lo : hi;
                const float keep = upper ? hi : lo;
                part[j][c] = keep + __shfl_xor_sync(kFullMask, send, d);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) res[c] = part[0][c];
}

// ------------------------------------------------------------------------------------------------------------
// backward (reference cuh:87-159 + cuh:301-403):
//   grad_value[corner rows] += w_corner * a * g            (16-byte vector reductions, fp32 accumulator)
//   grad_attn[tap]  = sum_c g[c] * val[c]
//   grad_loc[tap].x = W_l * a * sum_c g[c] * gw[c] ;  .y = H_l * a * sum_c g[c] * gh[c]
// Per tap only the four corner dot products  dot_k = sum_c g[c] * V_k[c]  cross lanes; the bilinear coefficients are
// applied afterwards by the single lane that owns the tap.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int D, int LP_MAX>
__global__ void __launch_bounds__(kTiledThreads)
msda_bwd_tiled(const T *__restrict__ grad_out, const T *__restrict__ value,
               const int64_t *__restrict__ shapes, const int64_t *__restrict__ lsi,
               const float *__restrict__ loc, const float *__restrict__ attn,
               int S, int M, int L, int Lq, int P, long long npairs, int pairs_per_cta,
               float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn)
{
    constexpr int VEC = RowVec<T>::kElems;
    constexpr int LPR = D / VEC;
    constexpr int GPW = 32 / LPR;
    constexpr int NSL = LP_MAX / LPR;
    static_assert(D % VEC == 0 && (LPR & (LPR - 1)) == 0 && LPR <= 32 && LP_MAX % LPR == 0, "bad tiling");

    __shared__ LevelSmem lv;
    load_levels(lv, shapes, lsi, L);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int sub = lane % LPR, grp = lane / LPR;
    const int LP = L * P;
    const size_t row_elems = (size_t)M * D;
    const long long cta_begin = (long long)blockIdx.x * pairs_per_cta;
    const long long cta_end = min(npairs, cta_begin + (long long)pairs_per_cta);

    for (long long p0 = cta_begin + (long long)warp * GPW; p0 < cta_end; p0 += (long long)nwarps * GPW) {
        const long long pair_raw = p0 + grp;
        const bool active = pair_raw < cta_end;
        const long long pair = active ? pair_raw : cta_end - 1;
        const int m = (int)(pair % M);
        const int b = (int)((pair / M) / Lq);

        float g[VEC];
        RowVec<T>::load(grad_out + (size_t)pair * D + (size_t)sub * VEC, g);

        // ---- stage 1 ----
        float tw[NSL][4], tlh[NSL], tlw[NSL], ta[NSL];
        int tr0[NSL], tr1[NSL], tl[NSL];
        unsigned tm[NSL];
#pragma unroll
        for (int k = 0; k < NSL; ++k) {
            const int s = sub + k * LPR;
            tw[k][0] = tw[k][1] = tw[k][2] = tw[k][3] = 0.f;
            tr0[k] = 0; tr1[k] = 0; tl[k] = 0; tm[k] = 0; tlh[k] = 0.f; tlw[k] = 0.f; ta[k] = 0.f;
            if (s < LP && active) {
                const long long t = pair * LP + s;
                const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + t);
                const float a = __ldg(attn + t);
                const int l = s / P;
                const TapGeom gm = tap_geometry(xy.x, xy.y, lv.H[l], lv.W[l], lv.start[l]);
                const float hh = 1.f - gm.lh, hw = 1.f - gm.lw;
                tw[k][0] = (gm.mask & 1u) ? hh * hw * a : 0.f;
                tw[k][1] = (gm.mask & 2u) ? hh * gm.lw * a : 0.f;
                tw[k][2] = (gm.mask & 4u) ? gm.lh * hw * a : 0.f;
                tw[k][3] = (gm.mask & 8u) ? gm.lh * gm.lw * a : 0.f;
                tr0[k] = gm.r0;
                tr1[k] = gm.r1 | (gm.dw << 31);
                tl[k] = l; tm[k] = gm.mask; tlh[k] = gm.lh; tlw[k] = gm.lw; ta[k] = a;
            }
        }

        const size_t slab = (size_t)b * S * row_elems + (size_t)m * D + (size_t)sub * VEC;
        const T *base = value + slab;
        float *gbase = grad_value + slab;

        // ---- stage 2 ----
#pragma unroll
        for (int k = 0; k < NSL; ++k) {
            float part[LPR][4];
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                float w[4];
                w[0] = group_bcast<LPR>(tw[k][0], j);
                w[1] = group_bcast<LPR>(tw[k][1], j);
                w[2] = group_bcast<LPR>(tw[k][2], j);
                w[3] = group_bcast<LPR>(tw[k][3], j);
                const int r0 = group_bcast<LPR>(tr0[k], j);
                const int r1x = group_bcast<LPR>(tr1[k], j);
                const size_t dwo = (r1x < 0) ? row_elems : 0;
                size_t off[4];
                off[0] = (size_t)r0 * row_elems;
                off[1] = off[0] + dwo;
                off[2] = (size_t)(r1x & 0x7fffffff) * row_elems;
                off[3] = off[2] + dwo;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[VEC];
                    RowVec<T>::load(base + off[c], v);
                    float dsum = 0.f;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) dsum = fmaf(g[e], v[e], dsum);
                    part[j][c] = dsum;
                    if (w[c] != 0.f) {            // masked-out corners and idle groups carry weight 0
#pragma unroll
                        for (int e = 0; e < VEC; e += 4)
                            red_add_v4(gbase + off[c] + e, w[c] * g[e], w[c] * g[e + 1], w[c] * g[e + 2], w[c] * g[e + 3]);
                    }
                }
            }
            float dot[4];
            group_reduce_scatter<LPR>(part, sub, dot);

            // ---- the lane that owns tap (sub + k*LPR) finishes it ----
            const int s = sub + k * LPR;
            if (s < LP && active) {
                const unsigned mk = tm[k];
                const float d0 = (mk & 1u) ? dot[0] : 0.f, d1 = (mk & 2u) ? dot[1] : 0.f;
                const float d2 = (mk & 4u) ? dot[2] : 0.f, d3 = (mk & 8u) ? dot[3] : 0.f;
                const float lh = tlh[k], lw = tlw[k], hh = 1.f - lh, hw = 1.f - lw;
                const float val = hh * hw * d0 + hh * lw * d1 + lh * hw * d2 + lh * lw * d3;     // cuh:155-156
                const float gw = hh * (d1 - d0) + lh * (d3 - d2);                                 // cuh:124,133,142,151
                const float gh = hw * (d2 - d0) + lw * (d3 - d1);                                 // cuh:123,132,141,150
                const long long t = pair * LP + s;
                grad_attn[t] = val;
                const float a = ta[k];
                reinterpret_cast<float2 *>(grad_loc)[t] =
                    make_float2((float)lv.W[tl[k]] * a * gw, (float)lv.H[tl[k]] * a * gh);        // cuh:157-158
            }
        }
    }
}

}  // namespace msda
