// msda_tmem.cuh -- backward with the coarse levels of grad_value accumulated in TENSOR MEMORY (D = 32, L*P <= 16).
//
// Why: the backward is bound by each SM's path into the crossbar (~22-25 B/clk/SM of red payload; §5b of DESIGN.md).
// Combining the coarse levels' contributions inside the SM halves that traffic, but doing it in SHARED memory costs 4 LSU
// data-pipe wavefronts per row-add on the pipe the gathers already saturate (msda_slab.cuh: -43 % red sectors, +36 % time).
// Tensor memory has its own port (tcgen05.ld / tcgen05.st, 12-cycle load latency) and is otherwise idle in this kernel.
//
// Layout: TMEM is 128 lanes x 512 columns x 32 bit per SM; a warp reaches only the 32 lanes of quadrant (warp % 4).  A
// window row (32 channels, fp32) is ONE COLUMN of one quadrant: row r -> quadrant r % 4, column r / 4, lane = channel.
// 4 x 512 columns = 2048 rows: at cfg2 the levels 13x21 and 25x42 (1323 rows) fit.
//
// Roles (512 threads, one CTA per SM):
//   warps 0..11  producers -- the msda_bwd_tiled work for 4 pairs each (tap geometry, row gathers, dot products, grad_loc /
//                grad_attn, red.global for the non-window levels).  Non-zero corners that fall into a window level are
//                appended to the warp's PRIVATE list for the row's quadrant (positions from ballots: no atomics) and
//                their weight is cleared in the tap record, which skips the red.global.
//   warps 12..15 consumers -- warp 12 + q owns quadrant q exclusively: it drains the 12 producer lists of its quadrant
//                with tcgen05.ld -> FFMA -> tcgen05.st, 8 columns in flight (entries of a batch that repeat a column are
//                applied afterwards, one by one).  Exclusive ownership = no atomics anywhere.
//   Lists and the stashed grad_out rows are double-buffered per tile: producers fill buffer t while consumers drain t-1;
//   one __syncthreads per tile.  When the CTA's tile range moves to another (batch, head) slab, and at the end, the
//   consumers add their columns to grad_value (red.global.add.f32, one 128-byte row per instruction) and zero them.
// A full list falls back to red.global, so capacity never affects the result.
#pragma once

#include "msda_slab.cuh"

namespace msda {

constexpr int kTmThreads = 512;
constexpr int kTmProd = 12;                    // producer warps
constexpr int kTmCons = 4;                     // consumer warps = TMEM lane quadrants
constexpr int kTmTile = kTmProd * 4;           // pairs per tile
constexpr int kTmCols = 512;                   // TMEM columns allocated (all of them: one CTA per SM)
constexpr int kTmBatch = 8;                    // columns in flight per consumer step

__host__ __device__ inline size_t bwd_tmem_smem_bytes(int list_cap) {
    return 2 * ((size_t)kTmProd * kTmCons * list_cap * 8        // entry lists
                + (size_t)kTmProd * kTmCons * 4                 // entry counts
                + (size_t)kTmTile * 128)                        // grad_out rows of the tile
           + (size_t)kTmProd * TapSlab<8>::kBytes;               // tap slabs
}

__device__ __forceinline__ void tm_ld(uint32_t taddr, uint32_t &v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
}
__device__ __forceinline__ void tm_st(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tm_wait_ld(uint32_t (&v)[kTmBatch]) {      // the registers are only valid after the wait
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}
__device__ __forceinline__ void tm_wait_ld1(uint32_t &v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v) :: "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename T, int LP_MAX>
__global__ void __launch_bounds__(kTmThreads, 1)
msda_bwd_tmem(const T *__restrict__ grad_out, const T *__restrict__ value, const int64_t *__restrict__ shapes,
              const int64_t *__restrict__ lsi, const float *__restrict__ loc, const float *__restrict__ attn,
              int N, int S, int M, int L, int Lq, int P, int sms, int list_cap,
              float *__restrict__ grad_value, float *__restrict__ grad_loc, float *__restrict__ grad_attn)
{
    constexpr int D = 32, VEC = 4, LPR = 8, NSL = LP_MAX / LPR;
    static_assert(LP_MAX % LPR == 0, "tap capacity must be whole record rounds");
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ SlabMap sm;
    __shared__ uint32_t tmem_slot;

    // dynamic shared memory: [2] lists | [2] counts | [2] g stash | tap slabs
    const size_t list_bytes = (size_t)kTmProd * kTmCons * list_cap * 8;
    uint2 *lists = reinterpret_cast<uint2 *>(dyn);                                        // [buf][warp][quad][cap]
    int *counts = reinterpret_cast<int *>(dyn + 2 * list_bytes);                          // [buf][warp][quad]
    float *gstash = reinterpret_cast<float *>(dyn + 2 * list_bytes + 2 * kTmProd * kTmCons * 4);      // [buf][slot][32]
    unsigned char *slabs = reinterpret_cast<unsigned char *>(gstash) + 2 * (size_t)kTmTile * 128;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    build_slab_map(sm, shapes, lsi, L, N, M, Lq, kTmCons * kTmCols, kTmTile);
    if (warp == kTmProd) {                       // one warp allocates tensor memory (and frees it at the end)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_slot)), "r"((unsigned)kTmCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const int wrows = sm.wrows;
    const int LP = L * P;
    const unsigned row_elems = (unsigned)(M * D);

    unsigned first, last;
    cta_tile_range(sm.ntiles, (unsigned)sms, first, last);
    const unsigned ntile = last > first ? last - first : 0u;

    if (warp >= kTmProd) {
        // ======================================= consumers =======================================
        const int q = warp - kTmProd;                                        // == warp % 4: the quadrant this warp may touch
        const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
        const int ncols = wrows > q ? (wrows - q + 3) / 4 : 0;               // window rows r with r % 4 == q  ->  column r / 4
        for (int c = 0; c < ncols; ++c) tm_st(tq + (uint32_t)c, 0u);
        tm_wait_st();

        auto flush = [&](unsigned sl) {                                       // columns -> grad_value of slab `sl`, then zero
            const int b = (int)(sl / (unsigned)M), m = (int)(sl - (unsigned)b * (unsigned)M);
            float *gslab = grad_value + ((size_t)b * S * M + m) * D + lane;
            for (int c0 = 0; c0 < ncols; c0 += kTmBatch) {
                uint32_t v[kTmBatch];
#pragma unroll
                for (int j = 0; j < kTmBatch; ++j) { v[j] = 0u; if (c0 + j < ncols) tm_ld(tq + (uint32_t)(c0 + j), v[j]); }
                tm_wait_ld(v);
#pragma unroll
                for (int j = 0; j < kTmBatch; ++j) {
                    if (c0 + j < ncols) {
                        const int wr = (c0 + j) * 4 + q;
                        int row = 0;
                        for (int l = 0; l < L; ++l) {
                            const int wb = sm.wbase[l];
                            if (wb >= 0 && wr >= wb && wr < wb + sm.H[l] * sm.W[l]) row = sm.start[l] + (wr - wb);
                        }
                        const float f = __uint_as_float(v[j]);
                        if (f != 0.f) atomicAdd(gslab + (size_t)row * row_elems, f);          // RED.E.ADD.F32: 128 B per warp
                        tm_st(tq + (uint32_t)(c0 + j), 0u);
                    }
                }
            }
            tm_wait_st();
        };

        for (unsigned it = 0; it <= ntile; ++it) {
            if (it > 0 && wrows > 0) {
                const unsigned tile = first + it - 1;
                const int buf = (int)((it - 1) & 1u);
                const float *gs = gstash + (size_t)buf * kTmTile * 32 + lane;
                for (int w = 0; w < kTmProd; ++w) {
                    const int n = min(counts[(buf * kTmProd + w) * kTmCons + q], list_cap);
                    const uint2 *lst = lists + ((size_t)(buf * kTmProd + w) * kTmCons + q) * list_cap;
                    for (int i0 = 0; i0 < n; i0 += kTmBatch) {
                        const int nb = min(kTmBatch, n - i0);
                        uint32_t col[kTmBatch], acc[kTmBatch];
                        float wt[kTmBatch], gv[kTmBatch];
                        unsigned dup = 0;                                     // bit j: entry j repeats the column of an earlier one
#pragma unroll
                        for (int j = 0; j < kTmBatch; ++j) {
                            const uint2 e = j < nb ? lst[i0 + j] : make_uint2(0xffffff00u, 0u);
                            col[j] = e.x >> 8;
                            wt[j] = __uint_as_float(e.y);
                            gv[j] = j < nb ? gs[(e.x & 255u) * 32] : 0.f;
                            acc[j] = 0u;
#pragma unroll
                            for (int k = 0; k < j; ++k) if (j < nb && col[j] == col[k]) dup |= 1u << j;
                        }
#pragma unroll
                        for (int j = 0; j < kTmBatch; ++j) if (j < nb && !((dup >> j) & 1u)) tm_ld(tq + col[j], acc[j]);
                        tm_wait_ld(acc);
#pragma unroll
                        for (int j = 0; j < kTmBatch; ++j)
                            if (j < nb && !((dup >> j) & 1u))
                                tm_st(tq + col[j], __float_as_uint(fmaf(wt[j], gv[j], __uint_as_float(acc[j]))));
                        tm_wait_st();
                        if (dup) {                                            // rare: same row twice within 8 entries
#pragma unroll
                            for (int j = 1; j < kTmBatch; ++j) {
                                if ((dup >> j) & 1u) {
                                    uint32_t a = 0u;
                                    tm_ld(tq + col[j], a);
                                    tm_wait_ld1(a);
                                    tm_st(tq + col[j], __float_as_uint(fmaf(wt[j], gv[j], __uint_as_float(a))));
                                    tm_wait_st();
                                }
                            }
                        }
                    }
                }
                const unsigned sl = tile / sm.tiles_per_slab;
                if (it == ntile || (tile + 1) / sm.tiles_per_slab != sl) flush(sl);
            }
            __syncthreads();
        }
    } else {
        // ======================================= producers =======================================
        const int sub = lane % LPR, grp = lane / LPR;
        const int slot = warp * 4 + grp;
        TapSlab<LPR> slab(slabs + warp * TapSlab<LPR>::kBytes, grp);
        const unsigned lt_mask = (1u << lane) - 1u;

        for (unsigned it = 0; it <= ntile; ++it) {
            if (it < ntile) {
                const unsigned tile = first + it;
                const int buf = (int)(it & 1u);
                const unsigned sl = tile / sm.tiles_per_slab, qt = tile - sl * sm.tiles_per_slab;
                const int b = (int)(sl / (unsigned)M), m = (int)(sl - (unsigned)b * (unsigned)M);
                const int qi = (int)qt * kTmTile + slot;
                const bool active = qi < Lq;
                const size_t pair = ((size_t)b * Lq + (active ? qi : Lq - 1)) * M + m;
                uint2 *mylist = lists + (size_t)(buf * kTmProd + warp) * kTmCons * list_cap;
                int cnt[kTmCons] = {0, 0, 0, 0};                              // warp-uniform list lengths

                float g[VEC];
                RowVec<T, VEC>::load(grad_out + pair * D + (size_t)sub * VEC, g);
                if (!active) { g[0] = g[1] = g[2] = g[3] = 0.f; }
                *reinterpret_cast<float4 *>(gstash + ((size_t)buf * kTmTile + slot) * 32 + sub * 4) = make_float4(g[0], g[1], g[2], g[3]);

                float4 tw[NSL];
                int2 tr[NSL];
                float tlh[NSL], tlw[NSL], ta[NSL];
                unsigned tmeta[NSL];
#pragma unroll
                for (int k = 0; k < NSL; ++k) {
                    const int sidx = sub + k * LPR;
                    float wc[4] = {0.f, 0.f, 0.f, 0.f};
                    int wr[4] = {-1, -1, -1, -1};                             // window row of each corner, -1 = not a window level
                    tr[k] = make_int2(0, 0);
                    tlh[k] = tlw[k] = ta[k] = 0.f; tmeta[k] = 0;
                    if (sidx < LP && active) {
                        const size_t t = pair * LP + sidx;
                        const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc) + t);
                        const float a = __ldg(attn + t);
                        const int l = sidx / P;
                        const TapGeom gm = tap_geometry(xy.x, xy.y, sm.H[l], sm.W[l], sm.start[l]);
                        const float4 w4 = masked_weights(gm, a);
                        wc[0] = w4.x; wc[1] = w4.y; wc[2] = w4.z; wc[3] = w4.w;
                        tr[k] = make_int2(gm.r0, gm.r1 | (gm.dw << 31));
                        tlh[k] = gm.lh; tlw[k] = gm.lw; ta[k] = a; tmeta[k] = gm.mask | ((unsigned)l << 4);
                        const int wb = sm.wbase[l];
                        if (wb >= 0) {
                            const int w0 = wb + (gm.r0 - sm.start[l]), w1 = wb + (gm.r1 - sm.start[l]);
                            wr[0] = w0; wr[1] = w0 + gm.dw; wr[2] = w1; wr[3] = w1 + gm.dw;
                        }
                    }
                    // converged again: warp-aggregated append of the non-zero window corners (no atomics; order = lane order)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool want = wr[c] >= 0 && wc[c] != 0.f;
#pragma unroll
                        for (int qd = 0; qd < kTmCons; ++qd) {
                            const unsigned mk = __ballot_sync(kFullMask, want && (wr[c] & 3) == qd);
                            if (want && (wr[c] & 3) == qd) {
                                const int pos = cnt[qd] + __popc(mk & lt_mask);
                                if (pos < list_cap) {
                                    mylist[qd * list_cap + pos] = make_uint2(((unsigned)(wr[c] >> 2) << 8) | (unsigned)slot,
                                                                             __float_as_uint(wc[c]));
                                    wc[c] = 0.f;                              // in the list: no red.global for this corner
                                }
                            }
                            cnt[qd] += __popc(mk);
                        }
                    }
                    tw[k] = make_float4(wc[0], wc[1], wc[2], wc[3]);
                }
                if (lane < kTmCons)
                    counts[(buf * kTmProd + warp) * kTmCons + lane] = lane == 0 ? cnt[0] : lane == 1 ? cnt[1] : lane == 2 ? cnt[2] : cnt[3];

                const size_t slab_off = ((size_t)b * S * M + m) * D + (size_t)sub * VEC;
                const T *base = value + slab_off;
                float *gbase = grad_value + slab_off;
#pragma unroll
                for (int k = 0; k < NSL; ++k) {
                    __syncwarp();
                    slab.put(sub, tw[k], tr[k]);
                    __syncwarp();
                    float part[LPR][4];
#pragma unroll
                    for (int j = 0; j < LPR; ++j) {
                        const float4 w4 = slab.weights(j);
                        const int2 rr = slab.rows(j);
                        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
                        const unsigned dwo = (rr.y < 0) ? row_elems : 0u;
                        unsigned long long off[4];
                        off[0] = (unsigned long long)(unsigned)rr.x * row_elems;
                        off[1] = off[0] + dwo;
                        off[2] = (unsigned long long)(unsigned)(rr.y & 0x7fffffff) * row_elems;
                        off[3] = off[2] + dwo;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float v[VEC];
                            RowVec<T, VEC>::load(base + off[c], v);
                            float dsum = 0.f;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) dsum = fmaf(g[e], v[e], dsum);
                            part[j][c] = dsum;
                            if (w[c] != 0.f)
                                red_add_v4(gbase + off[c], w[c] * g[0], w[c] * g[1], w[c] * g[2], w[c] * g[3]);
                        }
                    }
                    float dot[4];
                    group_reduce_scatter<LPR>(part, sub, dot);

                    const int sidx = sub + k * LPR;
                    if (sidx < LP && active) {
                        const unsigned mk = tmeta[k];
                        const int l = (int)(mk >> 4);
                        const float d0 = (mk & 1u) ? dot[0] : 0.f, d1 = (mk & 2u) ? dot[1] : 0.f;
                        const float d2 = (mk & 4u) ? dot[2] : 0.f, d3 = (mk & 8u) ? dot[3] : 0.f;
                        const float lh = tlh[k], lw = tlw[k], hh = 1.f - lh, hw = 1.f - lw;
                        const float val = hh * hw * d0 + hh * lw * d1 + lh * hw * d2 + lh * lw * d3;   // cuh:155-156
                        const float gw = hh * (d1 - d0) + lh * (d3 - d2);                               // cuh:124,133,142,151
                        const float gh = hw * (d2 - d0) + lw * (d3 - d1);                               // cuh:123,132,141,150
                        const size_t t = pair * LP + sidx;
                        grad_attn[t] = val;
                        const float a = ta[k];
                        reinterpret_cast<float2 *>(grad_loc)[t] =
                            make_float2((float)sm.W[l] * a * gw, (float)sm.H[l] * a * gh);              // cuh:157-158
                    }
                }
            }
            __syncthreads();
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kTmProd)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)kTmCols) : "memory");
}

}  // namespace msda
