// msda_gemm_sm100.cu -- hand-written tcgen05 GEMM for the Linears that bracket the op (north_star: "tensor cores only
// on the fused value_proj / output_proj GEMMs"):      C[M, N] = A[M, K] . W[N, K]^T + bias[N]        (fp32 in/out, TF32 MMA)
//
// Shape regime: M = N_batch * S tokens (tens of thousands), K = d_model (256), N in {256, 384}.  Memory-bound: A is read
// once, C written once, W (<= 384 KB) is re-read by every CTA from L2.
//
// Structure (one CTA = one 128-row tile of C, all N columns):
//   warp 0   : TMA producer -- cp.async.bulk.tensor 2D loads of A[128 x 32] and W[N x 32] (128-byte rows, SWIZZLE_128B)
//              into a ring of shared-memory stages, completion on mbarriers.
//   warp 1   : TMEM allocation + MMA issue -- one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N<=256,
//              K=8 per instruction, 4 per 32-wide k-block), accumulating in TMEM; tcgen05.commit releases stages.
//   warps 2-5: epilogue -- tcgen05.ld 32 lanes x 32 columns at a time, + bias, 16-byte global stores (one row per thread).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <mutex>

#include "../../include/msda_b200.h"

extern std::atomic<uint64_t> g_msda_gemm_launches;
std::atomic<uint64_t> g_msda_gemm_launches{0};

namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                 // fp32 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;                   // tf32: 32 bytes per instruction
constexpr int kThreads = 192;
constexpr int kMaxN = 512;
constexpr int kEpiWarpsWS2 = 8, kThreadsWS2 = 64 + 32 * kEpiWarpsWS2;   // 2-CTA W-stationary kernel: producer, MMA, 8 epilogue warps

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
                 "[%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// L2 prefetch of one TMA box (no shared-memory destination, no barrier): hides the DRAM part of a later tile load's latency
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset [16,30) (unused for swizzled K-major: 1), stride byte offset [32,46) = 1024 B between 8-row groups,
// version = 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t umma_desc(const void *smem, unsigned k_byte_offset) {
    const uint64_t addr = (smem_u32(smem) + k_byte_offset) >> 4;
    return (addr & 0x3fffull) | (1ull << 16) | ((1024ull >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor for kind::tf32: c_format F32 (1) at [4,6), a/b format TF32 (2) at [7,10)/[10,13), both
// K-major, n_dim = N >> 3 at [17,23), m_dim = M >> 4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` (both CTAs of the pair share each W stage)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Params {
    long long M;
    int N, K, stages, tmem_cols, w_box_rows, acc_bufs;      // acc_bufs = 2 when two accumulators fit in TMEM (N <= 256)
    const float *bias;
    float *C;
};

// Persistent: cluster c (a CTA pair) handles row-tile pairs c, c+G/2, ...; CTA `rank` of the pair owns tile 2*pair+rank.
// Three pipelines: smem stages (TMA <-> MMA, full/empty), TMEM accumulators (MMA <-> epilogue, tmem_full/tmem_empty,
// double-buffered when N <= 256), and the tile loop itself.  The pair shares W: each CTA loads half of W's rows per
// k-block and TMA-multicasts it into both CTAs' stage (halves the L2 -> SM traffic of the operand every tile re-reads);
// a stage is refilled only after BOTH CTAs' MMAs have consumed it (empty barriers count 2, signalled by multicast commits).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
linear_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p)
{
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B atoms (8 rows x 128 B) must start 1024-byte aligned
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = p.N, KB = p.K / BLOCK_K;
    const unsigned a_bytes = BLOCK_M * BLOCK_K * 4, w_bytes = (unsigned)N * BLOCK_K * 4, stage_bytes = a_bytes + w_bytes;
    const long long ntiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const long long npairs = (ntiles + 1) / 2;
    const unsigned rank = cluster_ctarank();
    const long long pair0 = blockIdx.x / 2, pair_stride = gridDim.x / 2;
    // per epilogue warp: 32x33 fp32 transpose tile, placed after the stage ring
    float (*xpose)[32][33] = reinterpret_cast<float (*)[32][33]>(smem + (size_t)p.stages * stage_bytes);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                       // one warp allocates TMEM (power-of-two columns) and later frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((unsigned)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                    // the peer's barriers exist before anything is multicast into this CTA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            unsigned s = 0, ph = 0;                                      // ring position across tiles (no div / mod)
            const int half = N / 2;                                      // rows of W this CTA loads and multicasts
            for (long long pr = pair0; pr < npairs; pr += pair_stride) {
                const int row0 = (int)((pr * 2 + rank) * BLOCK_M);       // may lie beyond M for the last pair: TMA zero-fills
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);                    // both CTAs have drained this stage
                    unsigned char *sa = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full_bar[s], stage_bytes);           // own A tile + both halves of W
                    tma_load_2d(sa, &map_a, kb * BLOCK_K, row0, &full_bar[s]);
                    tma_load_2d_mc(sa + a_bytes + (size_t)rank * half * BLOCK_K * 4, &map_w, kb * BLOCK_K, (int)rank * half,
                                   &full_bar[s], (uint16_t)3);
                    if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        unsigned it = 0, t = 0, s = 0, ph = 0;
        // one thread issues every MMA: descriptors are built once, per MMA only their 16-byte-unit address field moves
        const uint64_t a_desc0 = umma_desc(smem, 0), w_desc0 = umma_desc(smem + a_bytes, 0);
        const uint32_t st_step = stage_bytes >> 4, k_step = (UMMA_K * 4) >> 4, n_step = (256u * BLOCK_K * 4) >> 4;
        const uint32_t idesc0 = umma_idesc_tf32(BLOCK_M, min(256, N)), idesc1 = N > 256 ? umma_idesc_tf32(BLOCK_M, N - 256) : 0u;
        for (long long pr = pair0; pr < npairs; pr += pair_stride, ++t) {
            const unsigned buf = t % p.acc_bufs, use = t / p.acc_bufs;
            mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);              // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + buf * (uint32_t)N;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t ad = a_desc0 + (uint64_t)(s * st_step), wd = w_desc0 + (uint64_t)(s * st_step);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_tf32(acc, ad + (uint64_t)(k * k_step), wd + (uint64_t)(k * k_step), idesc0, (kb | k) != 0);
                    if (N > 256) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_tf32(acc + 256u, ad + (uint64_t)(k * k_step), wd + (uint64_t)(n_step + k * k_step), idesc1, (kb | k) != 0);
                    }
                    umma_commit_mc(&empty_bar[s], (uint16_t)3);          // tell both producers this CTA is done with the stage
                    if (kb == KB - 1) umma_commit(&tmem_full_bar[buf]);  // accumulator complete
                }
                __syncwarp();
                if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> smem transpose -> (+bias) -> coalesced global stores =====
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;                          // a warp may touch TMEM lanes 32*(warp%4) .. +31
        float (*xp)[33] = xpose[ew];
        unsigned t = 0;
        for (long long pr = pair0; pr < npairs; pr += pair_stride, ++t) {
            const unsigned buf = t % p.acc_bufs, use = t / p.acc_bufs;
            mbar_wait(&tmem_full_bar[buf], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long rbase = (pr * 2 + rank) * BLOCK_M + lane_base;
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + buf * (uint32_t)N + (uint32_t)c0, v);
                const float bia = p.bias != nullptr ? __ldg(p.bias + c0 + lane) : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) xp[lane][i] = v[i];        // lane = row; conflict-free with the 33 pad
                __syncwarp();
#pragma unroll 8
                for (int r = 0; r < 32; ++r)                             // lane = column: 128 contiguous bytes per row
                    if (rbase + r < p.M) p.C[(rbase + r) * (long long)N + c0 + lane] = xp[r][lane] + bia;
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[buf])) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                    // neither CTA leaves while the peer may still multicast into it
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)p.tmem_cols) : "memory");
}

// =====================================================================================================================
// W-STATIONARY variant for the module's own shapes (K <= 256, N <= 256: value_proj / output_proj, 256 -> 256).
//
// The streaming kernel above re-reads W (256 KB) from L2 for every 128-row tile -- as many bytes as the A tile itself --
// and its 48 KB stages leave room for 4 of them.  Here the roles of the pair's operands are swapped:
//   * CTA r of the pair owns the output COLUMNS [r N/2, (r+1) N/2) and keeps its half of W (N/2 x K fp32 <= 128 KB)
//     resident in shared memory for the whole kernel (loaded once);
//   * both CTAs need the same A tile: each loads 64 of its 128 rows per k-block and TMA-multicasts them into both CTAs'
//     stage, so A crosses L2 -> SM once per pair;  stages are 16 KB -> 5 in flight next to the resident W;
//   * MMA stays cta_group::1 (M = 128, N = N/2), accumulators double-buffered in tensor memory;
//   * epilogue straight from registers: tcgen05.ld gives every thread 32 consecutive columns of its row = one 128-byte
//     segment, written with 16-byte stores after the fused tail  (+ bias) -> (row mask -> 0) -> (ReLU).
// The row mask is ops/modules/ms_deform_attn.py:96-97 (`value.masked_fill(input_padding_mask[..., None], 0)`) folded into
// value_proj's epilogue; ReLU is the FFN's first activation.
struct ParamsWS {
    long long M;
    int N, K, stages, tmem_cols, relu;
    int multicast;                      // 1: the pair shares each A tile by TMA multicast; 0: every CTA loads its own copy
    int direct_epilogue;                // 1: registers -> 32-byte global stores (no shared memory); 0: smem transpose
    const float *bias;
    const unsigned char *row_mask;      // [M] bytes, non-zero = zero the whole output row; may be null
    float *C;
    long long rows_per_pair;            // 2-CTA kernel: contiguous rows per CTA pair (multiple of 32)
    int dbg;                            // 2-CTA kernel: 1 = record the phase timeline
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
linear_tf32_ws_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_full,
                      const __grid_constant__ CUtensorMap map_w, const ParamsWS p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NH = p.N / 2, KB = p.K / BLOCK_K;
    const unsigned a_bytes = BLOCK_M * BLOCK_K * 4;                     // one A stage: 128 rows x 128 B
    const unsigned wk_bytes = (unsigned)NH * BLOCK_K * 4;               // one k-block of this CTA's half of W
    unsigned char *w_res = smem;                                        // KB x [NH x 128 B], resident
    unsigned char *stages = smem + (size_t)KB * wk_bytes;
    const long long ntiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    const unsigned rank = cluster_ctarank();
    const long long tile0 = blockIdx.x / 2, tile_stride = gridDim.x / 2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_full) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], p.multicast ? 2 : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
        mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((unsigned)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer: W once, then the A stream =====
        if (lane == 0) {
            mbar_expect_tx(&w_bar, (unsigned)KB * wk_bytes);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(w_res + (size_t)kb * wk_bytes, &map_w, kb * BLOCK_K, (int)rank * NH, &w_bar);
            unsigned it = 0, s = 0, ph = 0;
            // The ring holds only 4-5 A stages next to the resident W (64-80 KB in flight per SM), too little to cover the
            // DRAM latency of a tile load: each CTA therefore prefetches ITS half of the NEXT tile into L2 (no smem, no
            // barrier) while the current tile streams, so that the ring's loads are L2 hits.
            for (int kb = 0; kb < KB && tile0 < ntiles; ++kb)
                tma_prefetch_2d(&map_a, kb * BLOCK_K, (int)(tile0 * BLOCK_M) + (int)rank * (BLOCK_M / 2));
            for (long long t = tile0; t < ntiles; t += tile_stride) {
                const int row0 = (int)(t * BLOCK_M) + (int)rank * (BLOCK_M / 2);       // this CTA's 64 rows of the shared tile
                if (t + tile_stride < ntiles)
                    for (int kb = 0; kb < KB; ++kb)
                        tma_prefetch_2d(&map_a, kb * BLOCK_K, (int)((t + tile_stride) * BLOCK_M) + (int)rank * (BLOCK_M / 2));
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1);                    // the MMAs (of both CTAs when sharing) have drained it
                    mbar_expect_tx(&full_bar[s], a_bytes);               // whole tile: own half + the peer's half, or own copy
                    if (p.multicast)
                        tma_load_2d_mc(stages + (size_t)s * a_bytes + (size_t)rank * (a_bytes / 2), &map_a, kb * BLOCK_K, row0,
                                       &full_bar[s], (uint16_t)3);
                    else
                        tma_load_2d(stages + (size_t)s * a_bytes, &map_a_full, kb * BLOCK_K, (int)(t * BLOCK_M), &full_bar[s]);
                    if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        mbar_wait(&w_bar, 0);
        unsigned it = 0, tt = 0;
        const uint32_t idesc = umma_idesc_tf32(BLOCK_M, NH);
        // The issuing lane is ONE thread: everything it executes per MMA is on the critical path of the tensor pipe.  The
        // descriptors are therefore built once; per MMA only their address field (bits 0-13, units of 16 bytes) is bumped.
        const uint64_t a_desc0 = umma_desc(stages, 0), w_desc0 = umma_desc(w_res, 0);
        const uint32_t a_step = a_bytes >> 4, w_step = wk_bytes >> 4, k_step = (UMMA_K * 4) >> 4;
        unsigned s = 0, ph = 0;                                          // ring position of `it` without div / mod
        for (long long t = tile0; t < ntiles; t += tile_stride, ++tt) {
            const unsigned buf = tt & 1, use = tt >> 1;
            mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + buf * (uint32_t)NH;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t ad = a_desc0 + (uint64_t)(s * a_step), wd = w_desc0 + (uint64_t)((unsigned)kb * w_step);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_tf32(acc, ad + (uint64_t)(k * k_step), wd + (uint64_t)(k * k_step), idesc, (kb | k) != 0);
                    if (p.multicast) umma_commit_mc(&empty_bar[s], (uint16_t)3);   // free in BOTH CTAs once both have committed
                    else umma_commit(&empty_bar[s]);
                    if (kb == KB - 1) umma_commit(&tmem_full_bar[buf]);
                }
                __syncwarp();
                if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> per-warp smem transpose -> (+bias, mask, ReLU) -> coalesced 16-byte stores =====
        // tcgen05.ld hands every thread 32 consecutive columns of ITS row; storing those directly makes each warp store
        // touch 32 different lines (measured: 5 us per tile, the whole kernel 31 us).  Through a 32 x 36 staging tile a
        // warp store covers 4 rows x 128 B instead: lanes (r, c4) = (lane / 8, lane % 8).
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        float (*xp)[36] = reinterpret_cast<float (*)[36]>(stages + (size_t)p.stages * a_bytes + (size_t)ew * (32 * 36 * 4));
        const int rr = lane >> 3, c4 = (lane & 7) * 4;
        unsigned tt = 0;
        for (long long t = tile0; t < ntiles; t += tile_stride, ++tt) {
            const unsigned buf = tt & 1, use = tt >> 1;
            mbar_wait(&tmem_full_bar[buf], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long row_base = t * BLOCK_M + lane_base;
            unsigned dead_bits = 0;                                     // bit i: row (row_base + rr + 4 i) is masked
            if (p.row_mask != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long r = row_base + rr + 4 * i;
                    if (r < p.M && p.row_mask[r] != 0) dead_bits |= 1u << i;
                }
            }
            if (p.direct_epilogue >= 2) {            // timing experiments: 2 = nothing, 3 = TMEM loads only
                if (p.direct_epilogue == 3)
                    for (int c0 = 0; c0 < NH; c0 += 32) {
                        float v[32];
                        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + buf * (uint32_t)NH + (uint32_t)c0, v);
                        if (v[0] == 123.456f && row_base < 0) p.C[0] = v[1];
                    }
            } else if (p.direct_epilogue) {
                // The tf32 MMA reads its operands from shared memory at ~119 B/clk (8 KB per 69-cycle instruction) -- the
                // whole shared-memory port.  A transpose through shared memory therefore cannot overlap the next tile's
                // MMAs (measured: MMA phase + epilogue ADD UP, 4.1 us per tile).  Here every thread writes its own row
                // straight from registers, one full 32-byte sector per store.
                const long long r = row_base + lane;
                const bool live = r < p.M;
                const bool dead = live && p.row_mask != nullptr && p.row_mask[r] != 0;
                float *crow = p.C + (live ? r : 0) * (long long)p.N + (long long)rank * NH;
                for (int c0 = 0; c0 < NH; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + buf * (uint32_t)NH + (uint32_t)c0, v);
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                o[e] = v[i + e] + (p.bias != nullptr ? __ldg(p.bias + (size_t)rank * NH + c0 + i + e) : 0.f);
                                if (p.relu) o[e] = fmaxf(o[e], 0.f);
                                if (dead) o[e] = 0.f;
                            }
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(crow + c0 + i), "f"(o[0]),
                                         "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
                        }
                    }
                }
            } else
            for (int c0 = 0; c0 < NH; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + buf * (uint32_t)NH + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4 *>(&xp[lane][i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + (size_t)rank * NH + c0 + c4));
                float4 o8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o8[i] = *reinterpret_cast<const float4 *>(&xp[rr + 4 * i][c4]);     // all loads first
                __syncwarp();                                            // staging tile free for the next chunk
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long r = row_base + rr + 4 * i;
                    float4 o = o8[i];
                    o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (dead_bits & (1u << i)) o = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < p.M) *reinterpret_cast<float4 *>(p.C + r * (long long)p.N + (long long)rank * NH + c0 + c4) = o;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[buf])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)p.tmem_cols) : "memory");
}

// =====================================================================================================================
// W-STATIONARY, 2-CTA MMA (cta_group::2).  Same residency as above -- CTA r of the pair keeps rows [r N/2, (r+1) N/2) of W
// in shared memory -- but the pair now works on a 256-row tile with ONE tcgen05.mma.cta_group::2 (M = 256, N = N) per k-step,
// issued by the leader CTA: each CTA supplies its own 128 rows of A and its half of W, and receives the accumulator of
// its 128 rows x all N columns in its own tensor memory.  Per output row that halves the two big shared-memory streams of
// the 1-CTA version (the MMA's operand reads -- a tf32 MMA pulls 8 KB per 69-cycle instruction, the whole port -- and the
// TMA writes: no CTA receives rows it does not own), which is what bounded it (shared memory: 512 KB per 128 rows there,
// 640 KB per 256 rows here).
//   barriers: the A stage of BOTH CTAs completes on the LEADER's full barrier (the peer's TMA signals it through the
//   shared::cluster address with the CTA-rank bit cleared); tcgen05.commit multicasts "stage free" and "accumulator full" to
//   both CTAs; the epilogue warps of both CTAs arrive on the leader's "accumulator empty" barrier (count 8).
__device__ __forceinline__ unsigned leader_addr(const void *p) { return smem_u32(p) & 0xFEFFFFFFu; }   // same offset in the pair's even CTA
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, int c0, int c1, unsigned bar_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(bar_addr) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// Diagnostic timeline (MSDA_GEMM_WS_DBG=1): per CTA, SM-clock stamps of the phases below, read back with
// msda_debug_gemm_timeline().  Slot 0: %globaltimer at entry (ns); 1: entry; 2: set-up done; 3: first stage landed;
// 4 + 2 i / 5 + 2 i: accumulator of tile i complete / stored (epilogue warp 2); 20: producer done; 21: MMA issue done; 22: exit.
constexpr int kTlSlots = 24;
__device__ unsigned long long g_ws2_timeline[160][kTlSlots];
__device__ __forceinline__ void tl_stamp(int dbg, int slot) {
    if (dbg && blockIdx.x < 160) g_ws2_timeline[blockIdx.x][slot] = (unsigned long long)clock64();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsWS2, 1)
linear_tf32_ws2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_c, const ParamsWS p)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2], w_bar[8];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.dbg && threadIdx.x == 0 && blockIdx.x < 160) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        g_ws2_timeline[blockIdx.x][0] = gt;
        tl_stamp(1, 1);
    }
    const int N = p.N, NH = N / 2, KB = p.K / BLOCK_K;
    const unsigned a_bytes = BLOCK_M * BLOCK_K * 4;                     // this CTA's 128 rows of one k-block
    const unsigned wk_bytes = (unsigned)NH * BLOCK_K * 4;               // one k-block of this CTA's half of W
    unsigned char *w_res = smem;
    unsigned char *stages = smem + (size_t)KB * wk_bytes;
    const unsigned rank = cluster_ctarank();
    const bool leader = rank == 0;
    // Rows are dealt out in equal contiguous ranges (multiples of 32 = one store box), not in whole 256-row tiles: with
    // 175 tiles on 74 pairs a third of the machine would run a third round alone.  A pair walks its range in 256-row
    // tiles; the last one may be short: CTA r then owns rows [ts + r h, ts + (r + 1) h) with h < 128 (the loads and the MMA
    // still cover 128 rows per CTA -- extra rows are real rows of A whose results are simply not stored).
    const long long range0 = (long long)(blockIdx.x / 2) * p.rows_per_pair;
    const long long range1 = range0 + p.rows_per_pair < p.M ? range0 + p.rows_per_pair : p.M;
    const int ntiles = range1 > range0 ? (int)((range1 - range0 + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) : 0;
    auto tile_rows = [&](int i, long long &row0) -> int {               // -> h, and this CTA's first row
        const long long ts = range0 + (long long)i * 2 * BLOCK_M, rem = range1 - ts;
        const int h = rem >= 2 * BLOCK_M ? BLOCK_M : (int)((rem + 63) / 64) * 32;
        row0 = ts + (long long)rank * h;
        return h;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 2 * kEpiWarpsWS2); }
        for (int i = 0; i < 8; ++i) mbar_init(&w_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((unsigned)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // (allocating AFTER the cluster barrier, overlapped with the first loads, measured slower: the pair's allocation takes
    // ~1.5 us and then sits on the leader's critical path, which must know the PEER's tensor memory is ready too)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    if (threadIdx.x == 0) tl_stamp(p.dbg, 2);

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own half of W once, then own 128 rows of every tile; completion on the leader =====
        if (lane == 0) {
            // W arrives k-block by k-block, each with its own barrier, interleaved with the first tile's A stages, so the
            // first MMA waits for 16 + 16 KB, not for the whole 128 KB of W.
            int kbw = 0;
            auto load_w = [&]() {
                if (leader) mbar_expect_tx(&w_bar[kbw], 2u * wk_bytes);
                tma_load_2d_2sm(w_res + (size_t)kbw * wk_bytes, &map_w, kbw * BLOCK_K, (int)rank * NH, leader_addr(&w_bar[kbw]));
                ++kbw;
            };
            unsigned s = 0, ph = 0;
            for (int i = 0; i < ntiles; ++i) {
                long long row0;
                tile_rows(i, row0);
                for (int kb = 0; kb < KB; ++kb) {
                    if (i == 0 && kb < p.stages && kbw < KB) load_w();
                    if (i == 0 && kb == p.stages) while (kbw < KB) load_w();         // before the first wait that can block
                    mbar_wait(&empty_bar[s], ph ^ 1);                    // the pair's MMA has drained this stage (commit reaches both CTAs)
                    if (leader) mbar_expect_tx(&full_bar[s], 2u * a_bytes);
                    tma_load_2d_2sm(stages + (size_t)s * a_bytes, &map_a, kb * BLOCK_K, (int)row0, leader_addr(&full_bar[s]));
                    if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
                }
                while (i == 0 && kbw < KB) load_w();                                 // K / 32 <= stages
            }
            tl_stamp(p.dbg, 20);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's elected lane drives both CTAs' tensor cores =====
        if (leader) {
            unsigned tt = 0;
            const uint32_t idesc = umma_idesc_tf32(2 * BLOCK_M, N);
            const uint64_t a_desc0 = umma_desc(stages, 0), w_desc0 = umma_desc(w_res, 0);
            const uint32_t a_step = a_bytes >> 4, w_step = wk_bytes >> 4, k_step = (UMMA_K * 4) >> 4;
            unsigned s = 0, ph = 0;
            for (; tt < (unsigned)ntiles; ++tt) {
                const unsigned buf = tt & 1, use = tt >> 1;
                mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + buf * (uint32_t)N;
                for (int kb = 0; kb < KB; ++kb) {
                    if (tt == 0) mbar_wait(&w_bar[kb], 0);               // both halves of this k-block of W are resident
                    mbar_wait(&full_bar[s], ph);
                    if (tt == 0 && kb == 0 && lane == 0) tl_stamp(p.dbg, 3);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (lane == 0) {
                        const uint64_t ad = a_desc0 + (uint64_t)(s * a_step), wd = w_desc0 + (uint64_t)((unsigned)kb * w_step);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma2_tf32(acc, ad + (uint64_t)(k * k_step), wd + (uint64_t)(k * k_step), idesc, (kb | k) != 0);
                        umma2_commit_mc(&empty_bar[s], (uint16_t)3);
                        if (kb == KB - 1) umma2_commit_mc(&tmem_full_bar[buf], (uint16_t)3);
                    }
                    __syncwarp();
                    if (++s == (unsigned)p.stages) { s = 0; ph ^= 1u; }
                }
            }
            if (lane == 0) tl_stamp(p.dbg, 21);
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows x N columns.  A warp may only touch the tensor-memory lanes of its quadrant
        // (warp % 4), so the 8 epilogue warps pair up per quadrant and split the columns.  Per 32-column chunk:
        // tcgen05.ld (thread = row, 32 consecutive columns) -> fused tail in registers -> the row goes into the warp's
        // 32 x 128-byte staging tile in TMA's SWIZZLE_128B order (16-byte chunk j of row r at chunk j ^ (r % 8)) -> one TMA
        // store of the 32 x 32 box (rows past M are clipped by the tensor map).  No shared-memory read-back by the warp and
        // no per-thread global stores: the 1-CTA kernel's transposing epilogue was its slowest phase (5.5 us per tile).
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        const int col_lo = (ew >> 2) * NH;                               // this warp's NH columns
        float *xp = reinterpret_cast<float *>(stages + (size_t)p.stages * a_bytes + (size_t)ew * (32 * 32 * 4));
        for (unsigned tt = 0; tt < (unsigned)ntiles; ++tt) {
            const unsigned buf = tt & 1, use = tt >> 1;
            mbar_wait(&tmem_full_bar[buf], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 2 && lane == 0 && tt < 8) tl_stamp(p.dbg, 4 + 2 * (int)tt);
            long long row0;
            const int h = tile_rows((int)tt, row0);
            const long long row_base = row0 + lane_base;
            const long long r = row_base + lane;
            const bool dead = p.row_mask != nullptr && r < p.M && p.row_mask[r] != 0;
            if (p.direct_epilogue != 2 && lane_base < h)                // a short tile leaves the upper quadrants without rows
            for (int c0 = col_lo; c0 < col_lo + NH; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + buf * (uint32_t)N + (uint32_t)c0, v);
                if (p.direct_epilogue == 3) {                            // timing experiment: tensor-memory loads only
                    if (v[0] == 123.456f && row_base < 0) p.C[0] = v[1];
                    continue;
                }
                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + 4 * j));     // warp-uniform address
                        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (p.relu) v[i] = fmaxf(v[i], 0.f);
                    if (dead) v[i] = 0.f;
                }
                // (borrowing the idle A stages as extra staging tiles in the last tile, so that no box waits for the previous
                // store to drain, was measured: no change -- the tail is the memory system draining, not this wait)
                float *tile = xp;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // previous box has left the staging tile
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(tile + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                        // generic-proxy writes -> visible to the TMA
                __syncwarp();
                if (lane == 0 && row_base < p.M) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                                 ::"l"(&map_c), "r"(c0), "r"((int)row_base), "r"(smem_u32(tile)) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (warp == 2 && lane == 0 && tt < 8) tl_stamp(p.dbg, 5 + 2 * (int)tt);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_addr(&tmem_empty_bar[buf])) : "memory");
        }
    }
    if (warp >= 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging tile has been read; the writes complete with the grid
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) tl_stamp(p.dbg, 22);
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((unsigned)p.tmem_cols) : "memory");
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

bool make_map(CUtensorMap *map, const float *base, long long rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gemm

namespace gemm {

// W-stationary eligibility: the pair's halves of W must fit next to >= 3 A stages.
bool ws_ok(int N, int K) {
    if (N % 64 || N < 64 || N > 256 || K % BLOCK_K) return false;
    const size_t w_half = (size_t)(N / 2) * K * 4;
    return w_half + 3u * (BLOCK_M * BLOCK_K * 4) + 4 * 32 * 36 * 4 + 2048 <= 232448 - 1024;
}

int sms_for_device(const void *kernel, int smem_bytes, std::atomic<int> (&cache)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    int sms = cache[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cache[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

int launch_ws(const float *A, const float *W, const float *bias, const unsigned char *row_mask, long long M, int N, int K,
              int relu, float *C, cudaStream_t stream) {
    CUtensorMap map_a, map_a_full, map_w;
    if (!make_map(&map_a, A, M, K, BLOCK_M / 2) || !make_map(&map_a_full, A, M, K, BLOCK_M) || !make_map(&map_w, W, N, K, N / 2))
        return MSDA_E_NODEVICE;
    static int mc = -1;
    if (mc < 0) { const char *e = getenv("MSDA_GEMM_WS_MC"); mc = (e && e[0] == '0') ? 0 : 1; }
    // epilogue: 0 = shared-memory transpose (default, measured faster), 1 = "direct" register stores;
    // 2 = "none", 3 = "ldonly": TIMING EXPERIMENTS ONLY (no / partial output), never used by the library's callers
    static int direct = -1;
    if (direct < 0) {
        const char *e = getenv("MSDA_GEMM_WS_EPI");
        direct = (e && e[0] == 'd') ? 1 : (e && e[0] == 'n') ? 2 : (e && e[0] == 'l') ? 3 : 0;
    }
    ParamsWS p;
    p.multicast = mc;
    p.rows_per_pair = 0;
    p.dbg = 0;
    p.direct_epilogue = direct >= 2 ? direct : (direct == 1 && ((reinterpret_cast<uintptr_t>(C) & 31u) == 0) && (N % 16 == 0)) ? 1 : 0;
    p.M = M; p.N = N; p.K = K; p.relu = relu; p.bias = bias; p.row_mask = row_mask; p.C = C;
    const size_t w_half = (size_t)(N / 2) * K * 4, a_stage = BLOCK_M * BLOCK_K * 4;
    constexpr size_t kDynMax = 232448 - 1024;
    constexpr size_t kXpose = 4 * 32 * 36 * 4;                                  // one 32 x 36 staging tile per epilogue warp
    int stages = (int)((kDynMax - 1024 - kXpose - w_half) / a_stage);
    if (stages > 8) stages = 8;
    if (stages < 2) return MSDA_E_BADARG;
    p.stages = stages;
    p.tmem_cols = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;          // 2 accumulators of N/2 columns
    const size_t smem = w_half + (size_t)stages * a_stage + kXpose + 1024;
    static std::atomic<int> cache[64];
    const int sms = sms_for_device(reinterpret_cast<const void *>(linear_tf32_ws_kernel), (int)kDynMax, cache);
    if (sms < 0) return MSDA_E_NODEVICE;
    const long long tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const long long max_clusters = sms / 2;
    const unsigned grid = 2u * (unsigned)(tiles < max_clusters ? tiles : max_clusters);
    linear_tf32_ws_kernel<<<grid, kThreads, smem, stream>>>(map_a, map_a_full, map_w, p);
    g_msda_gemm_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int launch_ws2(const float *A, const float *W, const float *bias, const unsigned char *row_mask, long long M, int N, int K,
               int relu, float *C, cudaStream_t stream) {
    CUtensorMap map_a, map_w, map_c;
    if (!make_map(&map_a, A, M, K, BLOCK_M) || !make_map(&map_w, W, N, K, N / 2) || !make_map(&map_c, C, M, N, 32))
        return MSDA_E_NODEVICE;
    static int epi = -1;                     // MSDA_GEMM_WS_EPI=none: timing experiment without the epilogue's work
    if (epi < 0) {                           // none / ldonly: timing experiments, no valid output
        const char *e = getenv("MSDA_GEMM_WS_EPI");
        epi = (e && e[0] == 'n') ? 2 : (e && e[0] == 'l') ? 3 : 0;
    }
    ParamsWS p;
    p.multicast = 0; p.direct_epilogue = epi;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("MSDA_GEMM_WS_DBG"); dbg = (e && e[0] == '1') ? 1 : 0; }
    p.dbg = dbg;
    p.M = M; p.N = N; p.K = K; p.relu = relu; p.bias = bias; p.row_mask = row_mask; p.C = C;
    const size_t w_half = (size_t)(N / 2) * K * 4, a_stage = BLOCK_M * BLOCK_K * 4;
    constexpr size_t kDynMax = 232448 - 1024;
    constexpr size_t kXpose = (size_t)kEpiWarpsWS2 * 32 * 32 * 4;              // one swizzled 32 x 32 staging tile per epilogue warp
    int stages = (int)((kDynMax - 1024 - kXpose - w_half) / a_stage);
    if (stages > 8) stages = 8;
    if (stages < 2) return MSDA_E_BADARG;
    p.stages = stages;
    p.tmem_cols = 2 * N <= 32 ? 32 : 2 * N <= 64 ? 64 : 2 * N <= 128 ? 128 : 2 * N <= 256 ? 256 : 512;    // 2 accumulators of N columns
    const size_t smem = w_half + (size_t)stages * a_stage + kXpose + 1024;
    static std::atomic<int> cache[64];
    const int sms = sms_for_device(reinterpret_cast<const void *>(linear_tf32_ws2_kernel), (int)kDynMax, cache);
    if (sms < 0) return MSDA_E_NODEVICE;
    const long long tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const long long max_clusters = sms / 2;
    const long long pairs = tiles < max_clusters ? tiles : max_clusters;
    p.rows_per_pair = (((M + pairs - 1) / pairs) + 31) / 32 * 32;
    const unsigned grid = 2u * (unsigned)pairs;
    linear_tf32_ws2_kernel<<<grid, kThreadsWS2, smem, stream>>>(map_a, map_w, map_c, p);
    g_msda_gemm_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// which W-stationary kernel: the 2-CTA MMA version unless MSDA_GEMM_WS2=0 (the 1-CTA kernel with multicast A, kept for A/B runs)
bool use_ws2() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("MSDA_GEMM_WS2"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

}  // namespace gemm

extern "C" int msda_linear_tf32_ex(const float *A, const float *W, const float *bias, const uint8_t *row_mask, int64_t M, int N,
                                   int K, int relu, float *C, void *stream) {
    using namespace gemm;
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0) return MSDA_E_BADARG;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15u) return MSDA_E_BADARG;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15u)) return MSDA_E_BADARG;
    if (!ws_ok(N, K)) return MSDA_E_BADARG;
    if (use_ws2() && K <= 8 * BLOCK_K)        // the 2-CTA kernel keeps one barrier per k-block of W (8)
        return launch_ws2(A, W, bias, row_mask, M, N, K, relu ? 1 : 0, C, static_cast<cudaStream_t>(stream));
    return launch_ws(A, W, bias, row_mask, M, N, K, relu ? 1 : 0, C, static_cast<cudaStream_t>(stream));
}

extern "C" int msda_linear_tf32_ws_ok(int N, int K) { return gemm::ws_ok(N, K) ? 1 : 0; }

extern "C" int msda_debug_gemm_timeline(unsigned long long *out, int n_words) {
    if (!out || n_words <= 0) return MSDA_E_BADARG;
    const size_t want = sizeof(unsigned long long) * (size_t)n_words, have = sizeof(unsigned long long) * 160 * gemm::kTlSlots;
    return (int)cudaMemcpyFromSymbol(out, gemm::g_ws2_timeline, want < have ? want : have);
}

extern "C" int msda_linear_tf32(const float *A, const float *W, const float *bias, int64_t M, int N, int K, float *C,
                                void *stream) {
    using namespace gemm;
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || K % BLOCK_K || N % 32 || N > kMaxN) return MSDA_E_BADARG;   // epilogue reads 32 TMEM columns at a time
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15u) return MSDA_E_BADARG;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15u)) return MSDA_E_BADARG;
    {
        static int ws = -1;
        if (ws < 0) { const char *e = getenv("MSDA_GEMM_WS"); ws = (e && e[0] == '0') ? 0 : 1; }
        if (ws && ws_ok(N, K))
            return use_ws2() && K <= 8 * BLOCK_K ? launch_ws2(A, W, bias, nullptr, M, N, K, 0, C, static_cast<cudaStream_t>(stream))
                             : launch_ws(A, W, bias, nullptr, M, N, K, 0, C, static_cast<cudaStream_t>(stream));
    }
    if (N > 256 && (N % 64)) return MSDA_E_BADARG;      // MMA N tiles are multiples of 16, W halves whole swizzle atoms
    const int w_box_rows = N / 2;                       // each CTA of the pair loads (and multicasts) half of W
    CUtensorMap map_a, map_w;
    if (!make_map(&map_a, A, M, K, BLOCK_M) || !make_map(&map_w, W, N, K, w_box_rows)) return MSDA_E_NODEVICE;
    Params p;
    p.w_box_rows = w_box_rows;
    p.M = M; p.N = N; p.K = K; p.bias = bias; p.C = C;
    const unsigned stage_bytes = (BLOCK_M + N) * BLOCK_K * 4;
    constexpr unsigned kXposeBytes = 4 * 32 * 33 * 4, kDynMax = 232448 - 1024;      // 227 KB minus the static part
    int stages = (int)((kDynMax - 1024u - kXposeBytes) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages > K / BLOCK_K) stages = K / BLOCK_K;
    if (stages < 1) return MSDA_E_BADARG;
    p.stages = stages;
    p.acc_bufs = (N <= 256) ? 2 : 1;
    const int need = N * p.acc_bufs;
    p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + kXposeBytes;
    // function attributes and SM counts are per DEVICE: cache them per ordinal (a process may drive several GPUs)
    constexpr int kMaxDev = 64;
    static std::atomic<int> sms_of[kMaxDev];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) dev = 0;
    int sms = sms_of[dev].load(std::memory_order_relaxed);
    if (sms == 0) {
        cudaError_t attr_err = cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024);
        if (attr_err != cudaSuccess) return (int)attr_err;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        sms_of[dev].store(sms, std::memory_order_relaxed);
    }
    const long long tiles = (M + BLOCK_M - 1) / BLOCK_M, pairs = (tiles + 1) / 2;
    const long long max_clusters = sms / 2;
    const unsigned grid = 2u * (unsigned)(pairs < max_clusters ? pairs : max_clusters);
    linear_tf32_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(map_a, map_w, p);
    g_msda_gemm_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}
