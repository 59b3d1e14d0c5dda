// Micro-benchmark: what does it cost to ADD a 128-byte fp32 row into a shared-memory accumulator (the candidate for
// privatising grad_value's coarse levels inside one SM), against the L2 path (red.global.add.v4.f32)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_smem_rmw tools/ubench_smem_rmw.cu
// Access shape = the backward's: 8 lanes x 16 B cover one row, 4 random rows per warp instruction.
// sm_100a has NO native fp32 shared-memory atomic add (atomicAdd(float*) on shared compiles to an
// LDS + FADD + ATOMS.CAST.SPIN loop); candidates:
//   f32x4   : 4 x atomicAdd(float) per lane                      (ATOMS.CAST.SPIN loop, what CUDA C++ gives)
//   cas64x2 : LDS.64 + ATOMS.CAS.64 loop, 2 per lane
//   cas128  : LDS.128 + ATOMS.CAS.128 loop, 1 per lane
//   rmw     : LDS.128 + 4 FADD + STS.128, NOT atomic (upper bound: what an ownership scheme could reach)
//   int     : 4 x ATOMS.ADD (u32, native) per lane                (what a native fp32 atomic would cost, if it existed)
//   global  : red.global.add.v4.f32 into an L2-resident table     (the path the kernel uses today)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Mode { F32X4 = 0, CAS64X2 = 1, CAS128 = 2, RMW = 3, INT = 4, GLOBAL = 5 };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void add_cas128(float4 *p, float4 v) {
    const unsigned a = smem_u32(p);
    float4 cur = *p;
    while (true) {
        const float4 nw = make_float4(cur.x + v.x, cur.y + v.y, cur.z + v.z, cur.w + v.w);
        const unsigned long long c0 = ((unsigned long long)__float_as_uint(cur.y) << 32) | __float_as_uint(cur.x);
        const unsigned long long c1 = ((unsigned long long)__float_as_uint(cur.w) << 32) | __float_as_uint(cur.z);
        const unsigned long long n0 = ((unsigned long long)__float_as_uint(nw.y) << 32) | __float_as_uint(nw.x);
        const unsigned long long n1 = ((unsigned long long)__float_as_uint(nw.w) << 32) | __float_as_uint(nw.z);
        unsigned long long o0, o1;
        asm volatile("{\n\t.reg .b128 c, n, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 n, {%5, %6};\n\t"
                     "atom.shared.cas.b128 o, [%2], c, n;\n\tmov.b128 {%0, %1}, o;\n\t}"
                     : "=l"(o0), "=l"(o1) : "r"(a), "l"(c0), "l"(c1), "l"(n0), "l"(n1) : "memory");
        if (o0 == c0 && o1 == c1) break;
        cur.x = __uint_as_float((unsigned)o0); cur.y = __uint_as_float((unsigned)(o0 >> 32));
        cur.z = __uint_as_float((unsigned)o1); cur.w = __uint_as_float((unsigned)(o1 >> 32));
    }
}

__device__ __forceinline__ void add_cas64(float2 *p, float2 v) {
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p);
    unsigned long long cur = *q;
    while (true) {
        const float x = __uint_as_float((unsigned)cur) + v.x, y = __uint_as_float((unsigned)(cur >> 32)) + v.y;
        const unsigned long long nw = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
        const unsigned long long old = atomicCAS(q, cur, nw);
        if (old == cur) break;
        cur = old;
    }
}

template <int MODE, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) rowadd(float *gtable, unsigned rows, int iters, float *out, unsigned seed) {
    extern __shared__ __align__(16) float acc[];          // rows x 32 floats
    const unsigned srows = MODE == GLOBAL ? 8u : rows;
    for (unsigned i = threadIdx.x; i < srows * 32; i += THREADS) acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane % 8, grp = lane / 8;
    unsigned s = seed + (blockIdx.x * THREADS + threadIdx.x) / 8 * 2654435761u + grp * 40503u;
    for (int it = 0; it < iters; ++it) {
        unsigned r[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = s * 1664525u + 1013904223u;
            r[u] = (unsigned)(((unsigned long long)(s >> 4) * rows) >> 28);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float x = (float)(r[u] & 3u);
            float *row = acc + (size_t)r[u] * 32 + sub * 4;
            if (MODE == F32X4) {
                atomicAdd(row, x); atomicAdd(row + 1, 1.f); atomicAdd(row + 2, 2.f); atomicAdd(row + 3, 3.f);
            } else if (MODE == CAS64X2) {
                add_cas64(reinterpret_cast<float2 *>(row), make_float2(x, 1.f));
                add_cas64(reinterpret_cast<float2 *>(row + 2), make_float2(2.f, 3.f));
            } else if (MODE == CAS128) {
                add_cas128(reinterpret_cast<float4 *>(row), make_float4(x, 1.f, 2.f, 3.f));
            } else if (MODE == RMW) {
                float4 v;
                const unsigned a = smem_u32(row);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
                v.x += x; v.y += 1.f; v.z += 2.f; v.w += 3.f;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            } else if (MODE == INT) {
                unsigned *ri = reinterpret_cast<unsigned *>(row);
                atomicAdd(ri, r[u] & 3u); atomicAdd(ri + 1, 1u); atomicAdd(ri + 2, 2u); atomicAdd(ri + 3, 3u);
            } else {
                float *a = gtable + ((size_t)r[u] * 8 + sub) * 4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(1.f), "f"(2.f), "f"(3.f) : "memory");
            }
        }
    }
    __syncthreads();
    float t = 0.f;
    for (unsigned i = threadIdx.x; i < srows * 32; i += THREADS) t += acc[i];
    if (t == 123.456f) out[0] = t;
}

// LDG.256 gather (sm_100: ld.global.v8.f32): 4 lanes x 32 B cover one 128-byte row, 8 rows per warp instruction.
template <int UNROLL>
__global__ void __launch_bounds__(256) gather256(const float *__restrict__ table, unsigned rows, int iters, float *out,
                                                 unsigned seed) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % 4, grp = lane / 4;
    unsigned s = seed + (blockIdx.x * blockDim.x + threadIdx.x) / 4 * 2654435761u + grp * 40503u;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float v[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            s = s * 1664525u + 1013904223u;
            const unsigned r = (unsigned)(((unsigned long long)(s >> 4) * rows) >> 28);
            const float *p = table + (size_t)r * 32 + sub * 8;
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                           "=f"(v[u][6]), "=f"(v[u][7]) : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += v[u][e];
    }
    if (acc == 123.456f) out[0] = acc;
}

void run_gather256(const float *tab, size_t table_bytes, int ctas_per_sm) {
    constexpr int UNROLL = 4;
    const unsigned rows = (unsigned)(table_bytes / 128);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; cudaMalloc(&out, 4);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather256<UNROLL><<<sms * ctas_per_sm, 256>>>(tab, rows, 10, out, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    gather256<UNROLL><<<sms * ctas_per_sm, 256>>>(tab, rows, iters, out, 7u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ldg_per_sm = (double)ctas_per_sm * 8 * iters * UNROLL;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("LDG.256 gather, 4 lanes/row  table %8.1f MB  ctas/sm %d : %6.2f cyc per 1024-B LDG.256 per SM  %7.1f B/clk/SM  %6.2f TB/s\n",
           table_bytes / 1e6, ctas_per_sm, cycles / ldg_per_sm, 1024.0 * ldg_per_sm / cycles,
           ldg_per_sm * sms * 1024.0 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

// red.global.add.v4.f32 from a SUBSET of the SMs: does the 6.4 TB/s ceiling belong to the L2 (stays when fewer SMs
// issue) or to each SM's path into the crossbar (scales with the SM count)?  One CTA per SM is forced with a large
// dynamic shared-memory request.
void run_red_subset(float *gtable, unsigned grows, int n_ctas, const char *tag) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto kern = rowadd<GLOBAL, 8, 512>;
    const size_t smem = 120 * 1024;                      // > half of the SM: at most one CTA per SM
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float *out; cudaMalloc(&out, 4);
    const int iters = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<n_ctas, 512, smem>>>(gtable, grows, 4, out, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<n_ctas, 512, smem>>>(gtable, grows, iters, out, 7u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows_per_cta = 64.0 * iters * 8;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("red.v4 %-10s %3d CTAs (1 per SM) x 512 thr, %8u rows : %6.2f cyc per 128-B row-add per SM  %6.2f B/clk/SM  %6.2f TB/s chip\n",
           tag, n_ctas, grows, cycles / rows_per_cta, 128.0 * rows_per_cta / cycles, rows_per_cta * n_ctas * 128.0 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

template <int MODE, int THREADS>
void run(const char *name, unsigned rows, int ctas_per_sm, float *gtable, unsigned grows) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    constexpr int UNROLL = 8;
    auto kern = rowadd<MODE, UNROLL, THREADS>;
    const size_t smem = MODE == GLOBAL ? 1024 : (size_t)rows * 128;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float *out; cudaMalloc(&out, 4);
    const int iters = MODE == F32X4 ? 100 : 400;
    const unsigned r = MODE == GLOBAL ? grows : rows;
    const unsigned srows = MODE == GLOBAL ? 8 : rows;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<sms * ctas_per_sm, THREADS, smem>>>(gtable, MODE == GLOBAL ? grows : srows, 4, out, 1u);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%-10s launch failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return; }
    cudaEventRecord(e0);
    kern<<<sms * ctas_per_sm, THREADS, smem>>>(gtable, r, iters, out, 7u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows_per_sm = (double)ctas_per_sm * (THREADS / 8) * iters * UNROLL;     // 128-byte row-adds per SM
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-10s rows %5u (%6.1f KB)  ctas/sm %d x %4d thr : %7.2f cyc per 128-B row-add per SM   %6.2f TB/s chip payload\n",
           name, r, r * 128 / 1e3, ctas_per_sm, THREADS, cycles / rows_per_sm, rows_per_sm * sms * 128.0 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    float *gtab; const size_t gb = 46u << 20;
    cudaMalloc(&gtab, gb); cudaMemset(gtab, 0, gb);
    const unsigned grows = (unsigned)(gb / 128);
    // (a) the cfg2 coarse window: levels 2+3 of one (b, m) slab = 1050 + 273 rows = 169 KB -> one CTA per SM
    for (unsigned rows : {1323u, 273u}) {
        run<F32X4, 512>("f32x4", rows, 1, gtab, grows);
        run<CAS64X2, 512>("cas64x2", rows, 1, gtab, grows);
        run<CAS128, 512>("cas128", rows, 1, gtab, grows);
        run<RMW, 512>("rmw", rows, 1, gtab, grows);
        run<INT, 512>("int", rows, 1, gtab, grows);
        run<CAS128, 1024>("cas128", rows, 1, gtab, grows);
        run<RMW, 1024>("rmw", rows, 1, gtab, grows);
    }
    // (b) level 3 only (35 KB) leaves room for several CTAs per SM
    run<CAS128, 256>("cas128", 273, 4, gtab, grows);
    run<RMW, 256>("rmw", 273, 4, gtab, grows);
    run<F32X4, 256>("f32x4", 273, 4, gtab, grows);
    // (c) the L2 path for comparison
    run<GLOBAL, 256>("global", 8, 2, gtab, grows);
    run<GLOBAL, 256>("global", 8, 4, gtab, grows);
    run<GLOBAL, 512>("global", 8, 2, gtab, grows);
    for (int n : {18, 37, 74, 111, 148}) run_red_subset(gtab, grows, n, "48MB");
    for (int n : {37, 148}) run_red_subset(gtab, 16u * 1323u, n, "hot-2.7MB");
    for (size_t sz : {(size_t)32 << 10, (size_t)128 << 10, (size_t)8 << 20, (size_t)46 << 20})
        for (int c : {2, 4}) run_gather256(gtab, sz, c);
    // (d) L2 path, hot set the size of the coarse levels of all 16 slabs (16 x 169 KB = 2.7 MB): contention on few lines
    run<GLOBAL, 256>("global-hot", 8, 4, gtab, 16u * 1323u);
    return 0;
}
