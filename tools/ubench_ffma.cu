// ubench_ffma.cu -- fp32 FMA issue rate on sm_100a: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), with and without a
// multiplier shared by consecutive instructions (operand-reuse cache / scalar-broadcast operand).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_ffma tools/ubench_ffma.cu && ./ubench_ffma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ float lo(f2 a) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); return x + y; }

constexpr int CH = 8, IT = 512;

template <int MODE>
__global__ void k(const float *in, float *out, long long *cyc) {
    float a[CH], b[CH], w[CH], u[CH];                // w: per-thread values (vector registers); u: warp-uniform (uniform registers)
    f2 pa[CH], pb[CH], pw[CH], pu[CH];
    for (int i = 0; i < CH; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; w[i] = in[threadIdx.x + 64 + i]; u[i] = in[64 + i];
        pa[i] = mk(a[i], b[i]); pb[i] = mk(b[i], a[i]); pw[i] = mk(w[i], w[(i + 1) % CH]); pu[i] = mk(u[i], u[(i + 1) % CH]);
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int r = 0; r < CH; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) a[i] = fmaf(w[i], b[i], a[i]);                      // three distinct registers
                if (MODE == 1) a[i] = fmaf(w[r], b[i], a[i]);                      // multiplier shared by 8 consecutive FFMAs
                if (MODE == 2) pa[i] = fma2(pw[i], pb[i], pa[i]);                  // three distinct pairs
                if (MODE == 3) pa[i] = fma2(pw[r], pb[i], pa[i]);                  // pair multiplier shared
                if (MODE == 4) pa[i] = fma2(mk(w[r], w[r]), pb[i], pa[i]);         // scalar-broadcast multiplier shared
                if (MODE == 5) a[i] = fmaf(a[i], 1.0001f, b[i]);                   // immediate multiplier
                if (MODE == 6) a[i] = fmaf(u[r], b[i], a[i]);                      // warp-uniform multiplier
                if (MODE == 7) pa[i] = fma2(pu[r], pb[i], pa[i]);                  // warp-uniform pair multiplier
                if (MODE == 8) pa[i] = fma2(mk(u[r], u[r]), pb[i], pa[i]);         // warp-uniform scalar-broadcast multiplier
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < CH; ++i) s += a[i] + lo(pa[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, const float *in, float *out, long long *cyc) {
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(in, out, cyc);
        k<MODE><<<148, warps * 32>>>(in, out, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double per = (double)c / ((double)IT * CH * CH * (warps / 4));
        printf("%-46s warps/SMSP %d: %.2f cycles per warp-instruction per scheduler\n", name, warps / 4, per);
    }
}

int main() {
    float *in, *out; long long *cyc;
    cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("FFMA  3 distinct registers", in, out, cyc);
    run<1>("FFMA  shared multiplier (reuse)", in, out, cyc);
    run<5>("FFMA  immediate multiplier", in, out, cyc);
    run<2>("FFMA2 3 distinct pairs", in, out, cyc);
    run<3>("FFMA2 shared pair multiplier", in, out, cyc);
    run<4>("FFMA2 shared scalar-broadcast multiplier", in, out, cyc);
    run<6>("FFMA  warp-uniform multiplier", in, out, cyc);
    run<7>("FFMA2 warp-uniform pair multiplier", in, out, cyc);
    run<8>("FFMA2 warp-uniform scalar-broadcast multiplier", in, out, cyc);
    return cudaGetLastError() != cudaSuccess;
}
