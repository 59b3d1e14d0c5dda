// msda_condinst.cuh -- CondInst dynamic mask head, the mask branch of UNINEXT's video / instance
// configs (uninext/models/ddetrs.py:488-598 `mask_heads_forward` + `dynamic_mask_with_coords`, :895-944 helpers).
//
// Reference formulation: for every selected instance a 3-layer MLP of 1x1 convolutions whose 169 weights come from the
// controller (10 -> 8 -> 8 -> 1 channels; input = 2 relative coordinates + the 8 mask-feature channels of the instance's
// image), evaluated at every pixel of the stride-8 mask feature map.  The reference materialises the input as
// [1, I * 10, H, W] with `repeat` (201 MB for 300 instances at 100 x 168) and runs three grouped convolutions with
// groups = I -- a shape cuDNN handles badly -- then upsamples with `aligned_bilinear`.
//
// Here: one thread per pixel keeps the pixel's 8 feature channels and its location in registers and loops over the
// instances of its image, whose parameters sit in shared memory (broadcast reads); nothing is materialised.
//   forward : logits[i, y, x]                                         (ddetrs.py:493-505, 523-566)
//   backward: grad_feats (accumulated over the instance loop in registers: no atomics), grad_params / grad_ref
//             (block reduction over pixels, one atomicAdd per value and block).
//   aligned_bilinear forward / backward (ddetrs.py:921-942): replicate-pad, align_corners bilinear x factor, shift by
//   factor / 2 -- closed form, gather on both passes.
// Parameter layout of one instance (parse_dynamic_params, ddetrs.py:895-918): w1[8][10] | w2[8][8] | w3[8] | b1[8] | b2[8] | b3,
// input channel order (rel_x, rel_y, feat_0..7) with rel = reference_point - (pixel * stride + stride / 2)  (:533-541, :944-958).
#pragma once

#include "msda_common.cuh"

namespace msda {

constexpr int kCiFeat = 8;                  // mask feature channels (hidden_dim / 32)
constexpr int kCiCh = 8;                    // dynamic_mask_channels
constexpr int kCiIn = kCiFeat + 2;
constexpr int kCiParams = kCiCh * kCiIn + kCiCh * kCiCh + kCiCh + kCiCh + kCiCh + 1;      // 169
constexpr int kCiW1 = 0, kCiW2 = kCiCh * kCiIn, kCiW3 = kCiW2 + kCiCh * kCiCh, kCiB1 = kCiW3 + kCiCh, kCiB2 = kCiB1 + kCiCh,
              kCiB3 = kCiB2 + kCiCh;
constexpr int kCiThreads = 256;
constexpr int kCiChunk = 16;                // instances whose parameters are staged in shared memory at a time

struct CiPixel {
    float x[kCiIn];                         // filled per instance: rel coords + features
};

__device__ __forceinline__ float ci_forward(const float *__restrict__ p, const float (&x)[kCiIn], float (&h1)[kCiCh],
                                            float (&h2)[kCiCh]) {
#pragma unroll
    for (int o = 0; o < kCiCh; ++o) {
        float a = p[kCiB1 + o];
#pragma unroll
        for (int c = 0; c < kCiIn; ++c) a = fmaf(p[kCiW1 + o * kCiIn + c], x[c], a);
        h1[o] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int o = 0; o < kCiCh; ++o) {
        float a = p[kCiB2 + o];
#pragma unroll
        for (int c = 0; c < kCiCh; ++c) a = fmaf(p[kCiW2 + o * kCiCh + c], h1[c], a);
        h2[o] = fmaxf(a, 0.f);
    }
    float out = p[kCiB3];
#pragma unroll
    for (int c = 0; c < kCiCh; ++c) out = fmaf(p[kCiW3 + c], h2[c], out);
    return out;
}

// grid: (pixel tiles, instance chunks, images).  feats [N, 8, H*W]; params [I, 169]; refs [I, 2] (pixels of the input image);
// inst_start [N + 1] (instances of image b are [inst_start[b], inst_start[b + 1])); logits [I, H*W].
__global__ void __launch_bounds__(kCiThreads)
condinst_fwd(const float *__restrict__ feats, const float *__restrict__ params, const float *__restrict__ refs,
             const int *__restrict__ inst_start, int HW, int W, int stride, int rel_coord, float *__restrict__ logits)
{
    __shared__ float sp[kCiChunk][kCiParams + 3];
    const int b = blockIdx.z;
    const int i0 = inst_start[b] + blockIdx.y * kCiChunk, i1 = min(inst_start[b + 1], i0 + kCiChunk);
    if (i0 >= i1) return;
    for (int t = threadIdx.x; t < (i1 - i0) * (kCiParams + 2); t += kCiThreads) {
        const int k = t / (kCiParams + 2), j = t - k * (kCiParams + 2);
        sp[k][j] = j < kCiParams ? params[(size_t)(i0 + k) * kCiParams + j] : refs[(size_t)(i0 + k) * 2 + (j - kCiParams)];
    }
    __syncthreads();
    const int px = blockIdx.x * kCiThreads + threadIdx.x;
    if (px >= HW) return;
    float x[kCiIn];
#pragma unroll
    for (int c = 0; c < kCiFeat; ++c) x[2 + c] = __ldg(feats + ((size_t)b * kCiFeat + c) * HW + px);
    const float lx = (float)((px % W) * stride + stride / 2), ly = (float)((px / W) * stride + stride / 2);
    for (int k = 0; k < i1 - i0; ++k) {
        x[0] = rel_coord ? sp[k][kCiParams] - lx : 0.f;
        x[1] = rel_coord ? sp[k][kCiParams + 1] - ly : 0.f;
        float h1[kCiCh], h2[kCiCh];
        logits[(size_t)(i0 + k) * HW + px] = ci_forward(sp[k], x, h1, h2);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

// grid: (pixel tiles, 1, images): a block walks ALL instances of its image, so grad_feats needs no atomics.
// grad_params [I, 169] and grad_refs [I, 2] must be zero on entry (block partial sums are added atomically).
__global__ void __launch_bounds__(kCiThreads)
condinst_bwd(const float *__restrict__ grad_logits, const float *__restrict__ feats, const float *__restrict__ params,
             const float *__restrict__ refs, const int *__restrict__ inst_start, int HW, int W, int stride, int rel_coord,
             float *__restrict__ grad_feats, float *__restrict__ grad_params, float *__restrict__ grad_refs)
{
    __shared__ float sp[kCiChunk][kCiParams + 3];
    __shared__ float red[kCiThreads / 32][kCiParams + 3];
    const int b = blockIdx.z;
    const int ib0 = inst_start[b], ib1 = inst_start[b + 1];
    const int px = blockIdx.x * kCiThreads + threadIdx.x;
    const bool live = px < HW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x[kCiIn], gf[kCiFeat];
#pragma unroll
    for (int c = 0; c < kCiFeat; ++c) { x[2 + c] = live ? __ldg(feats + ((size_t)b * kCiFeat + c) * HW + px) : 0.f; gf[c] = 0.f; }
    const int pxs = live ? px : 0;
    const float lx = (float)((pxs % W) * stride + stride / 2), ly = (float)((pxs / W) * stride + stride / 2);

    for (int i0 = ib0; i0 < ib1; i0 += kCiChunk) {
        const int n = min(kCiChunk, ib1 - i0);
        __syncthreads();
        for (int t = threadIdx.x; t < n * (kCiParams + 2); t += kCiThreads) {
            const int k = t / (kCiParams + 2), j = t - k * (kCiParams + 2);
            sp[k][j] = j < kCiParams ? params[(size_t)(i0 + k) * kCiParams + j] : refs[(size_t)(i0 + k) * 2 + (j - kCiParams)];
        }
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            const float *p = sp[k];
            x[0] = rel_coord ? p[kCiParams] - lx : 0.f;
            x[1] = rel_coord ? p[kCiParams + 1] - ly : 0.f;
            float h1[kCiCh], h2[kCiCh];
            ci_forward(p, x, h1, h2);
            const float go = live ? __ldg(grad_logits + (size_t)(i0 + k) * HW + px) : 0.f;
            // layer 3:  out = b3 + w3 . h2
            float d2[kCiCh], d1[kCiCh], dx[kCiIn];
#pragma unroll
            for (int c = 0; c < kCiCh; ++c) d2[c] = h2[c] > 0.f ? go * p[kCiW3 + c] : 0.f;      // grad wrt layer-2 pre-activation
#pragma unroll
            for (int c = 0; c < kCiCh; ++c) {
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < kCiCh; ++o) a = fmaf(p[kCiW2 + o * kCiCh + c], d2[o], a);
                d1[c] = h1[c] > 0.f ? a : 0.f;                                                  // grad wrt layer-1 pre-activation
            }
#pragma unroll
            for (int c = 0; c < kCiIn; ++c) {
                float a = 0.f;
#pragma unroll
                for (int o = 0; o < kCiCh; ++o) a = fmaf(p[kCiW1 + o * kCiIn + c], d1[o], a);
                dx[c] = a;
            }
#pragma unroll
            for (int c = 0; c < kCiFeat; ++c) gf[c] += dx[2 + c];
            // parameter / reference-point gradients: sum over the pixels of this block.  Warp sums land in `red`,
            // warp 0 folds them and issues one atomicAdd per value.
            auto put = [&](int j, float v) { v = warp_sum(v); if (lane == 0) red[warp][j] = v; };
#pragma unroll
            for (int o = 0; o < kCiCh; ++o) {
#pragma unroll
                for (int c = 0; c < kCiIn; ++c) put(kCiW1 + o * kCiIn + c, d1[o] * x[c]);
#pragma unroll
                for (int c = 0; c < kCiCh; ++c) put(kCiW2 + o * kCiCh + c, d2[o] * h1[c]);
                put(kCiW3 + o, go * h2[o]);
                put(kCiB1 + o, d1[o]);
                put(kCiB2 + o, d2[o]);
            }
            put(kCiB3, go);
            put(kCiParams, rel_coord ? dx[0] : 0.f);
            put(kCiParams + 1, rel_coord ? dx[1] : 0.f);
            __syncthreads();
            for (int j = threadIdx.x; j < kCiParams + 2; j += kCiThreads) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kCiThreads / 32; ++w) s += red[w][j];
                if (s != 0.f) {
                    if (j < kCiParams) atomicAdd(grad_params + (size_t)(i0 + k) * kCiParams + j, s);
                    else atomicAdd(grad_refs + (size_t)(i0 + k) * 2 + (j - kCiParams), s);
                }
            }
            __syncthreads();
        }
    }
    if (live) {
#pragma unroll
        for (int c = 0; c < kCiFeat; ++c) grad_feats[((size_t)b * kCiFeat + c) * HW + px] = gf[c];
    }
}

// ---- aligned_bilinear (ddetrs.py:921-942): out[Y, X] of size (f*h, f*w) samples the replicate-padded input at
// ((Y - f/2)^+ / f, (X - f/2)^+ / f) with align_corners weights.  src index pair + fraction for one output coordinate:
__device__ __forceinline__ void ab_src(int o, int f, int n, int &i0, int &i1, float &fr) {
    const int p = max(o - f / 2, 0);
    i0 = p / f;
    fr = (float)(p - i0 * f) / (float)f;
    i1 = min(i0 + 1, n - 1);
}

__global__ void __launch_bounds__(256)
aligned_bilinear_fwd(const float *__restrict__ in, long long planes, int h, int w, int f, float *__restrict__ out) {
    const int oh = h * f, ow = w * f;
    const long long total = planes * oh * ow;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(t % ow), Y = (int)((t / ow) % oh);
        const long long pl = t / ((long long)ow * oh);
        int y0, y1, x0, x1; float fy, fx;
        ab_src(Y, f, h, y0, y1, fy);
        ab_src(X, f, w, x0, x1, fx);
        const float *src = in + pl * h * w;
        const float top = (1.f - fx) * __ldg(src + y0 * w + x0) + fx * __ldg(src + y0 * w + x1);
        const float bot = (1.f - fx) * __ldg(src + y1 * w + x0) + fx * __ldg(src + y1 * w + x1);
        out[t] = (1.f - fy) * top + fy * bot;
    }
}

// gather form of the transpose: input pixel (y, x) collects from every output whose source pair contains it.
__global__ void __launch_bounds__(256)
aligned_bilinear_bwd(const float *__restrict__ gout, long long planes, int h, int w, int f, float *__restrict__ gin) {
    const int oh = h * f, ow = w * f;
    const long long total = planes * h * w;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % w), y = (int)((t / w) % h);
        const long long pl = t / ((long long)w * h);
        const float *g = gout + pl * oh * ow;
        // outputs whose (i0, i1) can contain y: source positions in [y - 1, y + 1) -> Y in [f*(y-1) + f/2, f*(y+1) + f/2), plus
        // the clamped head [0, f/2) for y == 0
        const int Ya = y == 0 ? 0 : max(0, f * (y - 1) + f / 2), Yb = min(oh, f * (y + 1) + f / 2);
        const int Xa = x == 0 ? 0 : max(0, f * (x - 1) + f / 2), Xb = min(ow, f * (x + 1) + f / 2);
        float acc = 0.f;
        for (int Y = Ya; Y < Yb; ++Y) {
            int y0, y1; float fy;
            ab_src(Y, f, h, y0, y1, fy);
            const float wy = (y0 == y ? 1.f - fy : 0.f) + (y1 == y ? fy : 0.f);
            if (wy == 0.f) continue;
            for (int X = Xa; X < Xb; ++X) {
                int x0, x1; float fx;
                ab_src(X, f, w, x0, x1, fx);
                const float wx = (x0 == x ? 1.f - fx : 0.f) + (x1 == x ? fx : 0.f);
                if (wx != 0.f) acc = fmaf(wy * wx, __ldg(g + (size_t)Y * ow + X), acc);
            }
        }
        gin[t] = acc;
    }
}

}  // namespace msda
