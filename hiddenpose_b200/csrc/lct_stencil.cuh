// method == 'bp' tail of the reference layer (/root/reference/models/tflct.py:164-175):
// ReplicationPad3d(2) -> conv3d with the 5x5x5 Laplacian-of-Gaussian -> zero time slice 0,
// and its transpose for the backward pass.  Unreachable from NlosPose
// (FeaturePropagation asserts mode == 'lct'), so these are plain gather kernels.
#pragma once

#include <cuda_runtime.h>

namespace lct {

struct StencilWeights { float w[125]; };

__device__ __forceinline__ int clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

// out[c,t,h,w] = sum_k lapw[kt,kh,kw] * vol[c, clamp(t+kt-2), clamp(h+kh-2), clamp(w+kw-2)];  out[c,0] = 0
__global__ void laplacian_kernel(const float* __restrict__ vol, float* __restrict__ out, int C, int M, int N, StencilWeights sw) {
    const size_t total = (size_t)C * M * N * N;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % N), h = (int)((i / N) % N), t = (int)((i / ((size_t)N * N)) % M);
    const size_t c = i / ((size_t)M * N * N);
    if (t == 0) { out[i] = 0.f; return; }
    const float* v = vol + c * (size_t)M * N * N;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        const int tt = clampi(t + a - 2, M);
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const int hh = clampi(h + b - 2, N);
            const float* rowp = v + ((size_t)tt * N + hh) * N;
#pragma unroll
            for (int d = 0; d < 5; ++d) acc += sw.w[(a * 5 + b) * 5 + d] * __ldg(rowp + clampi(w + d - 2, N));
        }
    }
    out[i] = acc;
}

// range of output coordinates p along one axis whose tap k reads input coordinate q (after clamping)
__device__ __forceinline__ void tap_range(int q, int k, int n, int& lo, int& hi) {
    lo = hi = q - k + 2;                       // interior: exactly one
    if (q == 0) { lo = 0; hi = 2 - k; }        // everything clamped onto the first sample
    if (q == n - 1) { hi = n - 1; lo = (q == 0) ? 0 : n + 1 - k; }
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
}

// gvol = (d out / d vol)^T gout
__global__ void laplacian_adjoint_kernel(const float* __restrict__ gout, float* __restrict__ gvol, int C, int M, int N, StencilWeights sw) {
    const size_t total = (size_t)C * M * N * N;
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % N), h = (int)((i / N) % N), t = (int)((i / ((size_t)N * N)) % M);
    const size_t c = i / ((size_t)M * N * N);
    const float* g = gout + c * (size_t)M * N * N;
    float acc = 0.f;
    for (int a = 0; a < 5; ++a) {
        int t0, t1; tap_range(t, a, M, t0, t1);
        if (t0 < 1) t0 = 1;                    // out[:, 0] is forced to zero: no gradient through it
        for (int b = 0; b < 5; ++b) {
            int h0, h1; tap_range(h, b, N, h0, h1);
            for (int d = 0; d < 5; ++d) {
                int w0, w1; tap_range(w, d, N, w0, w1);
                const float wt = sw.w[(a * 5 + b) * 5 + d];
                for (int tt = t0; tt <= t1; ++tt)
                    for (int hh = h0; hh <= h1; ++hh)
                        for (int ww = w0; ww <= w1; ++ww) acc += wt * __ldg(g + ((size_t)tt * N + hh) * N + ww);
            }
        }
    }
    gvol[i] = acc;
}

}  // namespace lct
