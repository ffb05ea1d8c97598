// method == 'bp' tail of the reference layer (/root/reference/models/tflct.py:164-175):
// ReplicationPad3d(2) -> conv3d with the 5x5x5 Laplacian-of-Gaussian -> zero time slice 0,
// and its transpose for the backward pass.  Unreachable from NlosPose (FeaturePropagation
// asserts mode == 'lct'), but a supported method of the layer.
//
// laplacian_tiled_kernel: one block owns RH rows x all N columns of a channel and marches along
// t.  Every input plane passes through shared memory once (a (RH+4) x (N+4) slab with the
// padding materialised, double-buffered with 4-byte cp.async) and is scattered into the five
// output planes it touches: 20 accumulators per thread (five planes x four consecutive w), ten
// 128-bit shared-memory loads and 500 FMAs per plane, weights straight from the constant bank.
// The gather form (laplacian_kernel, kept for volumes the tile shape does not fit) issued 125
// loads per output and took twice as long as the whole light-cone chain in front of it.
//
// Adjoint: with g zero-extended (and g[:, 0] = 0, the slice the forward zeroes), the transpose is
// gv[q] = sum_k W[k] g[q - k + 2] wherever q is not on the boundary of the volume -- the same
// kernel with zero padding and the taps flipped -- and on the boundary shell the clamped taps
// fold onto q: those voxels (under 5 % of a 128^3 volume) are redone by laplacian_adjoint_shell_kernel.
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

// ---------------------------------------------------------------------------
// tiled form
// ---------------------------------------------------------------------------
struct LapTiledParams {
    const float* in;
    float* out;
    int M, N;
    int rows;            // RH: output rows per block; blockDim.x = rows * N / 4
    int chunk;           // output planes per block along t
    int zero_pad;        // 0: replicate padding, output plane 0 forced to zero (forward)
                         // 1: zero padding, input plane 0 read as zero (interior of the adjoint; weights arrive flipped)
    StencilWeights w;
};

constexpr int kLapHalo = 2;
__host__ __device__ inline int lap_slab_stride(int N) { return N + 8; }                    // floats; keeps the rows 16-byte aligned
__host__ __device__ inline size_t lap_smem_bytes(int N, int rows) { return (size_t)2 * (rows + 2 * kLapHalo) * lap_slab_stride(N) * sizeof(float); }

// Where a thread's share of a slab comes from and goes to is the same for every plane: element e of thread t is slab
// slot i = t + e * blockDim.x, i.e. row r = i / (N + 4), column s = i % (N + 4), holding input (h0 + r - 2, s - 2) clamped
// into the plane.  Computed once per block (the division per element and plane was a third of the kernel's instructions).
template <int KE> struct LapSlots {
    int src[KE];          // offset inside a plane; -1: no such slot
    int dst[KE];          // offset inside the slab, bit 30 set when the source lies outside the plane (zero in the adjoint)
};
template <int KE>
__device__ __forceinline__ void lap_make_slots(const LapTiledParams& p, int h0, LapSlots<KE>& sl) {
    const int N = p.N, cols = N + 2 * kLapHalo, stride = lap_slab_stride(N), count = (p.rows + 2 * kLapHalo) * cols;
#pragma unroll
    for (int e = 0; e < KE; ++e) {
        const int i = threadIdx.x + e * blockDim.x;
        const int r = i / cols, s = i - r * cols;
        const int h = h0 + r - kLapHalo, w = s - kLapHalo;
        const bool inside = h >= 0 && h < N && w >= 0 && w < N;
        sl.src[e] = i < count ? clampi(h, N) * N + clampi(w, N) : -1;
        sl.dst[e] = (r * stride + s) | (inside ? 0 : (1 << 30));
    }
}
// plane `tv` (virtual index: outside [0, M) it is padding) of the channel -> slab[(rows + 4)][stride], columns s = w + 2
template <int KE>
__device__ __forceinline__ void lap_issue_plane(const LapTiledParams& p, const float* chan, float* slab, int tv, const LapSlots<KE>& sl) {
    const bool absent = p.zero_pad && (tv <= 0 || tv >= p.M);        // zero padding in t, and the slice the forward zeroes
    const float* plane = chan + (size_t)clampi(tv, p.M) * p.N * p.N;
    const unsigned base = (unsigned)__cvta_generic_to_shared(slab);
#pragma unroll
    for (int e = 0; e < KE; ++e) {
        if (sl.src[e] < 0) continue;
        const bool outside = (sl.dst[e] >> 30) & 1;
        const int nbytes = (absent || (p.zero_pad && outside)) ? 0 : 4;              // 0: the four bytes are zero-filled
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                     ::"r"(base + 4u * (unsigned)(sl.dst[e] & 0x3fffffff)), "l"(plane + sl.src[e]), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// one input plane (slot J of a group of five consecutive virtual planes) scattered into the five output planes it
// touches: output plane tv - a + 2 takes tap a; its accumulators sit at index (J - a + 2) mod 5
template <int J>
__device__ __forceinline__ void lap_accumulate(const LapTiledParams& p, const float* rowbase, int stride, float (&acc)[5][4]) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const float4 lo = *reinterpret_cast<const float4*>(rowbase + b * stride);
        const float4 hi = *reinterpret_cast<const float4*>(rowbase + b * stride + 4);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            constexpr int kFive = 5;
            const int slot = (J - a + 2 + kFive) % kFive;
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                const float wt = p.w.w[(a * 5 + b) * 5 + d];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[slot][q] = fmaf(wt, v[q + d], acc[slot][q]);
            }
        }
    }
}

template <int J, int KE>
__device__ __forceinline__ void lap_step(const LapTiledParams& p, const float* chan, float* out_chan, float* smem, float (&acc)[5][4],
                                         const LapSlots<KE>& sl, int tv, int tv_last, int h0, int hr, int w0, int t_begin, int t_end, int& buf) {
    if (tv > tv_last) return;                                            // block-uniform
    const int stride = lap_slab_stride(p.N), slab_floats = (p.rows + 2 * kLapHalo) * stride;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                     // plane tv has landed; everyone is done with the other buffer
    if (tv < tv_last) lap_issue_plane<KE>(p, chan, smem + (buf ^ 1) * slab_floats, tv + 1, sl);
    const bool absent = p.zero_pad && (tv <= 0 || tv >= p.M);
    if (!absent) lap_accumulate<J>(p, smem + buf * slab_floats + hr * stride + w0, stride, acc);
    buf ^= 1;
    // output plane tv - 2 has now seen all five of its input planes
    constexpr int done = (J + 3) % 5;
    const int t = tv - 2;
    if (t >= t_begin && t < t_end) {
        float4 r = make_float4(acc[done][0], acc[done][1], acc[done][2], acc[done][3]);
        if (!p.zero_pad && t == 0) r = make_float4(0.f, 0.f, 0.f, 0.f);  // tflct.py:175
        *reinterpret_cast<float4*>(out_chan + ((size_t)t * p.N + h0 + hr) * p.N + w0) = r;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[done][q] = 0.f;
}

// grid (N / rows, ceil(M / chunk), C); block rows * N / 4 threads; dynamic shared memory lap_smem_bytes(N, rows);
// KE >= ceil((rows + 4) * (N + 4) / blockDim.x)
template <int KE>
__global__ void __launch_bounds__(512, 2) laplacian_tiled_kernel(const LapTiledParams p) {
    extern __shared__ __align__(16) float lap_smem[];
    const int N = p.N, quads = N / 4;
    const int hr = threadIdx.x / quads, w0 = (threadIdx.x % quads) * 4;
    const int h0 = blockIdx.x * p.rows;
    const int t_begin = blockIdx.y * p.chunk, t_end = min(t_begin + p.chunk, p.M);
    const float* chan = p.in + (size_t)blockIdx.z * p.M * N * N;
    float* out_chan = p.out + (size_t)blockIdx.z * p.M * N * N;
    LapSlots<KE> sl;
    lap_make_slots<KE>(p, h0, sl);
    float acc[5][4];
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[s][q] = 0.f;
    const int tv_first = t_begin - 2, tv_last = t_end + 1;              // virtual planes the chunk's outputs read
    int buf = 0;
    lap_issue_plane<KE>(p, chan, lap_smem, tv_first, sl);
    for (int base = tv_first; base <= tv_last; base += 5) {
        lap_step<0, KE>(p, chan, out_chan, lap_smem, acc, sl, base + 0, tv_last, h0, hr, w0, t_begin, t_end, buf);
        lap_step<1, KE>(p, chan, out_chan, lap_smem, acc, sl, base + 1, tv_last, h0, hr, w0, t_begin, t_end, buf);
        lap_step<2, KE>(p, chan, out_chan, lap_smem, acc, sl, base + 2, tv_last, h0, hr, w0, t_begin, t_end, buf);
        lap_step<3, KE>(p, chan, out_chan, lap_smem, acc, sl, base + 3, tv_last, h0, hr, w0, t_begin, t_end, buf);
        lap_step<4, KE>(p, chan, out_chan, lap_smem, acc, sl, base + 4, tv_last, h0, hr, w0, t_begin, t_end, buf);
    }
}

// ---------------------------------------------------------------------------
// Boundary shell of the transpose: the voxels with t in {0, M-1}, h in {0, N-1} or w in {0, N-1}, enumerated compactly
// (t faces, then h faces, then w faces, each without the voxels of the faces before it) so that a warp is not held up by
// the one lane in thirty-two that sits on a face.
// ---------------------------------------------------------------------------
__global__ void laplacian_adjoint_shell_kernel(const float* __restrict__ gout, float* __restrict__ gvol, int C, int M, int N, StencilWeights sw) {
    const long long n_t = 2LL * N * N, n_h = 2LL * (M - 2) * N, n_w = 2LL * (M - 2) * (N - 2), per = n_t + n_h + n_w;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= per * C) return;
    const int c = (int)(i / per);
    long long j = i - c * per;
    int t, h, w;
    if (j < n_t) { t = j < (long long)N * N ? 0 : M - 1; j %= (long long)N * N; h = (int)(j / N); w = (int)(j % N); }
    else if (j < n_t + n_h) { j -= n_t; h = j < (long long)(M - 2) * N ? 0 : N - 1; j %= (long long)(M - 2) * N; t = 1 + (int)(j / N); w = (int)(j % N); }
    else { j -= n_t + n_h; w = j < (long long)(M - 2) * (N - 2) ? 0 : N - 1; j %= (long long)(M - 2) * (N - 2); t = 1 + (int)(j / (N - 2)); h = 1 + (int)(j % (N - 2)); }
    const float* g = gout + (size_t)c * M * N * N;
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
    gvol[(size_t)c * M * N * N + ((size_t)t * N + h) * N + w] = acc;
}

}  // namespace lct
