// On-device construction of the inverse light-cone filter (row f3 of SURVEY.md section 8).
//
// The reference builds it on the host: psf = definePsf(...) (utils/helper.py:72-125),
// fpsf = np.fft.fftn(psf), invpsf = conj(fpsf) / (1/snr + |fpsf|^2) (models/tflct.py:55-65) --
// 1.7 s / 11 s / 46 s for the BASELINE shapes.  The PSF has one voxel per (y, x) column (two on an
// exact tie), so its DFT along t is an analytic phase; here the host only finds the PSF support
// (operators.psf_support) and the GPU does the rest with the same register-resident FFT stages the
// data path uses:
//   psf_planes   P[kt][y][x] = val * w_2M^(kt z(y,x))                       (kt = 0..M)
//   ColumnFft    full-length FFT along y, in place, lanes along x
//   RowFftWiener full-length FFT along x -> Wiener (or conj for 'bp') -> scaled filter value, written
//                straight into the layout the data-path kernels read.
#pragma once

#include "lct_chain.cuh"

namespace lct {

struct FilterBuildParams {
    int M, N;                      // L = 2N
    float2* planes;                // (M+1, L, L) scratch, natural order
    float2* out;                   // final filter
    int fused_layout;              // 1: [kt][kw/2][plane row][kw&1]   0: [kt][kh][kw]   2: quarter [kt][kh <= N][kw <= N]   3: half rows [kt][kh <= N][kw]
    float inv_snr, scale;
    int conj_only;                 // method == 'bp'
};

// one thread per PSF voxel and plane
__global__ void psf_planes_kernel(float2* __restrict__ planes, const int* __restrict__ z, const int* __restrict__ flat,
                                  int count, float val, int M, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, kt = blockIdx.y;
    if (i >= count) return;
    const int idx = (int)(((long long)kt * z[i]) % (2 * M)) * (kTwN / (2 * M));
    const float2 w = g_tw[idx];
    float2* dst = planes + (size_t)kt * L * L + flat[i];
    atomicAdd(&dst->x, val * w.x);          // two voxels share a column only on exact ties
    atomicAdd(&dst->y, val * w.y);
}

// Full-length forward FFT along the strided axis of one (kt) plane, in place, CT columns per block.
template <class P, int CT_> struct ColumnFft {
    static constexpr int L = P::L, CT = CT_, kThreads = P::TL * CT;
    static constexpr size_t kSmem = (size_t)L * CT * sizeof(float2);
    static_assert(P::S == 2 || P::S == 3, "two or three stages");
};

template <class P, int CT> __global__ void __launch_bounds__(P::TL * CT)
column_fft_kernel(float2* __restrict__ planes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* zs = reinterpret_cast<float2*>(smem_raw);
    constexpr int L = P::L;
    const int tid = threadIdx.x, col = tid % CT, tau = tid / CT;
    float2* base = planes + (size_t)blockIdx.y * L * L + blockIdx.x * CT + col;
    auto ld_s = [&](int pos, int) { return zs[pos * CT + col]; };
    auto st_s = [&](int pos, int, float2 v) { zs[pos * CT + col] = v; };
    fwd_stage<P, 0, false, TwConst>(tau, [&](int pos, int) { return base[(size_t)pos * L]; }, st_s);
    __syncthreads();
    if constexpr (P::S == 3) {
        fwd_stage<P, 1, false, TwConst>(tau, ld_s, st_s);
        __syncthreads();
    }
    // every input of this tile is in shared memory by now, so writing the tile back in place is safe
    fwd_stage<P, P::S - 1, false, TwConst>(tau, ld_s,
        [&](int pos, int slot, float2 v) { base[(size_t)P::template freq_of<P::S - 1>(pos, slot) * L] = v; });
}

// Full-length forward FFT along the contiguous axis + Wiener formula; RB lines per block, warp-synchronous.
template <class P, class PH, int RB> __global__ void __launch_bounds__(P::TL * RB)
row_fft_wiener_kernel(FilterBuildParams p) {
    static_assert(P::S == 2 && 32 % P::TL == 0, "two-stage line plan with the line's threads in one warp");
    constexpr int L = P::L, N = L / 2, RS = L + P::R0 + ((P::TL < 16) ? 8 : 0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, tau = tid % P::TL, rl = tid / P::TL;
    const int kh = blockIdx.x * RB + rl, kt = blockIdx.y;
    float2* zs = reinterpret_cast<float2*>(smem_raw) + rl * RS;
    const float2* row = p.planes + ((size_t)kt * L + kh) * L;
    auto padpos = [](int pos) { return pos + pos / P::st(0); };
    fwd_stage<P, 0, false, TwGlobal>(tau,
        [&](int pos, int) { return row[pos]; },
        [&](int pos, int, float2 v) { zs[padpos(pos)] = v; });
    __syncwarp();
    const int plane_row = PH::freq_to_pos(kh);                 // where the data path keeps H-frequency kh
    fwd_stage<P, 1, false, TwGlobal>(tau,
        [&](int pos, int) { return zs[padpos(pos)]; },
        [&](int pos, int slot, float2 f) {
            const int kw = P::template freq_of<1>(pos, slot);
            float2 w = make_float2(f.x, -f.y);                 // conj(fpsf)                      tflct.py:60,62
            if (!p.conj_only) {
                const float den = p.inv_snr + f.x * f.x + f.y * f.y;
                w.x /= den;
                w.y /= den;
            }
            w.x *= p.scale;
            w.y *= p.scale;
            if (p.fused_layout == 2) {
                if (kh <= N && kw <= N) p.out[((size_t)kt * (N + 1) + kh) * (N + 1) + kw] = w;
                return;
            }
            if (p.fused_layout == 3) {
                if (kh <= N) p.out[((size_t)kt * (N + 1) + kh) * L + kw] = w;
                return;
            }
            const size_t o = p.fused_layout ? ((((size_t)kt * N + (kw >> 1)) * L + plane_row) * 2 + (kw & 1))
                                            : (((size_t)kt * L + kh) * L + kw);
            p.out[o] = w;
        });
}

}  // namespace lct
