// normalize_feature (/root/reference/models/feature_propagation.py:273-286): the op NlosPose applies
// right after the LCT (NlosPose.py:54).  Per (batch, channel) volume:
//     out = (x - min(x)) / (max(x - min(x)) + 1e-15) * 10
// (the nn.ReLU() on line 274 of the reference is computed and discarded, so negatives reach the min).
// The reference spends ~9 full passes over the volume on this (min, sub, max, div, mul + temporaries);
// here the per-channel min / max (with their positions, needed by the backward pass) are either taken
// from the LCT's last kernel, which reduces them while it writes the volume, or found by one reduction
// pass, and one more pass applies the affine map.
//
// min/max are carried as 64-bit keys  (order-preserving float bits << 32 | position)  reduced with
// atomicMin: the max side stores the complemented key.  Ties resolve to the lowest position.
// NaN wins both reductions, as it does in torch.min / torch.max (a diverged volume must turn the whole
// normalised output to NaN, as in the reference, not to finite garbage): it is keyed below -inf on the
// min side and above +inf on the max side, and both keys decode to a NaN.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lct {

__device__ __forceinline__ unsigned int float_order_key(float v) {
    const unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);          // monotone: larger float -> larger key
}
__device__ __forceinline__ float float_from_order_key(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long min_key(float v, unsigned int pos) {
    const unsigned int k = (v != v) ? 0u : float_order_key(v);                 // key 0 decodes to a NaN
    return ((unsigned long long)k << 32) | pos;
}
__device__ __forceinline__ unsigned long long max_key(float v, unsigned int pos) {       // stored complemented
    const unsigned int k = (v != v) ? 0xffffffffu : float_order_key(v);        // key ~0 decodes to a NaN
    return ~(((unsigned long long)k << 32) | (0xffffffffu - pos));
}
__device__ __forceinline__ void key_min_value(unsigned long long k, float& v, unsigned int& pos) {
    v = float_from_order_key((unsigned int)(k >> 32));
    pos = (unsigned int)k;
}
__device__ __forceinline__ void key_max_value(unsigned long long k, float& v, unsigned int& pos) {
    k = ~k;
    v = float_from_order_key((unsigned int)(k >> 32));
    pos = 0xffffffffu - (unsigned int)k;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other < k ? other : k;
    }
    return k;
}

// keys[c][0] = min key, keys[c][1] = complemented max key; both must be pre-set to all ones.
__global__ void minmax_kernel(const float* __restrict__ x, unsigned long long* __restrict__ keys, long long elems) {
    const int c = blockIdx.y;
    const float* xc = x + (size_t)c * elems;
    unsigned long long kmin = ~0ull, kmax = ~0ull;
    const long long n4 = elems / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xc) + i);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int pos = (unsigned int)(4 * i + e);
            const unsigned long long a = min_key(vv[e], pos), b = max_key(vv[e], pos);
            kmin = a < kmin ? a : kmin;
            kmax = b < kmax ? b : kmax;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < elems; i += blockDim.x) {
            const unsigned long long a = min_key(xc[i], (unsigned int)i), b = max_key(xc[i], (unsigned int)i);
            kmin = a < kmin ? a : kmin;
            kmax = b < kmax ? b : kmax;
        }
    kmin = warp_min_u64(kmin);
    kmax = warp_min_u64(kmax);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 2 * c, kmin);
        atomicMin(keys + 2 * c + 1, kmax);
    }
}

// out = (x - mn) / ((mx - mn) + 1e-15) * scale, in the reference's operation order (bit-compatible with torch).
// The keys may arrive from the LCT's last kernel with their position fields at "unknown" (it reduces values only);
// this pass sees every value, so it writes the positions of the extremes back into the keys (lowest position wins,
// as in the stand-alone reduction) for the backward pass to use.  The value half of a key never changes here.
__device__ __forceinline__ void resolve_positions(float v, unsigned int pos, float mn, float mx, unsigned long long* keys2) {
    if (v == mn) atomicMin(keys2, min_key(v, pos));
    if (v == mx) atomicMin(keys2 + 1, max_key(v, pos));
}

__global__ void normalize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                 unsigned long long* __restrict__ keys, long long elems, float scale) {
    const int c = blockIdx.y;
    float mn, mx;
    unsigned int p0, p1;
    key_min_value(keys[2 * c], mn, p0);
    key_max_value(keys[2 * c + 1], mx, p1);
    // always on: deciding from the keys' current state would let a block that starts late skip its part of the volume
    // and lose the lowest-position tie rule; the two compares per value are free in a pass bound by its memory traffic
    constexpr bool resolve = true;
    const float den = __fadd_rn(__fsub_rn(mx, mn), 1e-15f);
    const float* xc = x + (size_t)c * elems;
    float* oc = out + (size_t)c * elems;
    const long long n4 = elems / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(xc) + i);
        if (resolve) {
            resolve_positions(v.x, (unsigned int)(4 * i), mn, mx, keys + 2 * c);
            resolve_positions(v.y, (unsigned int)(4 * i + 1), mn, mx, keys + 2 * c);
            resolve_positions(v.z, (unsigned int)(4 * i + 2), mn, mx, keys + 2 * c);
            resolve_positions(v.w, (unsigned int)(4 * i + 3), mn, mx, keys + 2 * c);
        }
        v.x = __fmul_rn(__fdiv_rn(__fsub_rn(v.x, mn), den), scale);
        v.y = __fmul_rn(__fdiv_rn(__fsub_rn(v.y, mn), den), scale);
        v.z = __fmul_rn(__fdiv_rn(__fsub_rn(v.z, mn), den), scale);
        v.w = __fmul_rn(__fdiv_rn(__fsub_rn(v.w, mn), den), scale);
        reinterpret_cast<float4*>(oc)[i] = v;
    }
    if (blockIdx.x == 0)
        for (long long i = n4 * 4 + threadIdx.x; i < elems; i += blockDim.x) {
            if (resolve) resolve_positions(xc[i], (unsigned int)i, mn, mx, keys + 2 * c);
            oc[i] = __fmul_rn(__fdiv_rn(__fsub_rn(xc[i], mn), den), scale);
        }
}

// Backward.  With m = min, R = max - min, s = scale / (R + eps):  out_i = s (x_i - m), so
//   gx_j = s g_j - [j = argmin] s G + ([j = argmin] - [j = argmax]) * scale/(R + eps)^2 * H,
//   G = sum_i g_i,  H = sum_i g_i (x_i - m).
// sums[c][0] = G, sums[c][1] = H are accumulated in double by the first kernel (pre-zeroed).
__global__ void normalize_bwd_sums_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                          const unsigned long long* __restrict__ keys, double* __restrict__ sums,
                                          long long elems) {
    const int c = blockIdx.y;
    float mn;
    unsigned int p0;
    key_min_value(keys[2 * c], mn, p0);
    const float* xc = x + (size_t)c * elems;
    const float* gc = g + (size_t)c * elems;
    double G = 0.0, H = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
        const float gi = __ldg(gc + i);
        G += gi;
        H += (double)gi * (double)(__ldg(xc + i) - mn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        G += __shfl_xor_sync(0xffffffffu, G, o);
        H += __shfl_xor_sync(0xffffffffu, H, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sums + 2 * c, G);
        atomicAdd(sums + 2 * c + 1, H);
    }
}

__global__ void normalize_bwd_kernel(const float* __restrict__ g, float* __restrict__ gx,
                                     const unsigned long long* __restrict__ keys, const double* __restrict__ sums,
                                     long long elems, float scale) {
    const int c = blockIdx.y;
    float mn, mx;
    unsigned int pmin, pmax;
    key_min_value(keys[2 * c], mn, pmin);
    key_max_value(keys[2 * c + 1], mx, pmax);
    const double den = (double)__fadd_rn(__fsub_rn(mx, mn), 1e-15f);
    const double s = (double)scale / den;
    const double G = sums[2 * c], H = sums[2 * c + 1];
    const double corr = (double)scale / (den * den) * H;
    const float* gc = g + (size_t)c * elems;
    float* oc = gx + (size_t)c * elems;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < elems; i += (long long)gridDim.x * blockDim.x) {
        double v = s * (double)__ldg(gc + i);
        if (i == (long long)pmin) v += -s * G + corr;
        if (i == (long long)pmax) v -= corr;
        oc[i] = (float)v;
    }
}

}  // namespace lct
