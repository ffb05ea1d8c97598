// The skip branch of the reference's FeatureExtraction, the op that feeds the LCT
// (/root/reference/models/feature_extraction.py:141-145,166-171; NlosPose.py:51-53):
//
//     x_conv2 = F.conv3d(x, weights, stride 1, padding 1)      weights: learnable (1,1,3,3,3)
//     output  = x_conv1 + x_conv2                              x_conv2 broadcast over x_conv1's channels
//
// and its backward (gradient w.r.t. x through this branch, and w.r.t. the 27 weights).
// Streaming 27-tap stencils: a thread owns four neighbouring x positions of one (y) row and walks
// a chunk of time bins, keeping the 3 x 3 x 6 neighbourhood in registers (planes rotate through
// four register sets, the fourth being the prefetch of the next step), so every input value is
// loaded once per thread row instead of 27 times.
#pragma once

#include <cuda_runtime.h>

namespace lct {

constexpr int kSkipThreads = 256;

struct SkipParams {
    const float* src;     // kSkipForward / kSkipWeightGrad: x (B,1,T,N,N);  kSkipDataGrad: g (B,D,T,N,N)
    const float* add;     // kSkipForward: x_conv1 (B,D,T,N,N);  kSkipWeightGrad: g (B,D,T,N,N)
    float* out;           // kSkipForward: (B,D,T,N,N);  kSkipDataGrad: gx (B,1,T,N,N);  kSkipWeightGrad: partial sums [blocks][27]
    const float* w;       // the 27 weights (device memory: they are a learnable parameter)
    int D, T, N;
    int chunk;            // time bins per thread, a multiple of 4 (the plane window rotates through 4 register sets)
};

enum { kSkipForward = 0, kSkipDataGrad = 1, kSkipWeightGrad = 2 };

#ifndef LCT_SKIP_PREFETCH
#define LCT_SKIP_PREFETCH 6           // time planes of look-ahead for the cache prefetch
#endif

// the register window only looks one step ahead -- too short for DRAM latency at two blocks per SM --
// so every 128-byte line a thread row will need is requested a few planes early
__device__ __forceinline__ void skip_prefetch(const float* ptr) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}

// one time plane of the neighbourhood: rows y-1..y+1, columns x0-1..x0+4, zero outside the volume;
// `nsum` channels `stride` apart are summed (the broadcast of x_conv2 turns into a sum in backward)
__device__ __forceinline__ void skip_load_plane(const float* base, size_t stride, int nsum, int t, int T, int N,
                                                int y, int x0, bool active, float (&pl)[3][6]) {
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f, r = 0.f;
        if (active && t >= 0 && t < T && yy >= 0 && yy < N) {
            const float* row = base + ((size_t)t * N + yy) * N + x0;
            for (int d = 0; d < nsum; ++d, row += stride) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(row));
                c.x += q.x; c.y += q.y; c.z += q.z; c.w += q.w;
                if (x0 > 0) l += __ldg(row - 1);
                if (x0 + 4 < N) r += __ldg(row + 4);
            }
        }
        pl[dy][0] = l; pl[dy][1] = c.x; pl[dy][2] = c.y; pl[dy][3] = c.z; pl[dy][4] = c.w; pl[dy][5] = r;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kSkipThreads, 2) skip_kernel(SkipParams p) {
    const int XQ = p.N >> 2;
    const int idx = blockIdx.x * kSkipThreads + threadIdx.x;
    const bool active = idx < p.N * XQ;
    const int y = active ? idx / XQ : 0, x0 = active ? (idx % XQ) * 4 : 0;
    const int b = blockIdx.z, t0 = blockIdx.y * p.chunk;
    const int t1 = min(t0 + p.chunk, p.T);
    const size_t vol = (size_t)p.T * p.N * p.N;

    // window source: x for the forward pass and the weight gradient, sum_d g[b,d] for the data gradient
    const float* win_base = (MODE == kSkipDataGrad) ? p.src + (size_t)b * p.D * vol : p.src + (size_t)b * vol;
    const int win_sum = (MODE == kSkipDataGrad) ? p.D : 1;

    float w[27];          // taps (forward / data gradient) or the 27 running sums (weight gradient)
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        if (MODE == kSkipForward) w[k] = __ldg(p.w + k);
        else if (MODE == kSkipDataGrad) w[k] = __ldg(p.w + 26 - k);      // transpose = all three axes reversed
        else w[k] = 0.f;
    }

    // four planes in registers: three are the current neighbourhood, the fourth is already in flight
    // for the next step (its loads are issued before this step's arithmetic)
    float win[4][3][6];
    skip_load_plane(win_base, vol, win_sum, t0 - 1, p.T, p.N, y, x0, active, win[0]);
    skip_load_plane(win_base, vol, win_sum, t0, p.T, p.N, y, x0, active, win[1]);
    skip_load_plane(win_base, vol, win_sum, t0 + 1, p.T, p.N, y, x0, active, win[2]);

    auto step = [&](int t, float (&lo)[3][6], float (&mid)[3][6], float (&hi)[3][6], float (&next)[3][6]) {
        const bool live = active && t < t1;
        const size_t at = ((size_t)t * p.N + y) * p.N + x0;
        // centre values first, so that they are in flight together with the next plane:
        // x_conv1 (forward, first channel) or sum_d g (weight gradient)
        float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE != kSkipDataGrad && live) {
            const float* cp = p.add + (size_t)b * p.D * vol + at;
            const int n = (MODE == kSkipForward) ? 1 : p.D;
            for (int d = 0; d < n; ++d, cp += vol) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(cp));
                ctr.x += q.x; ctr.y += q.y; ctr.z += q.z; ctr.w += q.w;
            }
        }
        if (LCT_SKIP_PREFETCH > 0 && active && (x0 & 31) == 0 && t + 2 + LCT_SKIP_PREFETCH <= t1) {
            const size_t ahead = at + (size_t)LCT_SKIP_PREFETCH * p.N * p.N;
            for (int d = 0; d < win_sum; ++d) skip_prefetch(win_base + d * vol + ahead + 2 * (size_t)p.N * p.N);
            if (MODE != kSkipDataGrad) {
                const int n = (MODE == kSkipForward) ? 1 : p.D;
                for (int d = 0; d < n; ++d) skip_prefetch(p.add + ((size_t)b * p.D + d) * vol + ahead);
            }
        }
        skip_load_plane(win_base, vol, win_sum, t + 2, p.T, p.N, y, x0, active && t + 2 <= t1, next);
        if (MODE == kSkipWeightGrad) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    w[0 * 9 + dy * 3 + dx] += ctr.x * lo[dy][dx] + ctr.y * lo[dy][dx + 1] + ctr.z * lo[dy][dx + 2] + ctr.w * lo[dy][dx + 3];
                    w[1 * 9 + dy * 3 + dx] += ctr.x * mid[dy][dx] + ctr.y * mid[dy][dx + 1] + ctr.z * mid[dy][dx + 2] + ctr.w * mid[dy][dx + 3];
                    w[2 * 9 + dy * 3 + dx] += ctr.x * hi[dy][dx] + ctr.y * hi[dy][dx + 1] + ctr.z * hi[dy][dx + 2] + ctr.w * hi[dy][dx + 3];
                }
        } else {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s[j] = fmaf(w[0 * 9 + dy * 3 + dx], lo[dy][dx + j], s[j]);
                        s[j] = fmaf(w[1 * 9 + dy * 3 + dx], mid[dy][dx + j], s[j]);
                        s[j] = fmaf(w[2 * 9 + dy * 3 + dx], hi[dy][dx + j], s[j]);
                    }
            if (!live) return;
            if (MODE == kSkipForward) {
                const float* ap = p.add + (size_t)b * p.D * vol + at;
                float* op = p.out + (size_t)b * p.D * vol + at;
                *reinterpret_cast<float4*>(op) = make_float4(ctr.x + s[0], ctr.y + s[1], ctr.z + s[2], ctr.w + s[3]);
                for (int d = 1; d < p.D; ++d) {
                    ap += vol; op += vol;
                    const float4 a = __ldg(reinterpret_cast<const float4*>(ap));
                    *reinterpret_cast<float4*>(op) = make_float4(a.x + s[0], a.y + s[1], a.z + s[2], a.w + s[3]);
                }
            } else {
                *reinterpret_cast<float4*>(p.out + (size_t)b * vol + at) = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
    };

    for (int t = t0; t < t1; t += 4) {
        step(t, win[0], win[1], win[2], win[3]);
        step(t + 1, win[1], win[2], win[3], win[0]);
        step(t + 2, win[2], win[3], win[0], win[1]);
        step(t + 3, win[3], win[0], win[1], win[2]);
    }

    if (MODE == kSkipWeightGrad) {
        // block sum of the 27 running sums in a fixed order -> one row of partial sums per block
        __shared__ float red[kSkipThreads / 32][27];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            float v = w[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < 27) {
            float v = 0.f;
            for (int i = 0; i < kSkipThreads / 32; ++i) v += red[i][threadIdx.x];
            const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            p.out[blk * 27 + threadIdx.x] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Same three passes with the window source staged through shared memory: a block owns RB rows x N
// columns and walks its chunk of time bins behind a ring of kSkipRing planes that cp.async keeps
// kSkipRing - 2 steps ahead of the arithmetic (16-byte copies, zero-filled outside the volume, so
// the borders need no branches in the inner loop).  Each plane is copied once per block and read
// once per thread row (3 x (LDS.128 + 2 LDS.32)) into the rotating register window.
// Used whenever the window source is a single channel (always, except the data gradient at D > 1).
// ---------------------------------------------------------------------------------------------
constexpr int kSkipRing = 8;               // a power of two (ring slots are masked, not divided)

__device__ __forceinline__ void skip_cp_async16(float* dst_smem, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;                                     // 0 source bytes: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void skip_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void skip_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

inline int skip_ring_rows(int N) { const int xq = N / 4; int rb = kSkipThreads / xq; return rb > N ? N : rb; }
inline size_t skip_ring_smem(int N) {          // window planes (halo rows, padded columns) + centre-operand planes
    return (size_t)kSkipRing * ((skip_ring_rows(N) + 2) * (N + 8) + skip_ring_rows(N) * N) * sizeof(float);
}
inline bool skip_ring_ok(int N) { return N / 4 <= kSkipThreads / 2 && skip_ring_smem(N) <= 100 * 1024; }

template <int MODE>
__global__ void __launch_bounds__(kSkipThreads, 2) skip_ring_kernel(SkipParams p) {
    extern __shared__ __align__(16) float ring[];
    const int tid = threadIdx.x;
    const int XQ = p.N >> 2;
    const int RB = min(kSkipThreads / XQ, p.N), RS = p.N + 8, plane = (RB + 2) * RS;
    const int yl = tid / XQ, x0 = (tid % XQ) * 4;
    const int y0 = blockIdx.x * RB, y = y0 + yl;
    const bool active = yl < RB && y < p.N;
    const int b = blockIdx.z, t0 = blockIdx.y * p.chunk;
    const int t1 = min(t0 + p.chunk, p.T);
    const size_t vol = (size_t)p.T * p.N * p.N;
    const float* src = p.src + (size_t)b * ((MODE == kSkipDataGrad) ? p.D : 1) * vol;
    // the centre operand (x_conv1 in the forward pass, g in the weight gradient) rides the same ring,
    // one 16-byte copy per thread and plane, read back by the thread that copied it
    const bool stage_ctr = MODE == kSkipForward || (MODE == kSkipWeightGrad && p.D == 1);
    const float* ctr_src = (MODE == kSkipDataGrad) ? nullptr : p.add + (size_t)b * p.D * vol;
    float* cring = ring + kSkipRing * plane;
    float* out_b = p.out + (size_t)b * ((MODE == kSkipDataGrad) ? 1 : p.D) * vol;      // unused by the weight gradient
    const int cplane = RB * p.N;

    float w[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        if (MODE == kSkipForward) w[k] = __ldg(p.w + k);
        else if (MODE == kSkipDataGrad) w[k] = __ldg(p.w + 26 - k);
        else w[k] = 0.f;
    }

    // pad columns (4 floats either side of every row) stay zero for the whole kernel
    for (int i = tid; i < kSkipRing * plane / 4; i += kSkipThreads) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // Everything that does not change from plane to plane is worked out once: this thread's (at most two) copy
    // slots of a window plane as {shared offset, offset inside a global plane, row inside the volume}, its centre
    // slot, and its three window rows.  Per plane only the ring slot and the plane base are added.
    const int NN = p.N * p.N;
    int cs_off[2], cg_off[2];
    bool cs_has[2], cs_row_ok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = tid + k * kSkipThreads;
        const int r = i / XQ, q = i - r * XQ, yy = y0 - 1 + r;
        cs_has[k] = i < (RB + 2) * XQ;
        cs_row_ok[k] = cs_has[k] && yy >= 0 && yy < p.N;
        cs_off[k] = r * RS + 4 + 4 * q;
        cg_off[k] = yy * p.N + 4 * q;
    }
    const int my_ctr = yl * p.N + x0;              // centre slot inside a centre plane
    const int my_row = yl * RS + 4 + x0;           // first of this thread's three window rows inside a window plane
    const int my_at = y * p.N + x0;                // this thread's four outputs inside a global plane

    auto ring_slot = [&](int u) { return (u + 1) & (kSkipRing - 1); };                  // u >= -1; kSkipRing is a power of two
    auto issue = [&](int u) {                                                           // plane u -> its slot, one group
        const int slot = ring_slot(u);
        float* dst = ring + slot * plane;
        const bool t_ok = u >= 0 && u < p.T && u <= t1;
        const float* gp = src + (size_t)(t_ok ? u : 0) * NN;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (cs_has[k]) {
                const bool ok = t_ok && cs_row_ok[k];
                skip_cp_async16(dst + cs_off[k], ok ? gp + cg_off[k] : src, ok);
            }
        if (MODE != kSkipDataGrad && stage_ctr && active) {
            const bool ok = u >= t0 && u < t1;
            skip_cp_async16(cring + slot * cplane + my_ctr, ok ? ctr_src + (size_t)u * NN + my_at : ctr_src, ok);
        }
        skip_cp_commit();
    };
    auto fetch = [&](int u, float (&pl)[3][6]) {                                        // this thread's rows of plane u
        const float* base = ring + ring_slot(u) * plane + my_row;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float* r = base + dy * RS;
            const float4 c = *reinterpret_cast<const float4*>(r);
            pl[dy][0] = r[-1]; pl[dy][1] = c.x; pl[dy][2] = c.y; pl[dy][3] = c.z; pl[dy][4] = c.w; pl[dy][5] = r[4];
        }
    };

    for (int u = t0 - 1; u < t0 - 1 + kSkipRing; ++u) issue(u);
    skip_cp_wait<kSkipRing - 3>();                 // the first three planes have landed
    __syncthreads();
    float win[3][3][6];
    if (active) { fetch(t0 - 1, win[0]); fetch(t0, win[1]); }
    __syncthreads();                               // plane t0 - 1's slot may be overwritten from here on

    auto step = [&](int t, float (&lo)[3][6], float (&mid)[3][6], float (&hi)[3][6]) {
        const bool live = active && t < t1;
        const size_t at = (size_t)t * NN + my_at;
        float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE != kSkipDataGrad && live) {
            if (stage_ctr) {
                ctr = *reinterpret_cast<const float4*>(cring + ring_slot(t) * cplane + my_ctr);
            } else {
                const float* cp = ctr_src + at;
                for (int d = 0; d < p.D; ++d, cp += vol) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(cp));
                    ctr.x += q.x; ctr.y += q.y; ctr.z += q.z; ctr.w += q.w;
                }
            }
        }
        issue(t + kSkipRing - 1);                  // into the slot of plane t - 1 (read by everyone two steps ago)
        if (active) fetch(t + 1, hi);
        if (MODE == kSkipWeightGrad) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    w[0 * 9 + dy * 3 + dx] += ctr.x * lo[dy][dx] + ctr.y * lo[dy][dx + 1] + ctr.z * lo[dy][dx + 2] + ctr.w * lo[dy][dx + 3];
                    w[1 * 9 + dy * 3 + dx] += ctr.x * mid[dy][dx] + ctr.y * mid[dy][dx + 1] + ctr.z * mid[dy][dx + 2] + ctr.w * mid[dy][dx + 3];
                    w[2 * 9 + dy * 3 + dx] += ctr.x * hi[dy][dx] + ctr.y * hi[dy][dx + 1] + ctr.z * hi[dy][dx + 2] + ctr.w * hi[dy][dx + 3];
                }
        } else if (live) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s[j] = fmaf(w[0 * 9 + dy * 3 + dx], lo[dy][dx + j], s[j]);
                        s[j] = fmaf(w[1 * 9 + dy * 3 + dx], mid[dy][dx + j], s[j]);
                        s[j] = fmaf(w[2 * 9 + dy * 3 + dx], hi[dy][dx + j], s[j]);
                    }
            if (MODE == kSkipForward) {
                const float* ap = ctr_src + at;
                float* op = out_b + at;
                *reinterpret_cast<float4*>(op) = make_float4(ctr.x + s[0], ctr.y + s[1], ctr.z + s[2], ctr.w + s[3]);
                for (int d = 1; d < p.D; ++d) {
                    ap += vol; op += vol;
                    const float4 a = __ldg(reinterpret_cast<const float4*>(ap));
                    *reinterpret_cast<float4*>(op) = make_float4(a.x + s[0], a.y + s[1], a.z + s[2], a.w + s[3]);
                }
            } else {
                *reinterpret_cast<float4*>(out_b + at) = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
        skip_cp_wait<kSkipRing - 3>();             // plane t + 2 has landed (this thread's copies) ...
        __syncthreads();                           // ... and everyone's; everyone is done reading plane t + 1
    };

    for (int t = t0; t < t1; t += 3) {
        step(t, win[0], win[1], win[2]);
        step(t + 1, win[1], win[2], win[0]);
        step(t + 2, win[2], win[0], win[1]);
    }
    skip_cp_wait<0>();

    if (MODE == kSkipWeightGrad) {
        __shared__ float red[kSkipThreads / 32][27];
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            float v = active ? w[k] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (tid < 27) {
            float v = 0.f;
            for (int i = 0; i < kSkipThreads / 32; ++i) v += red[i][tid];
            const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            p.out[blk * 27 + tid] = v;
        }
    }
}

inline dim3 skip_ring_grid(int B, int T, int N, int chunk) {
    const int rb = skip_ring_rows(N);
    return dim3((unsigned)((N + rb - 1) / rb), (unsigned)((T + chunk - 1) / chunk), (unsigned)B);
}

// gw[k] = sum over blocks of partial[blk][k], in double and in a fixed order (warp k, lanes stride the blocks)
__global__ void skip_weight_reduce_kernel(const float* __restrict__ partial, int blocks, float* __restrict__ gw) {
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0.0;
    for (int i = lane; i < blocks; i += 32) s += (double)partial[(size_t)i * 27 + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) gw[k] = (float)s;
}

inline dim3 skip_grid(int B, int T, int N, int chunk) {
    const int rows = N * (N / 4);
    return dim3((unsigned)((rows + kSkipThreads - 1) / kSkipThreads), (unsigned)((T + chunk - 1) / chunk), (unsigned)B);
}

// time bins per thread: the multiple of 4 that needs the fewest (waves of resident blocks) x (steps per block)
inline int skip_chunk(int B, int T, int N, int num_sms, bool ring) {
    const long slots = 2L * num_sms;                          // two 256-thread blocks per SM (register limit)
    int best = 4;
    long best_cost = -1;
    for (int chunk = 4; chunk <= 128; chunk += 4) {
        const dim3 g = ring ? skip_ring_grid(B, T, N, chunk) : skip_grid(B, T, N, chunk);
        const long blocks = (long)g.x * g.y * g.z;
        const long cost = ((blocks + slots - 1) / slots) * (chunk + (ring ? kSkipRing : 3));
        if (best_cost < 0 || cost < best_cost) { best = chunk; best_cost = cost; }
        if (chunk >= T) break;
    }
    return best;
}

}  // namespace lct
