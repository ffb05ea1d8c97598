// C ABI of the B200 LCT library (see include/hiddenpose_lct.h).
// Host side only: plan construction, workspace carving, kernel launches.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "hiddenpose_lct.h"
#include "lct_chain.cuh"
#include "lct_filter_build.cuh"
#include "lct_normalize.cuh"
#include "lct_skipconv.cuh"
#include "lct_stencil.cuh"
#include "lct_tables.h"

#ifndef LCT_DEFAULT_PDL
#define LCT_DEFAULT_PDL 1
#endif
#ifndef LCT_DEFAULT_STREAM_GROUPS
#define LCT_DEFAULT_STREAM_GROUPS 2
#endif

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_last_error = what;
    if (e != cudaSuccess) {
        g_last_error += ": ";
        g_last_error += cudaGetErrorString(e);
    }
    return code;
}

#define LCT_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return fail(LCT_ERR_CUDA, #call, e_); \
    } while (0)

constexpr int kMaxDevices = 64;
constexpr size_t kHeaderAlign = 256;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

std::mutex g_tw_mutex;
bool g_tw_ready[kMaxDevices] = {};

int upload_twiddles(int device) {
    std::lock_guard<std::mutex> lock(g_tw_mutex);
    if (g_tw_ready[device]) return LCT_OK;
    std::vector<float2> tw(lct::kTwN);
    for (int j = 0; j < lct::kTwN; ++j) {
        const double a = -2.0 * M_PI * (double)j / (double)lct::kTwN;
        tw[j] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    LCT_CUDA(cudaMemcpyToSymbol(lct::c_tw, tw.data(), sizeof(float2) * lct::kTwN));
    LCT_CUDA(cudaMemcpyToSymbol(lct::g_tw, tw.data(), sizeof(float2) * lct::kTwN));
    g_tw_ready[device] = true;
    return LCT_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct GpuLauncher {
    cudaStream_t stream;
    int device;
    cudaError_t err = cudaSuccess;
    void* const* events = nullptr;          // optional: 6 cudaEvent_t recorded around the stages
    int prefetch_ahead = 0;                 // SM count when the time kernels should warm L2 for later blocks, else 0
    int pdl = 0;                            // programmatic dependent launch mode (Params::pdl); off while stage events are recorded
    void mark(int i) {
        if (events && err == cudaSuccess) err = cudaEventRecord((cudaEvent_t)events[i], stream);
    }
    template <class K> int launch(const lct::Params& p) {
        static std::atomic<bool> attr_set[kMaxDevices] = {};
        auto kern = lct::lct_kernel<K>;
        if (!attr_set[device]) {
            if (K::kSmem > 48 * 1024) {
                err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::kSmem);
                if (err != cudaSuccess) return 1;
            }
            // ask for the full shared-memory carve-out so occupancy is not capped at one block's worth
            err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            if (err != cudaSuccess) return 1;
            attr_set[device] = true;
        }
        lct::Params q = p;
        q.ahead = prefetch_ahead > 0 ? prefetch_ahead * K::kMinBlocks : 0;      // resident blocks of this kernel
        int gx, gy;
        K::grid(q, gx, gy);
        q.pdl = events ? 0 : pdl;
        const int iters = K::iterations(q);
        if (q.pdl) {
            // the kernel may be scheduled before its predecessor in the stream has drained; it orders itself behind
            // that kernel's memory with griddepcontrol.wait (lct_kernel)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(gx, gy);
            cfg.blockDim = dim3(K::kThreads);
            cfg.dynamicSmemBytes = K::kSmem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            err = cudaLaunchKernelEx(&cfg, kern, q, iters);
        } else {
            kern<<<dim3(gx, gy), K::kThreads, K::kSmem, stream>>>(q, iters);
            err = cudaGetLastError();
        }
        return err == cudaSuccess ? 0 : 1;
    }
};

template <class T> int to_device(const T* host, size_t count, T** out) {
    LCT_CUDA(cudaMalloc((void**)out, count * sizeof(T)));
    LCT_CUDA(cudaMemcpy(*out, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return LCT_OK;
}

// Per-sample window begins travel to the device as kernel parameters (256 per launch), not as a copy from the
// caller's pageable array: a parameter block is captured by value, so ragged windows are legal under stream
// capture (a CUDA graph replays the windows it was captured with) and nothing reads host memory after the call.
constexpr int kWindowChunk = 256;
struct WindowChunk { int v[kWindowChunk]; };
__global__ void set_windows_kernel(int* dst, const WindowChunk w, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = w.v[threadIdx.x];
}

__global__ void scale_filter_kernel(float2* f, size_t n, float s) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 v = f[i];
        f[i] = make_float2(v.x * s, v.y * s);
    }
}

}  // namespace

// Builds the scaled filter on the device from the PSF support (lct_filter_build.cuh).
template <int N>
int build_filter_on_device(const lct_desc* d, float2* out, int layout, float scale) {
    using PH = typename lct::ColPlan<2 * N>::type;
    using PL = typename lct::LinePlan<2 * N>::type;
    constexpr int L = 2 * N, CT = (L >= 512) ? 16 : (L < 32 ? L : 32), RB = lct::LineRows<N>::RB;
    const int M = d->time_bins;
    const size_t nfilt = (size_t)(M + 1) * L * L;
    float2* planes = nullptr;
    int *dz = nullptr, *dyx = nullptr;
    auto cleanup = [&]() { cudaFree(planes); cudaFree(dz); cudaFree(dyx); };
#define LCT_TRYC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(LCT_ERR_CUDA, #call, e_); } } while (0)
    LCT_TRYC(cudaMalloc((void**)&planes, nfilt * sizeof(float2)));
    LCT_TRYC(cudaMemset(planes, 0, nfilt * sizeof(float2)));
    LCT_TRYC(cudaMalloc((void**)&dz, sizeof(int) * d->psf_count));
    LCT_TRYC(cudaMalloc((void**)&dyx, sizeof(int) * d->psf_count));
    LCT_TRYC(cudaMemcpy(dz, d->psf_z, sizeof(int) * d->psf_count, cudaMemcpyHostToDevice));
    LCT_TRYC(cudaMemcpy(dyx, d->psf_yx, sizeof(int) * d->psf_count, cudaMemcpyHostToDevice));
    lct::psf_planes_kernel<<<dim3((d->psf_count + 255) / 256, M + 1), 256>>>(planes, dz, dyx, d->psf_count, d->psf_value, M, L);
    {
        auto kern = lct::column_fft_kernel<PH, CT>;
        constexpr size_t smem = (size_t)L * CT * sizeof(float2);
        if (smem > 48 * 1024) LCT_TRYC(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(L / CT, M + 1), PH::TL * CT, smem>>>(planes);
    }
    {
        auto kern = lct::row_fft_wiener_kernel<PL, PH, RB>;
        constexpr int RS = L + PL::R0 + ((PL::TL < 16) ? 8 : 0);
        constexpr size_t smem = (size_t)RB * RS * sizeof(float2);
        if (smem > 48 * 1024) LCT_TRYC(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lct::FilterBuildParams fp{M, N, planes, out, layout, 1.0f / d->snr, scale, d->method_bp};
        // the quarter layout only keeps rows kh <= N: the blocks above them have nothing to write
        kern<<<dim3(layout >= 2 ? N / RB + 1 : L / RB, M + 1), PL::TL * RB, smem>>>(fp);
    }
    LCT_TRYC(cudaGetLastError());
    LCT_TRYC(cudaDeviceSynchronize());
#undef LCT_TRYC
    cleanup();
    return LCT_OK;
}

struct DeviceBand {
    float4* ell = nullptr;
    int* rowptr = nullptr;
    float* vals = nullptr;
    float4* pair = nullptr;            // mtx tables only (the time-forward kernel's pair records)
    lct::BandTable view() const { return lct::BandTable{ell, rowptr, vals, pair}; }
    void release() { cudaFree(ell); cudaFree(rowptr); cudaFree(vals); cudaFree(pair); ell = nullptr; rowptr = nullptr; vals = nullptr; pair = nullptr; }
};

struct lct_plan {
    int M = 0, N = 0, device = 0;
    DeviceBand mtx_falloff, mtx, mtxi, mtxi_falloff;
    // Channel groups run as independent kernel chains on side streams, so the tail of one group's
    // kernel overlaps the next kernel of another group (the chains differ in what bounds them).
    static constexpr int kMaxGroups = 8;
    int groups = 1;
    int prefetch_ahead = 0;                 // SM count, or 0 when LCT_L2_PREFETCH=0 (see GpuLauncher::prefetch_ahead)
    int pdl = 0;                            // LCT_PDL: programmatic dependent launch between the kernels of a chain
    // One set of side streams + fork/join events per caller stream (created up front, handed out first come first
    // served): two threads driving the plan on two streams get two sets, so neither waits on the other's kernels.
    // More distinct caller streams than sets share sets round robin -- still correct (every use is ordered by its
    // own fork/join events), only then with a false dependency between the sharers.
    static constexpr int kSideSets = 4;
    struct SideSet {
        cudaStream_t side[kMaxGroups] = {};
        cudaEvent_t fork = nullptr, join[kMaxGroups] = {};
        cudaStream_t owner = nullptr;
        bool owned = false;
    };
    mutable SideSet sets[kSideSets];
    mutable int next_set = 0;
    mutable std::mutex side_mutex;          // guards the set table and the enqueue order on a set
    SideSet& set_for(cudaStream_t caller) const {       // call with side_mutex held
        for (int i = 0; i < kSideSets; ++i)
            if (sets[i].owned && sets[i].owner == caller) return sets[i];
        for (int i = 0; i < kSideSets; ++i)
            if (!sets[i].owned) { sets[i].owned = true; sets[i].owner = caller; return sets[i]; }
        SideSet& s = sets[next_set];
        next_set = (next_set + 1) % kSideSets;
        return s;
    }
    float2* filt = nullptr;         // natural layout (unfused K3) or, with filt_sym, its quarter (M+1, N+1, N+1); or
    float2* filt_plane = nullptr;   // [kt][kw][plane row] (plane-fused kernel); exactly one of the two is set
    int filt_sym = 0;               // lct::FilterLayout of `filt`
    bool fused() const { return filt_plane != nullptr; }
    // S1 (C, M+1, N, N) c64, plus S2 (C, M+1, 2N, N) c64 when the middle stages are not fused
    size_t per_channel_bytes() const { return (size_t)(M + 1) * N * N * sizeof(float2) * (fused() ? 1 : 3); }
    lct::ChainTables tables() const {
        return lct::ChainTables{mtx_falloff.view(), mtx.view(), mtxi.view(), mtxi_falloff.view(), filt, filt_plane, filt_sym};
    }
};

namespace {
// The PSF support is mirror-symmetric in y and x about the half-sample centre (after the roll of helper.py:115-116:
// voxel (z, y, x) <-> (z, -1-y mod 2N, x) and (z, y, -1-x mod 2N)) -- checked, not assumed: the quarter filter
// layout is only valid then.
bool psf_support_is_mirror_symmetric(const lct_desc* d) {
    const int L = 2 * d->spatial;
    std::vector<long long> keys((size_t)d->psf_count);
    for (int i = 0; i < d->psf_count; ++i) keys[i] = (long long)d->psf_z[i] * L * L + d->psf_yx[i];
    std::sort(keys.begin(), keys.end());
    for (int i = 0; i < d->psf_count; ++i) {
        const int y = d->psf_yx[i] / L, x = d->psf_yx[i] % L;
        const long long my = (long long)d->psf_z[i] * L * L + (long long)((L - 1 - y) % L) * L + x;
        const long long mx = (long long)d->psf_z[i] * L * L + (long long)y * L + (L - 1 - x) % L;
        if (!std::binary_search(keys.begin(), keys.end(), my) || !std::binary_search(keys.begin(), keys.end(), mx)) return false;
    }
    return true;
}
// Same property for a caller-supplied half spectrum: W(kh) = w^kh W(2N - kh), W(kw) = w^kw W(2N - kw), to 1e-5 of
// the largest entry (a host-built filter carries the rounding of its own FFT).
bool host_filter_is_mirror_symmetric(const float2* f, int M, int N) {
    const int L = 2 * N;
    double big = 0.0;
    const size_t n = (size_t)(M + 1) * L * L;
    for (size_t i = 0; i < n; ++i) big = std::max(big, std::max(std::fabs((double)f[i].x), std::fabs((double)f[i].y)));
    const double tol = 1e-5 * big;
    std::vector<double> c(L), sn(L);
    for (int k = 0; k < L; ++k) { c[k] = std::cos(-2.0 * M_PI * k / L); sn[k] = std::sin(-2.0 * M_PI * k / L); }
    for (int kt = 0; kt <= M; ++kt)
        for (int kh = 0; kh < L; ++kh)
            for (int kw = 0; kw < L; ++kw) {
                const float2 v = f[((size_t)kt * L + kh) * L + kw];
                if (kh > N) {
                    const float2 u = f[((size_t)kt * L + (L - kh)) * L + kw];
                    if (std::fabs(u.x * c[kh] - u.y * sn[kh] - v.x) > tol || std::fabs(u.x * sn[kh] + u.y * c[kh] - v.y) > tol) return false;
                }
                if (kw > N) {
                    const float2 u = f[((size_t)kt * L + kh) * L + (L - kw)];
                    if (std::fabs(u.x * c[kw] - u.y * sn[kw] - v.x) > tol || std::fabs(u.x * sn[kw] + u.y * c[kw] - v.y) > tol) return false;
                }
            }
    return true;
}

int upload_band(const std::vector<lct::EllRow>& ell, const std::vector<int32_t>& rowptr, const std::vector<float>& vals,
                DeviceBand& out, const std::vector<lct::PairRow>* pairs = nullptr) {
    static_assert(sizeof(lct::EllRow) == sizeof(float4), "EllRow must be 16 bytes");
    static_assert(sizeof(lct::PairRow) == 2 * sizeof(float4), "PairRow must be 32 bytes");
    int rc = to_device(reinterpret_cast<const float4*>(ell.data()), ell.size(), &out.ell);
    if (rc) return rc;
    if (pairs) {
        rc = to_device(reinterpret_cast<const float4*>(pairs->data()), pairs->size() * 2, &out.pair);
        if (rc) return rc;
    }
    rc = to_device(rowptr.data(), rowptr.size(), &out.rowptr);
    if (rc) return rc;
    return to_device(vals.data(), vals.size(), &out.vals);
}
}  // namespace

extern "C" {

int lct_abi_version(void) { return LCT_ABI_VERSION; }

const char* lct_error_string(int code) {
    switch (code) {
        case LCT_OK: return "ok";
        case LCT_ERR_INVALID: return "invalid argument";
        case LCT_ERR_UNSUPPORTED: return "unsupported size (time_bins must be a power of two in [32,512], spatial in [8,256])";
        case LCT_ERR_CUDA: return "CUDA error";
        case LCT_ERR_WORKSPACE: return "workspace too small";
        case LCT_ERR_NOMEM: return "out of memory";
        default: return "unknown error";
    }
}

const char* lct_last_error(void) { return g_last_error.c_str(); }

void lct_plan_destroy(lct_plan* plan) {
    if (!plan) return;
    DeviceGuard g(plan->device);
    plan->mtx_falloff.release(); plan->mtx.release(); plan->mtxi.release(); plan->mtxi_falloff.release();
    cudaFree(plan->filt);
    cudaFree(plan->filt_plane);
    for (auto& set : plan->sets) {
        for (int g = 0; g < lct_plan::kMaxGroups; ++g) {
            if (set.side[g]) cudaStreamDestroy(set.side[g]);
            if (set.join[g]) cudaEventDestroy(set.join[g]);
        }
        if (set.fork) cudaEventDestroy(set.fork);
    }
    delete plan;
}

int lct_plan_create(const lct_desc* d, lct_plan** out) {
    if (!d || !out) return fail(LCT_ERR_INVALID, "null descriptor");
    *out = nullptr;
    if (!d->mtx_rowptr || !d->mtx_colidx || !d->mtx_vals) return fail(LCT_ERR_INVALID, "null operator table");
    const bool device_filter = d->filter_half == nullptr;
    if (device_filter && (!d->psf_z || !d->psf_yx || d->psf_count <= 0 || !(d->snr > 0.f) || !(d->psf_value > 0.f)))
        return fail(LCT_ERR_INVALID, "neither a filter nor a PSF support was given");
    const int M = d->time_bins, N = d->spatial;
    if (!lct::supported_M(M) || !lct::supported_N(N)) return fail(LCT_ERR_UNSUPPORTED, "size not compiled");
    if (d->device < 0 || d->device >= kMaxDevices) return fail(LCT_ERR_INVALID, "bad device ordinal");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LCT_ERR_CUDA, "cudaSetDevice failed");
    int rc = upload_twiddles(d->device);
    if (rc) return rc;

    // validate the operator and derive the banded row tables for mtx and mtxi = mtx^T (helper.py:61)
    lct::HostTables ht;
    const std::string why = lct::build_tables(M, d->mtx_rowptr, d->mtx_colidx, d->mtx_vals, d->falloff, lct::time_tail_rows(M), ht, lct::time_long_pairs(M), lct::time_tile_columns(M));
    if (!why.empty()) return fail(LCT_ERR_INVALID, why.c_str());

    lct_plan* p = new (std::nothrow) lct_plan();
    if (!p) return fail(LCT_ERR_NOMEM, "plan allocation");
    p->M = M; p->N = N; p->device = d->device;
    const size_t nfilt = (size_t)(M + 1) * 4 * N * N;
#define LCT_TRY(expr) do { rc = (expr); if (rc) { lct_plan_destroy(p); return rc; } } while (0)
    LCT_TRY(upload_band(ht.mtx_ell_falloff, ht.mtx_rowptr, ht.mtx_vals_falloff, p->mtx_falloff, &ht.mtx_pair_falloff));
    LCT_TRY(upload_band(ht.mtx_ell, ht.mtx_rowptr, ht.mtx_vals, p->mtx, &ht.mtx_pair));
    LCT_TRY(upload_band(ht.mtxi_ell, ht.mtxi_rowptr, ht.mtxi_vals, p->mtxi));
    LCT_TRY(upload_band(ht.mtxi_ell_falloff, ht.mtxi_rowptr, ht.mtxi_vals_falloff, p->mtxi_falloff));
    // fold torch.ifft's 1/(2M*2N*2N) (tflct.py:151) into the filter; a power of two, so exact
    const float scale = 1.0f / (8.0f * (float)M * (float)N * (float)N);
    const bool fused = lct::plane_fusable(N) && !(d->reserved & LCT_FLAG_NO_PLANE_FUSION);
    float2** slot = fused ? &p->filt_plane : &p->filt;
    const bool allow_sym = !fused && !(d->reserved & LCT_FLAG_FULL_FILTER);
    // which part of a symmetric filter the unfused column kernel of this size wants (lct_kernels.cuh, Params::filt_sym)
    const int sym_layout = N >= 256 ? lct::kFilterHalfRows : lct::kFilterQuarter;
    const int sym_pitch = sym_layout == lct::kFilterHalfRows ? 2 * N : N + 1;
    const size_t nquarter = (size_t)(M + 1) * (N + 1) * sym_pitch;
    if (device_filter) {
        for (int i = 0; i < d->psf_count; ++i)
            if (d->psf_z[i] < 0 || d->psf_z[i] >= 2 * M || d->psf_yx[i] < 0 || d->psf_yx[i] >= 4 * N * N) {
                lct_plan_destroy(p);
                return fail(LCT_ERR_INVALID, "PSF voxel out of range");
            }
        p->filt_sym = (allow_sym && psf_support_is_mirror_symmetric(d)) ? sym_layout : lct::kFilterFull;
        cudaError_t e = cudaMalloc((void**)slot, (p->filt_sym ? nquarter : nfilt) * sizeof(float2));
        if (e != cudaSuccess) { lct_plan_destroy(p); return fail(LCT_ERR_NOMEM, "filter allocation", e); }
        rc = -1;
        LCT_SWITCH_N(N, (build_filter_on_device<kN>(d, *slot, fused ? 1 : (p->filt_sym == lct::kFilterQuarter ? 2 : (p->filt_sym == lct::kFilterHalfRows ? 3 : 0)), scale)));
        if (rc) { lct_plan_destroy(p); return rc < 0 ? fail(LCT_ERR_UNSUPPORTED, "size not compiled") : rc; }
    } else if (allow_sym && host_filter_is_mirror_symmetric(reinterpret_cast<const float2*>(d->filter_half), M, N)) {
        const int L = 2 * N;
        const float2* nat = reinterpret_cast<const float2*>(d->filter_half);
        std::vector<float2> quarter(nquarter);
        for (int kt = 0; kt <= M; ++kt)
            for (int kh = 0; kh <= N; ++kh)
                std::memcpy(&quarter[((size_t)kt * (N + 1) + kh) * sym_pitch], &nat[((size_t)kt * L + kh) * L], sizeof(float2) * sym_pitch);
        LCT_TRY(to_device(quarter.data(), nquarter, slot));
        p->filt_sym = sym_layout;
        scale_filter_kernel<<<1024, 256>>>(*slot, nquarter, scale);
    } else {
        if (fused) {
            // fused layout: [kt][kw/2][plane row][kw&1] -- both output parities of a row in one 128-bit load,
            // rows in the order the forward H stages leave them
            const int L = 2 * N;
            const float2* nat = reinterpret_cast<const float2*>(d->filter_half);
            std::vector<float2> perm(nfilt);
            std::vector<int> kh_of_row(L);
            for (int r = 0; r < L; ++r) {
                int rc = -1;
                LCT_SWITCH_N(N, (lct::plane_row_freq<kN>(r)));
                kh_of_row[r] = rc;
            }
            for (int kt = 0; kt <= M; ++kt)
                for (int kw = 0; kw < L; ++kw)
                    for (int r = 0; r < L; ++r)
                        perm[(((size_t)kt * N + (kw >> 1)) * L + r) * 2 + (kw & 1)] = nat[((size_t)kt * L + kh_of_row[r]) * L + kw];
            LCT_TRY(to_device(perm.data(), nfilt, slot));
        } else {
            LCT_TRY(to_device(reinterpret_cast<const float2*>(d->filter_half), nfilt, slot));
        }
        scale_filter_kernel<<<1024, 256>>>(*slot, nfilt, scale);
    }
#undef LCT_TRY
    {
        const char* pf = std::getenv("LCT_L2_PREFETCH");
        int sms = 0;
        if ((!pf || std::atoi(pf) != 0) && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device) == cudaSuccess)
            p->prefetch_ahead = sms;
    }
    {
        const char* env = std::getenv("LCT_PDL");
        p->pdl = (env ? std::atoi(env) : LCT_DEFAULT_PDL) ? 1 : 0;
    }
    {
        const char* env = std::getenv("LCT_STREAM_GROUPS");
        // Measured with dependent launch on (round 2, one B200): two groups win wherever a kernel of one group has a
        // tail the other group can fill (8 x 256x64^2: 179 vs 191 us, 16 x 128^3: 740 vs 746, 4 x 256x256^2: 1897 vs
        // 1950), except at 512x128^2, where the time kernels hold a whole SM per block and two chains only get in
        // each other's way (2 / 4 / 8 channels: 457 / 815 / 1535 us on one stream against 496 / 845 / 1551 on two)
        const int by_shape = (p->M == 512 && p->N == 128) ? 1 : LCT_DEFAULT_STREAM_GROUPS;
        int g = env ? std::atoi(env) : by_shape;
        p->groups = g < 1 ? 1 : (g > lct_plan::kMaxGroups ? lct_plan::kMaxGroups : g);
        for (auto& set : p->sets) {
            for (int i = 1; i < p->groups; ++i) {
                if (cudaStreamCreateWithFlags(&set.side[i], cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&set.join[i], cudaEventDisableTiming) != cudaSuccess) {
                    lct_plan_destroy(p);
                    return fail(LCT_ERR_CUDA, "side stream creation");
                }
            }
            if (p->groups > 1 && cudaEventCreateWithFlags(&set.fork, cudaEventDisableTiming) != cudaSuccess) {
                lct_plan_destroy(p);
                return fail(LCT_ERR_CUDA, "fork event creation");
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { lct_plan_destroy(p); return fail(LCT_ERR_CUDA, "filter scaling", e); }
    *out = p;
    return LCT_OK;
}

int32_t lct_plan_time_bins(const lct_plan* plan) { return plan ? plan->M : 0; }
int32_t lct_plan_spatial(const lct_plan* plan) { return plan ? plan->N : 0; }

size_t lct_plan_workspace_bytes(const lct_plan* plan, int32_t channels) {
    if (!plan || channels <= 0) return 0;
    return kHeaderAlign * 16 + plan->per_channel_bytes() * (size_t)channels;
}

static int run(const lct_plan* plan, const float* in, const int32_t* tbe, const int32_t* ten,
               int B, int D, int Tin, float* out, void* ws, size_t ws_bytes, void* stream_, bool backward,
               void* const* events = nullptr, unsigned long long* minmax_keys = nullptr) {
    if (!plan || !in || !out || !tbe || !ten || !ws) return fail(LCT_ERR_INVALID, "null argument");
    if (B <= 0 || D <= 0 || Tin <= 0 || Tin > plan->M) return fail(LCT_ERR_INVALID, "bad shape");
    if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)ws) & 15) return fail(LCT_ERR_INVALID, "buffers must be 16-byte aligned");
    bool uniform = true;
    for (int b = 0; b < B; ++b) {
        if (tbe[b] < 0 || ten[b] > plan->M || ten[b] - tbe[b] != Tin) return fail(LCT_ERR_INVALID, "bad time window");
        uniform = uniform && tbe[b] == tbe[0];
    }
    const int M = plan->M, N = plan->N;
    const long long C = (long long)B * D;
    const size_t header = kHeaderAlign * 16;
    if (ws_bytes < header + plan->per_channel_bytes()) return fail(LCT_ERR_WORKSPACE, "workspace too small");
    if (!uniform && (size_t)B * sizeof(int) > header) return fail(LCT_ERR_INVALID, "too many distinct windows (batch > 1024)");
    long long chunk = (long long)((ws_bytes - header) / plan->per_channel_bytes());
    const long long grid_cap = 65535 / (M + 1);
    if (chunk > grid_cap) chunk = grid_cap;
    if (chunk > C) chunk = C;

    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(LCT_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned char* base = static_cast<unsigned char*>(ws);
    int* be_dev = nullptr;
    if (!uniform) {
        be_dev = reinterpret_cast<int*>(base);
        for (int b0 = 0; b0 < B; b0 += kWindowChunk) {
            WindowChunk w;
            const int n = B - b0 < kWindowChunk ? B - b0 : kWindowChunk;
            std::memcpy(w.v, tbe + b0, sizeof(int) * n);
            set_windows_kernel<<<1, kWindowChunk, 0, stream>>>(be_dev + b0, w, n);
        }
        LCT_CUDA(cudaGetLastError());
    }
    float2* s1 = reinterpret_cast<float2*>(base + header);
    float2* s2 = s1 + (size_t)chunk * (M + 1) * N * N;
    const size_t in_stride = (size_t)(backward ? M : Tin) * N * N;
    const size_t out_stride = (size_t)(backward ? Tin : M) * N * N;
    if (events && chunk < C) return fail(LCT_ERR_WORKSPACE, "stage events need a workspace for the whole batch");
    if (minmax_keys) LCT_CUDA(cudaMemsetAsync(minmax_keys, 0xFF, (size_t)C * 2 * sizeof(unsigned long long), stream));
    const lct::ChainTables t = plan->tables();
    if (plan->groups > 1 && chunk >= C && C >= 2 && !events) {     // per-kernel events need the single-stream order
        // one batch: split the channels into groups, each an independent chain on its own stream
        std::lock_guard<std::mutex> lock(plan->side_mutex);
        lct_plan::SideSet& set = plan->set_for(stream);
        const int G = (int)(C < plan->groups ? C : plan->groups);
        LCT_CUDA(cudaEventRecord(set.fork, stream));
        // Whatever fails below, every side stream that was given work is joined back into the caller's stream
        // before the error is returned: the caller is free to release the buffers as soon as its own stream has
        // passed this call, so nothing may be left running unordered on a side stream.
        int status = LCT_OK, forked = 0;
        for (int g = 0; g < G && status == LCT_OK; ++g) {
            const long long c0 = C * g / G, c1 = C * (g + 1) / G;
            cudaStream_t sg = g == 0 ? stream : set.side[g];
            if (g) {
                const cudaError_t e = cudaStreamWaitEvent(sg, set.fork, 0);
                if (e != cudaSuccess) { status = fail(LCT_ERR_CUDA, "cudaStreamWaitEvent(fork)", e); break; }
                forked = g;
            }
            GpuLauncher lg{sg, plan->device};
            lg.prefetch_ahead = plan->prefetch_ahead;
            lg.pdl = plan->pdl;
            const int rc = lct::run_chain(lg, t, M, N, (int)(c1 - c0), D, Tin, tbe[0], be_dev, (int)c0,
                                          in + (size_t)c0 * in_stride, out + (size_t)c0 * out_stride,
                                          s1 + (size_t)c0 * (M + 1) * N * N, s2 + (size_t)c0 * (M + 1) * 2 * N * N, backward,
                                          lct::kStageAll, minmax_keys);
            if (rc < 0) status = fail(LCT_ERR_UNSUPPORTED, "size not compiled");
            else if (rc) status = fail(LCT_ERR_CUDA, "kernel launch", lg.err);
        }
        for (int g = 1; g <= forked; ++g) {
            cudaError_t e = cudaEventRecord(set.join[g], set.side[g]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, set.join[g], 0);
            if (e != cudaSuccess) {
                cudaStreamSynchronize(set.side[g]);         // last resort: the join could not be expressed as an event
                if (status == LCT_OK) status = fail(LCT_ERR_CUDA, "side stream join", e);
            }
        }
        return status;
    }
    GpuLauncher l{stream, plan->device};
    l.events = events;
    l.prefetch_ahead = plan->prefetch_ahead;
    l.pdl = plan->pdl;
    for (long long c0 = 0; c0 < C; c0 += chunk) {
        const int cn = (int)((C - c0 < chunk) ? (C - c0) : chunk);
        const int rc = lct::run_chain(l, t, M, N, cn, D, Tin, tbe[0], be_dev, (int)c0,
                                      in + (size_t)c0 * in_stride, out + (size_t)c0 * out_stride, s1, s2, backward,
                                      lct::kStageAll, minmax_keys);
        if (rc < 0) return fail(LCT_ERR_UNSUPPORTED, "size not compiled");
        if (rc) return fail(LCT_ERR_CUDA, "kernel launch", l.err);
    }
    if (l.err != cudaSuccess) return fail(LCT_ERR_CUDA, "cudaEventRecord", l.err);
    return LCT_OK;
}

int lct_forward(const lct_plan* plan, const float* x, const int32_t* tbe, const int32_t* ten,
                int32_t B, int32_t D, int32_t Tin, float* y, void* ws, size_t ws_bytes, void* stream) {
    return run(plan, x, tbe, ten, B, D, Tin, y, ws, ws_bytes, stream, false);
}

int lct_backward(const lct_plan* plan, const float* gy, const int32_t* tbe, const int32_t* ten,
                 int32_t B, int32_t D, int32_t Tin, float* gx, void* ws, size_t ws_bytes, void* stream) {
    return run(plan, gy, tbe, ten, B, D, Tin, gx, ws, ws_bytes, stream, true);
}

int lct_forward_minmax(const lct_plan* plan, const float* x, const int32_t* tbe, const int32_t* ten,
                       int32_t B, int32_t D, int32_t Tin, float* y, void* minmax_keys,
                       void* ws, size_t ws_bytes, void* stream) {
    if (!minmax_keys || ((uintptr_t)minmax_keys & 7)) return fail(LCT_ERR_INVALID, "bad minmax buffer");
    return run(plan, x, tbe, ten, B, D, Tin, y, ws, ws_bytes, stream, false, nullptr,
               static_cast<unsigned long long*>(minmax_keys));
}

// blocks per channel for the grid-stride passes (more blocks -- a full wave -- measured slower for the min/max
// pass, 21 vs 16.5 us at 8 x 1 Mi elements, and no faster for the affine pass)
static unsigned reduce_blocks(long long elems, int /*channels*/) {
    long long b = (elems / 4 + 255) / 256;
    return (unsigned)(b < 1 ? 1 : (b > 64 ? 64 : b));
}

int lct_minmax(const float* x, int32_t channels, int64_t elems, void* keys, void* stream_) {
    if (!x || !keys || channels <= 0 || elems <= 0 || elems > 0xffffffffLL || ((uintptr_t)x & 15) || ((uintptr_t)keys & 7))
        return fail(LCT_ERR_INVALID, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    LCT_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t)channels * 2 * sizeof(unsigned long long), stream));
    lct::minmax_kernel<<<dim3(reduce_blocks(elems, channels), channels), 256, 0, stream>>>(x, static_cast<unsigned long long*>(keys), elems);
    LCT_CUDA(cudaGetLastError());
    return LCT_OK;
}

int lct_normalize_feature(const float* x, void* keys, float* out, int32_t channels, int64_t elems,
                          float scale, void* stream_) {
    if (!x || !keys || !out || channels <= 0 || elems <= 0 || (((uintptr_t)x | (uintptr_t)out) & 15))
        return fail(LCT_ERR_INVALID, "bad argument");
    lct::normalize_kernel<<<dim3(reduce_blocks(elems, channels), channels), 256, 0, (cudaStream_t)stream_>>>(
        x, out, static_cast<unsigned long long*>(keys), elems, scale);
    LCT_CUDA(cudaGetLastError());
    return LCT_OK;
}

int lct_normalize_feature_backward(const float* x, const float* gout, const void* keys, float* gx, void* sums,
                                   int32_t channels, int64_t elems, float scale, void* stream_) {
    if (!x || !gout || !keys || !gx || !sums || channels <= 0 || elems <= 0) return fail(LCT_ERR_INVALID, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    LCT_CUDA(cudaMemsetAsync(sums, 0, (size_t)channels * 2 * sizeof(double), stream));
    const dim3 grid(reduce_blocks(elems, channels), channels);
    lct::normalize_bwd_sums_kernel<<<grid, 256, 0, stream>>>(x, gout, static_cast<const unsigned long long*>(keys),
                                                             static_cast<double*>(sums), elems);
    lct::normalize_bwd_kernel<<<grid, 256, 0, stream>>>(gout, gx, static_cast<const unsigned long long*>(keys),
                                                        static_cast<const double*>(sums), elems, scale);
    LCT_CUDA(cudaGetLastError());
    return LCT_OK;
}

int lct_run_staged(const lct_plan* plan, const float* in, const int32_t* tbe, const int32_t* ten,
                   int32_t B, int32_t D, int32_t Tin, float* out, void* ws, size_t ws_bytes, void* stream,
                   int32_t backward, void* const* events6) {
    if (!events6) return fail(LCT_ERR_INVALID, "null events");
    return run(plan, in, tbe, ten, B, D, Tin, out, ws, ws_bytes, stream, backward != 0, events6);
}

int lct_bp_laplacian(const lct_plan* plan, const float* vol, float* out, int32_t channels,
                     const float* lapw, int32_t adjoint, void* stream_) {
    if (!plan || !vol || !out || !lapw || channels <= 0 || vol == out) return fail(LCT_ERR_INVALID, "bad argument");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(LCT_ERR_CUDA, "cudaSetDevice failed");
    lct::StencilWeights w;
    std::memcpy(w.w, lapw, sizeof(float) * 125);
    const int M = plan->M, N = plan->N;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t total = (size_t)channels * M * N * N;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    static const bool gather_only = [] { const char* e = std::getenv("LCT_BP_GATHER"); return e && std::atoi(e) != 0; }();
    if (gather_only || N % 4 != 0 || channels > 65535) {          // the plain gather kernels (measurement aid; any shape)
        if (adjoint) lct::laplacian_adjoint_kernel<<<blocks, threads, 0, stream>>>(vol, out, channels, M, N, w);
        else lct::laplacian_kernel<<<blocks, threads, 0, stream>>>(vol, out, channels, M, N, w);
        LCT_CUDA(cudaGetLastError());
        return LCT_OK;
    }
    lct::LapTiledParams p{};
    p.in = vol; p.out = out; p.M = M; p.N = N;
    p.rows = N < 16 ? N : 16;
    while (p.rows * N / 4 > 512) p.rows /= 2;
    // enough blocks for about four waves of one block per SM and warp set, but at least 16 planes per block: a block
    // reads four planes more than it writes
    const long long per_plane = (long long)(N / p.rows) * channels;
    long long nchunks = (592 + per_plane - 1) / per_plane;
    if (nchunks > M / 16) nchunks = M / 16;
    if (nchunks < 1) nchunks = 1;
    p.chunk = (int)((M + nchunks - 1) / nchunks);
    p.zero_pad = adjoint ? 1 : 0;
    if (adjoint) for (int k = 0; k < 125; ++k) p.w.w[k] = w.w[124 - k];       // taps flipped in all three axes
    else p.w = w;
    const dim3 grid(N / p.rows, (M + p.chunk - 1) / p.chunk, channels);
    const int nthreads = p.rows * N / 4, slots = (p.rows + 2 * lct::kLapHalo) * (N + 2 * lct::kLapHalo);
    const int ke = (slots + nthreads - 1) / nthreads;
    const size_t smem = lct::lap_smem_bytes(N, p.rows);
    if (ke <= 6) lct::laplacian_tiled_kernel<6><<<grid, nthreads, smem, stream>>>(p);
    else if (ke <= 7) lct::laplacian_tiled_kernel<7><<<grid, nthreads, smem, stream>>>(p);
    else if (ke <= 10) lct::laplacian_tiled_kernel<10><<<grid, nthreads, smem, stream>>>(p);
    else return fail(LCT_ERR_UNSUPPORTED, "laplacian tile shape");
    LCT_CUDA(cudaGetLastError());
    if (adjoint) {                                                 // the boundary shell, where clamped taps fold onto a voxel
        const long long shell = (2LL * N * N + 2LL * (M - 2) * N + 2LL * (M - 2) * (N - 2)) * channels;
        lct::laplacian_adjoint_shell_kernel<<<(unsigned)((shell + threads - 1) / threads), threads, 0, stream>>>(vol, out, channels, M, N, w);
        LCT_CUDA(cudaGetLastError());
    }
    return LCT_OK;
}

static bool skip_args_ok(int32_t B, int32_t D, int32_t T, int32_t N) {
    return B > 0 && B <= 65535 && D > 0 && T > 0 && N >= 4 && (N % 4) == 0;
}

extern "C++" {
// launch shape of one skip-branch pass on the current device (the SM count is looked up once per device)
struct SkipLaunch { bool ring; int chunk; dim3 grid; size_t smem; };

static SkipLaunch skip_launch(int32_t B, int32_t T, int32_t N, bool single_channel_window) {
    static int sms[64] = {0};
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
        if (sms[dev] == 0 && (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0))
            sms[dev] = 148;
        n = sms[dev];
    }
    SkipLaunch l;
    l.ring = single_channel_window && lct::skip_ring_ok(N);
    l.chunk = lct::skip_chunk(B, T, N, n, l.ring);
    l.grid = l.ring ? lct::skip_ring_grid(B, T, N, l.chunk) : lct::skip_grid(B, T, N, l.chunk);
    l.smem = l.ring ? lct::skip_ring_smem(N) : 0;
    return l;
}

template <int MODE> static cudaError_t skip_run(const SkipLaunch& l, const lct::SkipParams& p, cudaStream_t stream) {
    if (l.ring) {
        static std::mutex mu;
        static bool ready[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lock(mu);
            if (dev >= 0 && dev < 64 && !ready[dev]) {
                cudaError_t e = cudaFuncSetAttribute(lct::skip_ring_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                if (e != cudaSuccess) return e;
                ready[dev] = true;
            }
        }
        lct::skip_ring_kernel<MODE><<<l.grid, lct::kSkipThreads, l.smem, stream>>>(p);
    } else {
        lct::skip_kernel<MODE><<<l.grid, lct::kSkipThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}
}  // extern "C++"

size_t lct_skip_workspace_bytes(int32_t B, int32_t T, int32_t N) {
    if (!skip_args_ok(B, 1, T, N)) return 0;
    const SkipLaunch l = skip_launch(B, T, N, true);
    return (size_t)l.grid.x * l.grid.y * l.grid.z * 27 * sizeof(float);
}

int lct_skip_sum(const float* feat, const float* x, const float* w, int32_t B, int32_t D, int32_t T, int32_t N,
                 float* out, void* stream_) {
    if (!feat || !x || !w || !out || !skip_args_ok(B, D, T, N)) return fail(LCT_ERR_INVALID, "bad argument");
    if (((uintptr_t)feat | (uintptr_t)x | (uintptr_t)out) & 15) return fail(LCT_ERR_INVALID, "buffers must be 16-byte aligned");
    if (x == out) return fail(LCT_ERR_INVALID, "x must not alias out");
    const SkipLaunch l = skip_launch(B, T, N, true);
    lct::SkipParams p{x, feat, out, w, D, T, N, l.chunk};
    LCT_CUDA(skip_run<lct::kSkipForward>(l, p, (cudaStream_t)stream_));
    return LCT_OK;
}

int lct_skip_sum_backward(const float* g, const float* x, const float* w, int32_t B, int32_t D, int32_t T, int32_t N,
                          float* gx, float* gw, void* ws, size_t ws_bytes, void* stream_) {
    if (!g || !skip_args_ok(B, D, T, N) || (!gx && !gw)) return fail(LCT_ERR_INVALID, "bad argument");
    if (((uintptr_t)g | (uintptr_t)x | (uintptr_t)gx) & 15) return fail(LCT_ERR_INVALID, "buffers must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (gx) {
        if (!w || gx == g) return fail(LCT_ERR_INVALID, "bad argument");
        const SkipLaunch l = skip_launch(B, T, N, D == 1);          // the window is sum_d g[b, d]: staged copies need D == 1
        lct::SkipParams p{g, nullptr, gx, w, D, T, N, l.chunk};
        LCT_CUDA(skip_run<lct::kSkipDataGrad>(l, p, stream));
    }
    if (gw) {
        const SkipLaunch l = skip_launch(B, T, N, true);
        const size_t blocks = (size_t)l.grid.x * l.grid.y * l.grid.z;
        if (!x || !ws || ws_bytes < blocks * 27 * sizeof(float)) return fail(LCT_ERR_WORKSPACE, "workspace too small");
        lct::SkipParams p{x, g, static_cast<float*>(ws), nullptr, D, T, N, l.chunk};
        LCT_CUDA(skip_run<lct::kSkipWeightGrad>(l, p, stream));
        lct::skip_weight_reduce_kernel<<<1, 27 * 32, 0, stream>>>(static_cast<const float*>(ws), (int)blocks, gw);
        LCT_CUDA(cudaGetLastError());
    }
    return LCT_OK;
}

int lct_forward_host(const lct_plan* plan, const float* x_host, const int32_t* tbe, const int32_t* ten,
                     int32_t B, int32_t D, int32_t Tin, float* y_host, void* stream_) {
    if (!plan || !x_host || !y_host) return fail(LCT_ERR_INVALID, "null argument");
    if (B <= 0 || D <= 0 || Tin <= 0 || Tin > plan->M) return fail(LCT_ERR_INVALID, "bad shape");
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(LCT_ERR_CUDA, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t C = (size_t)B * D, NN = (size_t)plan->N * plan->N;
    const size_t xb = C * Tin * NN * sizeof(float), yb = C * plan->M * NN * sizeof(float);
    const size_t wb = lct_plan_workspace_bytes(plan, (int32_t)C);
    unsigned char* dev = nullptr;
    LCT_CUDA(cudaMallocAsync((void**)&dev, align_up(xb, 256) + align_up(yb, 256) + wb, stream));
    float* dx = reinterpret_cast<float*>(dev);
    float* dy = reinterpret_cast<float*>(dev + align_up(xb, 256));
    void* ws = dev + align_up(xb, 256) + align_up(yb, 256);
    int rc = LCT_OK;
    cudaError_t e = cudaMemcpyAsync(dx, x_host, xb, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
        rc = lct_forward(plan, dx, tbe, ten, B, D, Tin, dy, ws, wb, stream);
        if (rc == LCT_OK) e = cudaMemcpyAsync(y_host, dy, yb, cudaMemcpyDeviceToHost, stream);
    }
    cudaFreeAsync(dev, stream);
    cudaError_t es = cudaStreamSynchronize(stream);
    if (rc != LCT_OK) return rc;
    if (e != cudaSuccess) return fail(LCT_ERR_CUDA, "host copy", e);
    if (es != cudaSuccess) return fail(LCT_ERR_CUDA, "stream synchronize", es);
    return LCT_OK;
}

}  // extern "C"
