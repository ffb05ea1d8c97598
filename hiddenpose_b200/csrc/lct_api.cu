// C ABI of the B200 LCT library (see include/hiddenpose_lct.h).
// Host side only: plan construction, workspace carving, kernel launches.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "hiddenpose_lct.h"
#include "lct_chain.cuh"
#include "lct_stencil.cuh"

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
    void mark(int i) {
        if (events && err == cudaSuccess) err = cudaEventRecord((cudaEvent_t)events[i], stream);
    }
    template <class K> int launch(const lct::Params& p) {
        static bool attr_set[kMaxDevices] = {};
        auto kern = lct::lct_kernel<K>;
        if (K::kSmem > 48 * 1024 && !attr_set[device]) {
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::kSmem);
            if (err != cudaSuccess) return 1;
            attr_set[device] = true;
        }
        int gx, gy;
        K::grid(p, gx, gy);
        kern<<<dim3(gx, gy), K::kThreads, K::kSmem, stream>>>(p, K::iterations(p));
        err = cudaGetLastError();
        return err == cudaSuccess ? 0 : 1;
    }
};

template <class T> int to_device(const T* host, size_t count, T** out) {
    LCT_CUDA(cudaMalloc((void**)out, count * sizeof(T)));
    LCT_CUDA(cudaMemcpy(*out, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return LCT_OK;
}

__global__ void scale_filter_kernel(float2* f, size_t n, float s) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 v = f[i];
        f[i] = make_float2(v.x * s, v.y * s);
    }
}

}  // namespace

struct lct_plan {
    int M = 0, N = 0, device = 0;
    int *mtx_rowptr = nullptr, *mtx_colidx = nullptr, *mtxi_rowptr = nullptr, *mtxi_colidx = nullptr;
    float *mtx_vals = nullptr, *mtx_vals_falloff = nullptr, *mtxi_vals = nullptr, *mtxi_vals_falloff = nullptr;
    float2* filt = nullptr;
    size_t per_channel_bytes() const { return (size_t)(M + 1) * N * N * sizeof(float2) * 3; }
    lct::ChainTables tables() const {
        return lct::ChainTables{mtx_rowptr, mtx_colidx, mtx_vals_falloff, mtx_vals,
                                mtxi_rowptr, mtxi_colidx, mtxi_vals, mtxi_vals_falloff, filt};
    }
};

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
    cudaFree(plan->mtx_rowptr); cudaFree(plan->mtx_colidx); cudaFree(plan->mtxi_rowptr); cudaFree(plan->mtxi_colidx);
    cudaFree(plan->mtx_vals); cudaFree(plan->mtx_vals_falloff); cudaFree(plan->mtxi_vals); cudaFree(plan->mtxi_vals_falloff);
    cudaFree(plan->filt);
    delete plan;
}

int lct_plan_create(const lct_desc* d, lct_plan** out) {
    if (!d || !out) return fail(LCT_ERR_INVALID, "null descriptor");
    *out = nullptr;
    if (!d->mtx_rowptr || !d->mtx_colidx || !d->mtx_vals || !d->filter_half) return fail(LCT_ERR_INVALID, "null operator table");
    const int M = d->time_bins, N = d->spatial;
    if (!lct::supported_M(M) || !lct::supported_N(N)) return fail(LCT_ERR_UNSUPPORTED, "size not compiled");
    if (d->device < 0 || d->device >= kMaxDevices) return fail(LCT_ERR_INVALID, "bad device ordinal");
    DeviceGuard guard(d->device);
    if (!guard.ok) return fail(LCT_ERR_CUDA, "cudaSetDevice failed");
    int rc = upload_twiddles(d->device);
    if (rc) return rc;

    // validate the CSR and build its transpose (mtxi = mtx^T, helper.py:61) on the host
    const int nnz = d->mtx_rowptr[M];
    if (d->mtx_rowptr[0] != 0 || nnz <= 0) return fail(LCT_ERR_INVALID, "bad CSR row pointers");
    for (int i = 0; i < M; ++i)
        if (d->mtx_rowptr[i + 1] < d->mtx_rowptr[i]) return fail(LCT_ERR_INVALID, "CSR row pointers not monotone");
    for (int e = 0; e < nnz; ++e)
        if (d->mtx_colidx[e] < 0 || d->mtx_colidx[e] >= M) return fail(LCT_ERR_INVALID, "CSR column index out of range");
    std::vector<float> fall(M, 1.0f);
    if (d->falloff) std::memcpy(fall.data(), d->falloff, sizeof(float) * M);
    std::vector<float> vals_f(nnz);
    std::vector<int> t_rowptr(M + 1, 0), t_colidx(nnz);
    std::vector<float> t_vals(nnz), t_vals_f(nnz);
    for (int e = 0; e < nnz; ++e) {
        vals_f[e] = d->mtx_vals[e] * fall[d->mtx_colidx[e]];
        t_rowptr[d->mtx_colidx[e] + 1]++;
    }
    for (int j = 0; j < M; ++j) t_rowptr[j + 1] += t_rowptr[j];
    {
        std::vector<int> cursor(t_rowptr.begin(), t_rowptr.end() - 1);
        for (int i = 0; i < M; ++i)
            for (int e = d->mtx_rowptr[i]; e < d->mtx_rowptr[i + 1]; ++e) {
                const int j = d->mtx_colidx[e], dst = cursor[j]++;
                t_colidx[dst] = i;
                t_vals[dst] = d->mtx_vals[e];
                t_vals_f[dst] = d->mtx_vals[e] * fall[j];
            }
    }

    lct_plan* p = new (std::nothrow) lct_plan();
    if (!p) return fail(LCT_ERR_NOMEM, "plan allocation");
    p->M = M; p->N = N; p->device = d->device;
    const size_t nfilt = (size_t)(M + 1) * 4 * N * N;
#define LCT_TRY(expr) do { rc = (expr); if (rc) { lct_plan_destroy(p); return rc; } } while (0)
    LCT_TRY(to_device(d->mtx_rowptr, (size_t)M + 1, &p->mtx_rowptr));
    LCT_TRY(to_device(d->mtx_colidx, (size_t)nnz, &p->mtx_colidx));
    LCT_TRY(to_device(d->mtx_vals, (size_t)nnz, &p->mtx_vals));
    LCT_TRY(to_device(vals_f.data(), (size_t)nnz, &p->mtx_vals_falloff));
    LCT_TRY(to_device(t_rowptr.data(), (size_t)M + 1, &p->mtxi_rowptr));
    LCT_TRY(to_device(t_colidx.data(), (size_t)nnz, &p->mtxi_colidx));
    LCT_TRY(to_device(t_vals.data(), (size_t)nnz, &p->mtxi_vals));
    LCT_TRY(to_device(t_vals_f.data(), (size_t)nnz, &p->mtxi_vals_falloff));
    LCT_TRY(to_device(reinterpret_cast<const float2*>(d->filter_half), nfilt, &p->filt));
#undef LCT_TRY
    // fold torch.ifft's 1/(2M*2N*2N) (tflct.py:151) into the filter; a power of two, so exact
    scale_filter_kernel<<<1024, 256>>>(p->filt, nfilt, 1.0f / (8.0f * (float)M * (float)N * (float)N));
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
               void* const* events = nullptr) {
    if (!plan || !in || !out || !tbe || !ten || !ws) return fail(LCT_ERR_INVALID, "null argument");
    if (B <= 0 || D <= 0 || Tin <= 0 || Tin > plan->M) return fail(LCT_ERR_INVALID, "bad shape");
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
        LCT_CUDA(cudaMemcpyAsync(be_dev, tbe, sizeof(int) * B, cudaMemcpyHostToDevice, stream));
    }
    float2* s1 = reinterpret_cast<float2*>(base + header);
    float2* s2 = s1 + (size_t)chunk * (M + 1) * N * N;
    const size_t in_stride = (size_t)(backward ? M : Tin) * N * N;
    const size_t out_stride = (size_t)(backward ? Tin : M) * N * N;
    if (events && chunk < C) return fail(LCT_ERR_WORKSPACE, "stage events need a workspace for the whole batch");
    GpuLauncher l{stream, plan->device};
    l.events = events;
    const lct::ChainTables t = plan->tables();
    for (long long c0 = 0; c0 < C; c0 += chunk) {
        const int cn = (int)((C - c0 < chunk) ? (C - c0) : chunk);
        const int rc = lct::run_chain(l, t, M, N, cn, D, Tin, tbe[0], be_dev, (int)c0,
                                      in + (size_t)c0 * in_stride, out + (size_t)c0 * out_stride, s1, s2, backward);
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
    const size_t total = (size_t)channels * plan->M * plan->N * plan->N;
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (adjoint)
        lct::laplacian_adjoint_kernel<<<blocks, threads, 0, (cudaStream_t)stream_>>>(vol, out, channels, plan->M, plan->N, w);
    else
        lct::laplacian_kernel<<<blocks, threads, 0, (cudaStream_t)stream_>>>(vol, out, channels, plan->M, plan->N, w);
    LCT_CUDA(cudaGetLastError());
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
