/*
 * hiddenpose_lct.h -- C ABI of the B200 (sm_100a) light-cone-transform library.
 *
 * Drop-in boundary for the LCT layer of Hagtaril/HiddenPose.  The reference has
 * no FFI: its layer is a Python nn.Module over PyTorch library calls.  Each entry
 * point below names the piece of the reference it replaces:
 *
 *   lct_plan_create    <- lct.parpareparam + lct.todev      models/tflct.py:32-92
 *                         (== LCT._parpareparam + LCT.todev models/feature_propagation.py:71-109,173-184)
 *                         operators come from              utils/helper.py:35-69,72-125
 *   lct_forward        <- lct.forward                       models/tflct.py:94-179
 *                         (== LCT.forward                   models/feature_propagation.py:186-257)
 *   lct_backward       <- the autograd-derived backward of that forward (the reference has no
 *                         explicit backward; it is the adjoint of the same linear chain,
 *                         SURVEY.md section 3.5)
 *   lct_bp_laplacian   <- the method=='bp' tail             models/tflct.py:164-175
 *   lct_normalize_feature[_backward], lct_minmax, lct_forward_minmax
 *                      <- normalize_feature                 models/feature_propagation.py:273-286
 *   lct_skip_sum[_backward]
 *                      <- FeatureExtraction's skip branch   models/feature_extraction.py:141-145,166-171
 *                         (the op that produces the LCT's input, NlosPose.py:51-53)
 *
 * Plain C types only: pointers, sizes, integer return codes (0 = success).  Nothing
 * throws across this boundary.  Device pointers are raw CUDA device addresses; `stream`
 * is a cudaStream_t passed as void*.  The caller owns all data buffers and the
 * workspace; a plan owns only immutable device constants plus a few internal side streams,
 * so one plan may be shared by threads and streams as long as concurrent calls use distinct
 * workspaces.  The channel groups of one call run on side streams that belong to the caller's
 * stream (four sets per plan, handed out per distinct caller stream; callers beyond four share
 * sets, which stays correct but orders the sharers' kernels after one another).
 * Calls are asynchronous on `stream`; they never allocate, synchronise or read host memory
 * after returning (lct_forward_host is the exception and says so), so they may be captured in
 * a CUDA graph -- per-sample windows included: the window table travels as kernel parameters.
 * If a call fails after some of its kernels were queued, everything it queued on internal
 * streams is ordered before the caller's stream again before the error code is returned.
 */
#ifndef HIDDENPOSE_LCT_H_
#define HIDDENPOSE_LCT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCT_OK 0
#define LCT_ERR_INVALID 1      /* bad argument (shape, window, null pointer) */
#define LCT_ERR_UNSUPPORTED 2  /* M or N outside the compiled set */
#define LCT_ERR_CUDA 3         /* a CUDA runtime call failed; see lct_last_error() */
#define LCT_ERR_WORKSPACE 4    /* workspace smaller than lct_plan_workspace_bytes(plan, 1) */
#define LCT_ERR_NOMEM 5

#define LCT_ABI_VERSION 3

/* lct_desc.flags */
#define LCT_FLAG_NO_PLANE_FUSION 1   /* keep K2/K3/K4 as three kernels even when the plane fits on chip */
#define LCT_FLAG_FULL_FILTER 2       /* store the whole half-spectrum filter even when its mirror symmetry in kh / kw
                                        would allow the quarter layout (a quarter of the filter bytes per launch) */

typedef struct lct_plan lct_plan;

/* Operators, all in HOST memory, copied to the device by lct_plan_create. */
typedef struct lct_desc {
    int32_t time_bins;          /* M: `crop` / `time_size`; power of two in [32, 512]            */
    int32_t spatial;            /* N: `spatial` / `image_size` (H == W); power of two in [8, 256] */
    int32_t device;             /* CUDA device ordinal                                            */
    int32_t reserved;           /* flags: 0 or LCT_FLAG_*                                         */
    const int32_t* mtx_rowptr;  /* M+1   CSR row pointers of mtx (helper.py:35-69)                */
    const int32_t* mtx_colidx;  /* nnz                                                            */
    const float* mtx_vals;      /* nnz                                                            */
    const float* falloff;       /* M     gridz**4 | **2 (tflct.py:123-127); NULL = no falloff     */
    const float* filter_half;   /* (M+1, 2N, 2N, 2) interleaved re/im: planes kt = 0..M of invpsf
                                   (tflct.py:57-65), unscaled; or NULL to have the library build the
                                   filter on the device from the PSF support below                */
    /* PSF support (utils/helper.py:72-125): the non-zero voxels of definePsf's (2M, 2N, 2N) volume,
       after its roll/transpose.  Only read when filter_half is NULL.                              */
    const int32_t* psf_z;       /* psf_count  time index of each voxel, in [0, 2M)                */
    const int32_t* psf_yx;      /* psf_count  y * 2N + x of each voxel                            */
    int32_t psf_count;
    float psf_value;            /* common voxel value 1/sqrt(count) (helper.py:112)               */
    float snr;                  /* tflct.py:42                                                    */
    int32_t method_bp;          /* 0: 'lct' Wiener inverse (tflct.py:60); 1: 'bp' conj(fpsf) (:62) */
} lct_desc;

int lct_abi_version(void);
const char* lct_error_string(int code);
/* Detail of the last failure on the calling thread ("" if none). */
const char* lct_last_error(void);

int lct_plan_create(const lct_desc* desc, lct_plan** out);
void lct_plan_destroy(lct_plan* plan);
int32_t lct_plan_time_bins(const lct_plan* plan);
int32_t lct_plan_spatial(const lct_plan* plan);

/* Bytes of device scratch that let `channels` (= B*D) volumes run as one batch.  Any
 * workspace >= lct_plan_workspace_bytes(plan, 1) is accepted; smaller-than-full
 * workspaces make the call process the channels in several batches. */
size_t lct_plan_workspace_bytes(const lct_plan* plan, int32_t channels);

/*
 * y[b,d,:,:,:] = LCT(x[b,d,:,:,:]) with x placed at time bins [tbe[b], ten[b]).
 *   x   device, float32, (B, D, Tin, N, N) contiguous
 *   y   device, float32, (B, D, M,   N, N) contiguous
 *   tbe, ten  HOST arrays of B int32; 0 <= tbe[b], ten[b] <= M, ten[b]-tbe[b] == Tin
 */
int lct_forward(const lct_plan* plan, const float* x, const int32_t* tbe, const int32_t* ten,
                int32_t B, int32_t D, int32_t Tin, float* y,
                void* workspace, size_t workspace_bytes, void* stream);

/*
 * gx = (d LCT / d x)^T gy : gradient w.r.t. x of sum(y * gy).
 *   gy  device, float32, (B, D, M,   N, N);   gx  device, float32, (B, D, Tin, N, N)
 */
int lct_backward(const lct_plan* plan, const float* gy, const int32_t* tbe, const int32_t* ten,
                 int32_t B, int32_t D, int32_t Tin, float* gx,
                 void* workspace, size_t workspace_bytes, void* stream);

/*
 * normalize_feature, the op NlosPose applies to the LCT volume (models/feature_propagation.py:273-286,
 * NlosPose.py:54):  out = (x - min) / (max(x - min) + 1e-15) * scale  per channel of `elems` values.
 * min / max travel as two 64-bit keys per channel ({value, position}, device memory, 16 bytes per
 * channel).  lct_forward_minmax is lct_forward that also reduces the keys of the volume it writes
 * (in its last kernel, while the values are in registers) -- values only: the position halves of its
 * keys read "unknown" (0xffffffff) until lct_normalize_feature, which sees every value anyway, has filled
 * them in (it updates the key buffer in place; a NaN or infinity in the volume makes lct_forward_minmax
 * resolve the positions itself).  lct_minmax computes complete keys for any tensor.
 * The backward pass needs `sums`: 16 bytes of device scratch per channel.
 */
int lct_forward_minmax(const lct_plan* plan, const float* x, const int32_t* tbe, const int32_t* ten,
                       int32_t B, int32_t D, int32_t Tin, float* y, void* minmax_keys,
                       void* workspace, size_t workspace_bytes, void* stream);
int lct_minmax(const float* x, int32_t channels, int64_t elems, void* minmax_keys, void* stream);
int lct_normalize_feature(const float* x, void* minmax_keys, float* out, int32_t channels, int64_t elems,
                          float scale, void* stream);
int lct_normalize_feature_backward(const float* x, const float* gout, const void* minmax_keys, float* gx,
                                   void* sums, int32_t channels, int64_t elems, float scale, void* stream);

/*
 * Measurement hook: lct_forward (backward == 0) or lct_backward (backward != 0) with six
 * caller-created cudaEvent_t recorded on `stream`: events6[i] before kernel i (time-forward,
 * row-forward, column-filter, row-inverse, time-inverse) and events6[5] after the last one.
 * Needs a workspace for the whole batch.  Same results as the plain calls, but the kernels run
 * back to back on `stream` only (the plain calls split the channels over two internal streams so
 * that consecutive kernels of different channel groups overlap; set LCT_STREAM_GROUPS=1 to disable).
 */
int lct_run_staged(const lct_plan* plan, const float* in, const int32_t* tbe, const int32_t* ten,
                   int32_t B, int32_t D, int32_t Tin, float* out,
                   void* workspace, size_t workspace_bytes, void* stream,
                   int32_t backward, void* const* events6);

/*
 * method=='bp' tail (tflct.py:164-175): out = conv3d(replicate_pad(vol, 2), lapw 5x5x5),
 * then out[:, 0] = 0.  `adjoint` != 0 applies the transpose (for the backward pass).
 *   vol, out  device, float32, (C, M, N, N), must not alias;  lapw  HOST, 125 floats.
 */
int lct_bp_laplacian(const lct_plan* plan, const float* vol, float* out, int32_t channels,
                     const float* lapw, int32_t adjoint, void* stream);

/*
 * Skip branch of FeatureExtraction (feature_extraction.py:166-171), the op that feeds the LCT:
 *     out[b, d] = feat[b, d] + conv3d(x[b], w, stride 1, padding 1)         (zero padding,
 *     cross-correlation as F.conv3d; the one-channel result is broadcast over feat's D channels)
 *   feat, out  device, float32, (B, D, T, N, N); out may be feat (in place)
 *   x          device, float32, (B, 1, T, N, N); must not alias out
 *   w          DEVICE, 27 floats, the learnable (1,1,3,3,3) parameter in (t, y, x) order
 * N must be a multiple of 4 and the pointers 16-byte aligned.  Runs on the current device.
 *
 * lct_skip_sum_backward: given g = d loss / d out (B, D, T, N, N),
 *     gx[b]  = conv3d^T(sum_d g[b, d], w)            (B, 1, T, N, N), skipped when gx == NULL
 *     gw[k]  = sum_{b,d,p} g[b,d,p] * x[b, p + k - 1]  27 floats,     skipped when gw == NULL
 * (d loss / d feat is g itself).  gw needs lct_skip_workspace_bytes(B, T, N) bytes of device scratch;
 * the reduction order is fixed, so results are reproducible run to run.
 */
size_t lct_skip_workspace_bytes(int32_t B, int32_t T, int32_t N);
int lct_skip_sum(const float* feat, const float* x, const float* w, int32_t B, int32_t D, int32_t T, int32_t N,
                 float* out, void* stream);
int lct_skip_sum_backward(const float* g, const float* x, const float* w, int32_t B, int32_t D, int32_t T, int32_t N,
                          float* gx, float* gw, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Host-buffer convenience: copies x from host memory, runs lct_forward, copies y back and
 * SYNCHRONISES `stream`.  Allocates its device scratch from the stream-ordered allocator.
 * Pinned host buffers make the copies asynchronous to each other.
 */
int lct_forward_host(const lct_plan* plan, const float* x_host, const int32_t* tbe, const int32_t* ten,
                     int32_t B, int32_t D, int32_t Tin, float* y_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIDDENPOSE_LCT_H_ */
