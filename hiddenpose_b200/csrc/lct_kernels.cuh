// The LCT kernels (sm_100a), written as sequences of barrier-separated phases so that the very
// same code can be stepped thread by thread on the CPU by tests/emu (LCT_EMULATE; test
// infrastructure only, never a product path).
//
// Data flow for C = B*D channels, M time bins, N x N spatial grid
// (reference: /root/reference/models/tflct.py:94-179):
//
//   K1 TimeFwd   x (C,Tin,N,N) f32 -> S1 (C,M+1,N,N) c64
//                time window (tflct.py:104-110), falloff (:123-127), sqrt(t)
//                resample as a banded gather (:135-138), zero-extend to 2M (:140)
//                and real FFT along T (first axis of :144), half spectrum.
//   K2 RowFwd    S1 -> S2 (C,M+1,2N,N): zero-extended FFT along H (:144).
//   K3 ColFilter S2 in place: zero-extended FFT along W, Wiener multiply
//                (:145-150), inverse FFT along W, crop to N (:151,153).
//   K4 RowInv    S2 -> S1 (C,M+1,N,N): inverse FFT along H, crop (:151,153).
//   K5 TimeInv   S1 -> y (C,Tout,N,N) f32: Hermitian inverse FFT along T, crop
//                to M, real part (:153), inverse resample mtxi as a banded gather
//                (:156-159) [+ falloff and window crop for the backward pass]; optionally
//                reduces the volume's per-channel min/max for normalize_feature.
//   PlaneFilter  K2 + K3 + K4 in one kernel with the (c, kt) plane resident in shared memory;
//                used whenever the plane fits (N <= 64).
//   RowFwdSplit / ColFilterSplit / RowInvSplit
//                K2 / K3 / K4 for 512-point lines (N = 256): the zero-extended transform as two
//                256-point transforms by output parity (16-wide butterflies instead of 32-wide).
//
// The backward pass is the same chain with conj(filter), the falloff moved
// from K1's input to K5's output and the window cut out at the end
// (SURVEY.md section 3.5).
#pragma once

#include <limits>
#include <type_traits>

#include "lct_fft.cuh"
#include "lct_tables.h"

#ifndef LCT_TIME_PRELOAD
#define LCT_TIME_PRELOAD 1      // the persistent K1 requests its tables and first tile from the kernel driver
#endif
#ifndef LCT_K1_ASYNC_TILE_256
#define LCT_K1_ASYNC_TILE_256 1
#endif
#ifndef LCT_K1_BLOCKS_256
#define LCT_K1_BLOCKS_256 4      // blocks per SM the 256-thread time kernels (M = 128) are compiled for (4: 64 registers)
#endif
#ifndef LCT_K5_BLOCKS_256
#define LCT_K5_BLOCKS_256 4
#endif

namespace lct {

#ifdef LCT_EMULATE
extern float2 h_tw[kTwN];
struct TwConst {
    static inline float2 get(int i) { return h_tw[i]; }
    static inline float2 mul(float2 a, int i) { return cmul(a, get(i)); }
    static inline float2 mulc(float2 a, int i) { return cmulc(a, get(i)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) { return mul(a, (k * lo) * (kTwN / Ls)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) { return mulc(a, (k * lo) * (kTwN / Ls)); }
};
struct TwGlobal : TwConst {
    static constexpr size_t kBytes = 0;
    static inline void fill(unsigned char*, int, int) {}
};
#define LCT_LDG(p) (*(p))
#define LCT_LDCG(p) (*(p))
#else
__constant__ float2 c_tw[kTwN];          // exp(-2*pi*i*j/1024), built in double on the host
__device__ float2 g_tw[kTwN];            // same table in global memory for lane-divergent lookups
struct TwConst {
    static LCT_DEV float2 get(int i) { return c_tw[i]; }
    // scalar multiplies here: in the time kernels the packed form measured slower (register pairing)
    static LCT_DEV float2 mul(float2 a, int i) { return cmul_s(a, get(i)); }      // a * w^i
    static LCT_DEV float2 mulc(float2 a, int i) { return cmulc_s(a, get(i)); }    // a * conj(w^i)
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) { return mul(a, (k * lo) * (kTwN / Ls)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) { return mulc(a, (k * lo) * (kTwN / Ls)); }
};
struct TwGlobal {
    static constexpr size_t kBytes = 0;
    static LCT_DEV void fill(unsigned char*, int, int) {}
    static LCT_DEV float2 get(int i) { return __ldg(&g_tw[i]); }
    static LCT_DEV float2 mul(float2 a, int i) { return cmul(a, get(i)); }
    static LCT_DEV float2 mulc(float2 a, int i) { return cmulc(a, get(i)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) { return mul(a, (k * lo) * (kTwN / Ls)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) { return mulc(a, (k * lo) * (kTwN / Ls)); }
};
#define LCT_LDG(p) __ldg(p)
#define LCT_LDCG(p) __ldcg(p)           // L2 only: a line that is read once has no use for L1 (measured per kernel: it pays in
                                        // the H-axis kernels, changes nothing in the plane kernel; evict-first stores (__stcs) of
                                        // K5's output made that kernel 18 % slower at M = 128 and changed nothing elsewhere)
#endif

// Constant-bank twiddles behind the same interface as TwShared (no table, nothing to fill).
struct TwNone {
    static constexpr size_t kBytes = 0;
    static LCT_DEV float2 get(int i) { return TwConst::get(i); }
    static LCT_DEV float2 mul(float2 a, int i) { return TwConst::mul(a, i); }
    static LCT_DEV float2 mulc(float2 a, int i) { return TwConst::mulc(a, i); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) { return mul(a, (k * lo) * (kTwN / Ls)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) { return mulc(a, (k * lo) * (kTwN / Ls)); }
    static LCT_DEV void fill(unsigned char*, int, int) {}
};

// Twiddles from a block-local copy of the table at the very start of dynamic shared memory
// (entries w_L^j, j < L): a warp-uniform LDS broadcast has a shorter, steadier latency than an
// indexed constant-bank load, and these loads sit on the critical path of every butterfly.
template <int L> struct TwShared {
    // (a table that also held the conjugates, for two-instruction packed multiplies, was measured
    //  slower: its 128-bit broadcast loads cost twice the shared-memory pipe time of these 64-bit ones)
    static constexpr int kEntries = L;
    static constexpr size_t kBytes = (size_t)L * sizeof(float2);
#ifdef LCT_EMULATE
    static inline float2 get(int i) { return h_tw[i]; }
    static inline void fill(unsigned char*, int, int) {}
#else
    static LCT_DEV float2 get(int i) {
        extern __shared__ __align__(128) unsigned char lct_dyn_smem[];
        return reinterpret_cast<const float2*>(lct_dyn_smem)[i / (kTwN / L)];
    }
    static LCT_DEV void fill(unsigned char* smem, int tid, int nthreads) {
        // from the global copy of the table: every lane wants a different entry, which the constant bank would
        // serialise (32 passes per warp -- 6 % of the plane kernel's time went into this fill); one coalesced load here
        for (int j = tid; j < L; j += nthreads) reinterpret_cast<float2*>(smem)[j] = __ldg(&g_tw[j * (kTwN / L)]);
    }
#endif
    static LCT_DEV float2 mul(float2 a, int i) { return cmul(a, get(i)); }
    static LCT_DEV float2 mulc(float2 a, int i) { return cmulc(a, get(i)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) { return mul(a, (k * lo) * (kTwN / Ls)); }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) { return mulc(a, (k * lo) * (kTwN / Ls)); }
};

// Stage-0 twiddles of a line plan, laid out [k][lo] at the start of shared memory: the lanes of a
// line (consecutive lo) read consecutive entries, so the lane-divergent lookup is conflict-free.
template <class P> struct TwLine {
    static constexpr int L = P::L, STR0 = P::st(0);
    static constexpr size_t kBytes = (size_t)L * sizeof(float2);
#ifdef LCT_EMULATE
    static inline float2 at(int k, int lo) { return h_tw[(k * lo) * (kTwN / L)]; }
    static inline void fill(unsigned char*, int, int) {}
#else
    static LCT_DEV float2 at(int k, int lo) {
        extern __shared__ __align__(128) unsigned char lct_dyn_smem[];
        return reinterpret_cast<const float2*>(lct_dyn_smem)[k * STR0 + lo];
    }
    static LCT_DEV void fill(unsigned char* smem, int tid, int nthreads) {
        for (int j = tid; j < L; j += nthreads) {
            const int k = j / STR0, lo = j % STR0;
            reinterpret_cast<float2*>(smem)[j] = __ldg(&g_tw[((k * lo) % L) * (kTwN / L)]);   // lane-divergent: not the constant bank
        }
    }
#endif
    template <int Ls, int STR> static LCT_DEV float2 stage_mul(float2 a, int k, int lo) {
        static_assert(Ls == L && STR == STR0, "TwLine only serves the first stage");
        return cmul(a, at(k, lo));
    }
    template <int Ls, int STR> static LCT_DEV float2 stage_mulc(float2 a, int k, int lo) {
        static_assert(Ls == L && STR == STR0, "TwLine only serves the first stage");
        return cmulc(a, at(k, lo));
    }
};

// Line-thread index shared by a whole warp (column tiles are multiples of 32 wide): broadcasting it
// through a warp reduction lands it in a uniform register, so table lookups indexed by it can be
// uniform loads.
#ifdef LCT_EMULATE
static inline int warp_uniform(int v) { return v; }
#else
LCT_DEV int warp_uniform(int v) { return __reduce_max_sync(0xffffffffu, v); }
#endif
template <int LANES> LCT_DEV int line_thread(int tid) {
    if constexpr (LANES % 32 == 0) return warp_uniform(tid / LANES);
    else return tid / LANES;
}

#ifdef LCT_EMULATE
static inline int float_bits(float f) { int i; std::memcpy(&i, &f, 4); return i; }
#else
LCT_DEV int float_bits(float f) { return __float_as_int(f); }
#endif

// The parity-split transforms multiply element n of a line by w_2N^n (forward) or its conjugate (inverse).  A stage-0
// thread owns n = t + q * st with st = N / R0, so w_2N^n = w_2N^t * w_(2 R0)^q: one table value per thread (`wt`, read once)
// times a compile-time constant -- instead of one lane-divergent table load per element (those loads were a quarter of
// the memory instructions of the N = 256 row kernels, whose first stall is the memory-instruction queue).
#ifndef LCT_SPLIT_TW_CONST
#define LCT_SPLIT_TW_CONST 1
#endif
template <int R0, bool INV> LCT_DEV float2 split_twiddle(float2 a, float2 wt, int q) {
    static_assert(32 % (2 * R0) == 0, "w_(2 R0)^q must be a 32nd root of unity");
    const float2 b = INV ? cmulc(a, wt) : cmul(a, wt);
    return twmul32<INV>(b, q * (32 / (2 * R0)));
}

struct Params {
    int M, N, C, D;
    // K1 input placement: rows [in_be, in_be+in_T) of the M-bin time axis come from `in`
    // (C, in_T, N*N); in_be is per batch sample when in_be_dev != nullptr.
    // K5 output placement: rows [out_be, out_be+out_T) are written to `out` (C, out_T, N*N).
    int in_T, out_T;
    int be_uniform;
    const int* be_dev;          // B entries or nullptr
    int c_base;                 // global index of this launch's channel 0 (for be_dev lookups)
    const float* in;
    float* out;
    float2* s1;                 // (C, M+1, N, N)
    float2* s2;                 // (C, M+1, 2N, N)
    const float2* filt;         // (M+1, 2N, 2N), already scaled by 1/(8 M N N); or its quarter (filt_sym)
    int conj_filter;
    // filt_sym: only part of the filter is stored.  The light-cone PSF is mirror-symmetric in y and in x about a
    // half-sample centre (psf[-1-x] = psf[x] after the roll of helper.py:115-116), so along either axis
    // W(k) = w_2N^k W(2N - k) for k > N (and W(N) = 0): the mirrored part is the stored one times a twiddle.
    //   kFilterQuarter  (M+1, N+1, N+1): kh, kw <= N -- for the kernels that keep a filter row in registers across
    //                   the channel loop, where the per-value twiddle is paid once per block;
    //   kFilterHalfRows (M+1, N+1, 2N):  kh <= N     -- for the 512-point kernel, which re-reads the filter per
    //                   channel: one row-constant twiddle, no per-value table lookup.
    // lct_plan_create verifies the symmetry before choosing either layout.
    int filt_sym;
    // Resampling operator used by this launch (mtx rows for K1, mtxi rows for K5): one 16-byte
    // record per row {start * kEllStride, w0, w1, w2} (lct_tables.h); rows longer than 3
    // continue in vals[rowptr[row] + 3 ...].
    const float4* ell;
    const int* rowptr;
    const float* vals;
    const float4* pair;         // K1 only: pair records (two float4 per pair, lct_tables.h); pairs below kLongPairs unused
    // optional (K5, forward): per-channel {min key, complemented max key} of the volume it writes, reduced
    // with atomicMin while the values are still in registers (lct_normalize.cuh); pre-set to all ones
    unsigned long long* minmax_keys;
    // time kernels: how many blocks ahead (in launch order) the tile to warm in L2 lies; 0 = no prefetch.
    // The launcher sets it to the number of resident blocks, so the lines arrive about one block life early.
    int ahead;
    // programmatic dependent launch: 0 = plain stream order, 1 = every block releases the next kernel of the stream as
    // soon as it starts (that kernel's blocks then take the SMs the tail of this one leaves idle, run their prologue
    // and sleep in griddepcontrol.wait until this grid has finished and flushed).  Releasing only after the last phase
    // measured the same or slower.
    int pdl;
};

LCT_DEV void prefetch_l2(const void* ptr) {
#ifndef LCT_EMULATE
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

// 16-byte asynchronous global -> shared copy (LDGSTS); `valid` false zero-fills the 16 bytes instead.
// The emulator copies at issue time, one legal outcome of the asynchronous copy.
LCT_DEV void cp_async16(void* dst_smem, const void* src, bool valid) {
#ifndef LCT_EMULATE
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
#else
    if (valid) std::memcpy(dst_smem, src, 16); else std::memset(dst_smem, 0, 16);
#endif
}
LCT_DEV void cp_async_commit() {
#ifndef LCT_EMULATE
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
LCT_DEV void cp_async_wait_all() {
#ifndef LCT_EMULATE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// Bulk asynchronous copy (the TMA unit's one-dimensional form, cp.async.bulk) + mbarrier: one thread asks for a whole
// contiguous run of bytes; they land in shared memory without passing through any thread's registers or issue slots
// and complete the barrier's transaction count.  `bulk_load` both announces the bytes on the barrier and starts the
// copy, so every issuing thread accounts for exactly what it asked for.  Source, destination and size must be
// multiples of 16 bytes.  The emulator copies at issue time (one legal outcome) and its barriers are no-ops.
LCT_DEV void mbar_init(unsigned long long* bar, unsigned count) {
#ifndef LCT_EMULATE
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
    (void)bar; (void)count;
#endif
}
// one arrival that also announces `bytes` of pending bulk copies (the barrier's phase completes when every expected
// arrival has happened and every announced byte has landed)
LCT_DEV void mbar_arrive_expect(unsigned long long* bar, unsigned bytes) {
#ifndef LCT_EMULATE
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
#else
    (void)bar; (void)bytes;
#endif
}
LCT_DEV void bulk_load(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar) {
#ifndef LCT_EMULATE
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
#else
    (void)bar;
    std::memcpy(dst_smem, src, bytes);
#endif
}
LCT_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
#ifndef LCT_EMULATE
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LCT_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LCT_MBAR_DONE_%=;\n\t"
        "bra LCT_MBAR_WAIT_%=;\n\t"
        "LCT_MBAR_DONE_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
#else
    (void)bar; (void)parity;
#endif
}
// orders this thread's (and, after a barrier, the block's) earlier generic accesses to shared memory before later
// accesses of the asynchronous proxy: needed before a bulk copy overwrites a buffer the threads have just used
LCT_DEV void fence_proxy_async() {
#ifndef LCT_EMULATE
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// Persistent tile walk of the time kernels: min(tiles, resident blocks) blocks, block b walks tiles b, b + G, ...
// and copies tile i + 1 into shared memory (cp.async) while it still works on tile i.  K5 always runs this way;
// K1 only at one block per SM (M = 512) -- at two or more blocks per SM the plain loads with the L2 warm-up of a
// later block's tile measured faster (cfg2 38.0 vs 40.7 us, cfg4 91 vs 92 us, cfg3 280 vs 253 us).
// N is a power of two, so is the number of tiles per channel: split a tile index with a shift and a mask
LCT_DEV int ilog2_pow2(int v) {
#ifndef LCT_EMULATE
    return 31 - __clz(v);
#else
    return __builtin_ctz((unsigned)v);
#endif
}

struct TileWalk {
    int total, G;
    LCT_HD TileWalk(const Params& p, int tiles_per_channel, bool persist = true) {
        total = tiles_per_channel * p.C;
        G = (persist && p.ahead > 0 && p.ahead < total) ? p.ahead : total;
    }
    LCT_HD int iterations() const { return (total + G - 1) / G; }
};

#ifdef LCT_EMULATE
static const float kInfinity = std::numeric_limits<float>::infinity();
#else
#define kInfinity __int_as_float(0x7f800000)
#endif
// order-preserving key of a float and its position (see lct_normalize.cuh)
LCT_DEV unsigned long long minmax_key(float v, unsigned int pos, bool is_max) {
    const unsigned int b = (unsigned int)float_bits(v);
    unsigned int k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    if (v != v) k = is_max ? 0xffffffffu : 0u;         // NaN wins both reductions (torch.min / torch.max propagate it)
    return is_max ? ~(((unsigned long long)k << 32) | (0xffffffffu - pos)) : (((unsigned long long)k << 32) | pos);
}
// Block-level commit of the per-thread keys: warp shuffle reduction, then one shared-memory atomicMin pair per warp
// into the block's slot; the warp that arrives last (a shared counter tells) carries the slot to the channel's global
// keys and re-arms it.  One global atomic pair per block and tile instead of one per warp: at 16 x 128^3 the 16 warps
// x 512 tiles hammering one address per channel cost 36 us per step.  `slot` = {min key, max key, arrivals}, 24 bytes
// of shared memory initialised to {~0, ~0, 0}; every warp of the block must call (warp-uniformly).
struct MinMaxSlot { unsigned long long kmin, kmax; unsigned int arrived, pad; };
LCT_DEV void minmax_slot_init(MinMaxSlot* slot) { slot->kmin = ~0ull; slot->kmax = ~0ull; slot->arrived = 0u; }
LCT_DEV void minmax_commit(unsigned long long* keys, int c, unsigned long long kmin, unsigned long long kmax,
                           MinMaxSlot* slot, int nwarps) {
#ifdef LCT_EMULATE
    (void)slot; (void)nwarps;
    if (kmin < keys[2 * c]) keys[2 * c] = kmin;
    if (kmax < keys[2 * c + 1]) keys[2 * c + 1] = kmax;
#else
    LCT_UNROLL
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b < kmax ? b : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&slot->kmin, kmin);
        atomicMin(&slot->kmax, kmax);
        __threadfence_block();
        if (atomicAdd(&slot->arrived, 1u) == (unsigned int)nwarps - 1u) {      // last warp of the block for this tile
            __threadfence_block();
            atomicMin(keys + 2 * c, atomicExch(&slot->kmin, ~0ull));
            atomicMin(keys + 2 * c + 1, atomicExch(&slot->kmax, ~0ull));
            slot->arrived = 0u;               // nobody touches the slot again before the next tile's barriers
        }
    }
#endif
}

// sum_e w[e] * src[(start + e) * kEllStride] for one operator row; `src` must have two readable
// (finite) rows past the last one, because short rows still touch three.  The record carries
// start * kEllStride; rows longer than three continue in the CSR arrays, and the host guarantees
// (build_tables) that such rows exist only where the caller passes kTail = true.
template <bool kTail, int STRIDE = kEllStride>
LCT_DEV float band_dot(const Params& p, const float4* ell, int row, const float* src) {
    const float4 e = ell[row];                     // block-local copy in shared memory (warp-uniform: broadcast)
    // the low bits of the offset hold the number of entries past the third; they are zero wherever kTail is false
    const int off = float_bits(e.x), extra = kTail ? (off & (STRIDE - 1)) : 0;
    const float* s = src + (off - extra);
    float acc = e.y * s[0];
    acc = fmaf(e.z, s[STRIDE], acc);
    acc = fmaf(e.w, s[2 * STRIDE], acc);
    if constexpr (kTail) {
        if (extra) {                               // rare (the first ~sqrt(M)/6 rows of mtx): the CSR arrays hold the rest
            const int first = LCT_LDG(p.rowptr + row);
            const int len = extra < STRIDE - 1 ? 3 + extra : LCT_LDG(p.rowptr + row + 1) - first;      // the count saturates
            const float* v = p.vals + first;
            for (int k = 3; k < len; ++k) acc = fmaf(LCT_LDG(v + k), s[k * STRIDE], acc);
        }
    }
    return acc;
}

// Rows (2 pair, 2 pair + 1) of the operator applied together through one pair record (lct_tables.h): three tile
// loads, six multiply-adds.  Only valid for pairs >= kLongPairs (build_tables checks the operator).
template <int STRIDE = kEllStride>
LCT_DEV float2 pair_dot(const float4* pairs, int pair, const float* src) {
    const float4 a = pairs[2 * pair], b = pairs[2 * pair + 1];       // {offset, a0, a1, a2}, {b0, b1, b2, -}
    const float* s = src + float_bits(a.x);
    const float x0 = s[0], x1 = s[STRIDE], x2 = s[2 * STRIDE];
    return make_float2(fmaf(a.w, x2, fmaf(a.z, x1, a.y * x0)), fmaf(b.z, x2, fmaf(b.y, x1, b.x * x0)));
}

// Row j of mtxi = mtx^T (the inverse resampling): its band starts at row floor(j^2 / M) of the volume tile -- the closed
// form of helper.py:35-69's staircase, checked by build_tables -- so the tile address does not wait on the record load:
// the three tile loads and the record load issue together (the record -> address -> load chain was the first stall of
// the gather at 16 warps per SM).  LOGM = log2(M).
#ifndef LCT_MTXI_CLOSED_FORM
#define LCT_MTXI_CLOSED_FORM 1
#endif
template <int STRIDE, int LOGM>
LCT_DEV float band_dot_sq(const float4* ell, int j, const float* src) {
    const float4 e = ell[j];
    const float* s = src + ((j * j) >> LOGM) * STRIDE;
    return fmaf(e.w, s[2 * STRIDE], fmaf(e.z, s[STRIDE], e.y * s[0]));
}

LCT_DEV int window_begin(const Params& p, int c) {
    return p.be_dev ? LCT_LDG(p.be_dev + (p.c_base + c) / p.D) : p.be_uniform;
}

// Stage-0 butterfly inputs q < Q of the time-forward kernel cover the pairs whose two rows may span more than three
// columns together (pairs up to ~M/10 of helper.py:35-69's operator: 1 / 4 / 10 / 24 / 53 at M = 32 ... 512); the
// rest take one pair record each.  build_tables refuses an operator that does not fit.
// (Measured: the pair records take K1 from 246 to 226 us at 8 x 512x128x128 and from 94 to 90 us at 16 x 128^3; at
//  M = 256, where the kernel's time is fixed ramp-up and tail rather than instruction issue, they change nothing.
//  Keeping each thread's pair offsets in registers across the tile walk of the persistent kernel: no gain either.)
template <int M> struct TimeLongQ { static constexpr int Q = (M >= 512) ? 4 : ((M >= 128) ? 2 : 1); };

// ---------------------------------------------------------------------------
// K1: time window + falloff + resample + real FFT (2M, M non-zero) along T.
// Packed-real algorithm: z[n] = u[2n] + i u[2n+1] (n < M/2, zero above), Z = FFT_M(z),
// X[k] = Ev + w^k Od with Ev = (Z[k] + conj Z[M-k])/2, Od = -i (Z[k] - conj Z[M-k])/2.
// ---------------------------------------------------------------------------
template <class P, int CT_> struct TimeFwd {
    using TwS = TwNone;                        // constant-bank twiddles (a block-local table measured slower here)
    static constexpr int M = P::L, CT = CT_, kThreads = P::TL * CT;
    // phases: load x tile | gather + stage 0 -> zs | stages 1.. in place | post-process
    static constexpr int kPhases = 2 + (P::S - 1) + 1;
    // x tile, FFT buffer and the operator's row records side by side: the register file already caps
    // the kernel at two blocks per SM, so nothing is gained by aliasing them (and a phase is saved)
    static constexpr size_t kXs = ((size_t)(M + 2) * CT * sizeof(float) + 15) / 16 * 16;
    static constexpr size_t kWork = kXs + (size_t)M * CT * sizeof(float2);
    // operator tables behind the work buffers: the pair records (M/2 x 32 B) and the row records of the first
    // kLongRows rows, the only ones still applied row by row
    static constexpr int kLongQ = TimeLongQ<M>::Q, kLongPairs = kLongQ * P::st(0), kLongRows = 2 * kLongPairs;
    static_assert(P::TL == P::st(0), "stage-0 butterfly q of line thread tau starts at pair tau + q * st(0)");
    static constexpr size_t kSmem = TwS::kBytes + (kWork + (size_t)(M + kLongRows) * sizeof(float4));
    static constexpr bool kWarpSync = false;
    // 1024 threads/SM at <= 64 regs; the 32-wide butterflies need 128 regs (512 threads/SM)
    static constexpr int kMinBlocks = (P::E >= 32) ? (512 / kThreads > 0 ? 512 / kThreads : 1) : ((kThreads >= 1024) ? 1 : (kThreads == 256 ? LCT_K1_BLOCKS_256 : 1024 / kThreads));
    struct Regs {};
    static void grid(const Params& p, int& gx, int& gy) { gx = p.N * p.N / CT; gy = p.C; }
    static int iterations(const Params&) { return 1; }
    // One pass per block, but the loop stays: without it (kSinglePass) this kernel compiles to fewer registers and runs
    // slower (M = 128: 98 vs 90 us at 16 x 128^3; M = 256: no change).

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int col = tid % CT, tau = line_thread<CT>(tid);
        const int NN = p.N * p.N, col0 = bx * CT, c = by;
        float* xs = reinterpret_cast<float*>(smem);
        float2* zs = reinterpret_cast<float2*>(smem + kXs);
        if constexpr (PH == 0) {
            // x tile -> xs[(M+2)][CT] f32, zero outside the window [be, en) and in the two pad rows
            const int be = window_begin(p, c), en = be + p.in_T;
            constexpr int V4 = CT / 4, kSlots = (M + 2) * V4;
            const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)c * p.in_T * NN + col0);
            float4* xs4 = reinterpret_cast<float4*>(smem);
            for (int j = tid; j < M + kLongRows; j += kThreads)
                reinterpret_cast<float4*>(smem + kWork)[j] = j < M ? LCT_LDG(p.pair + j) : LCT_LDG(p.ell + (j - M));
            // slot i = tid + u * kThreads covers row t = i / V4, column quad q = i % V4: q is fixed per
            // thread and t advances by kThreads / V4 per slot, so the source offset is stepped, not recomputed
            static_assert(kThreads % V4 == 0, "column quad must be fixed per thread");
            constexpr int kRowStep = kThreads / V4, kIters = (kSlots + kThreads - 1) / kThreads;
            const int t0 = tid / V4;
            ptrdiff_t off = (ptrdiff_t)(t0 - be) * (NN / 4) + tid % V4;
            const ptrdiff_t off_step = (ptrdiff_t)kRowStep * (NN / 4);
            // M = 256: the tile goes straight into shared memory by 16-byte asynchronous copies (no staging registers, no
            // stores: K1 39.0 -> 36.9 us at 8 x 256x64^2, 115 -> 108 at 32); at M = 128 and M = 64 the loads through
            // registers are faster (90 vs 96 us at 16 x 128^3, 14.5 vs 16.4 at 8 x 64x64^2)
            constexpr bool kAsyncTile = LCT_K1_ASYNC_TILE_256 && M == 256;
            [[maybe_unused]] float4 v[kIters];
            LCT_UNROLL
            for (int u = 0; u < kIters; ++u, off += off_step) {
                const int t = t0 + u * kRowStep;
                const bool ok = t >= be && t < en;
                if constexpr (kAsyncTile) {                  // zero-filled outside the window
                    if (tid + u * kThreads < kSlots) cp_async16(xs4 + tid + u * kThreads, ok ? src + off : src, ok);
                } else {
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) v[u] = LCT_LDG(src + off);
                }
            }
            if constexpr (kAsyncTile) cp_async_commit();
            if (p.ahead > 0) {
                // own loads are in flight: ask L2 for the tile of the block that will run here about one block life later
                const int gx = NN / CT;
                const long long next = (long long)by * gx + bx + p.ahead;
                const int nc = (int)(next / gx), nb = (int)(next % gx);     // (shift/mask here: fewer instructions, yet K1 measured 2-10 % slower)
                if (nc < p.C) {
                    const float* ns = p.in + (size_t)nc * p.in_T * NN + (size_t)nb * CT;     // one 128-byte line per time bin
                    for (int t = tid; t < p.in_T; t += kThreads) prefetch_l2(ns + (size_t)t * NN);
                }
            }
            if constexpr (kAsyncTile) cp_async_wait_all();
            else {
                LCT_UNROLL
                for (int u = 0; u < kIters; ++u) {
                    const int i = tid + u * kThreads;
                    if (i < kSlots) xs4[i] = v[u];
                }
            }
        } else if constexpr (PH == 1) {
            fwd_stage<P, 0, true, TwS>(tau,
                [&](int pos, int slot) {
                    // input q of the butterfly is pair tau + q * st(0): from q = kLongQ on, one pair record; below, row
                    // by row -- and rows longer than three taps sit below 2 * st(0): the first input of a butterfly only
                    const float4* pairs = reinterpret_cast<const float4*>(smem + kWork);
                    const float4* ell = pairs + M;
                    if (slot % P::radix(0) >= kLongQ) return pair_dot<CT>(pairs, pos, xs + col);
                    if (slot % P::radix(0) == 0)
                        return make_float2(band_dot<true, CT>(p, ell, 2 * pos, xs + col), band_dot<true, CT>(p, ell, 2 * pos + 1, xs + col));
                    return make_float2(band_dot<false, CT>(p, ell, 2 * pos, xs + col), band_dot<false, CT>(p, ell, 2 * pos + 1, xs + col));
                },
                [&](int pos, int, float2 v) { zs[pos * CT + col] = v; });
        } else if constexpr (PH < 1 + P::S) {
            constexpr int s = PH - 1;
            fwd_stage<P, s, false, TwS>(tau,
                [&](int pos, int) { return zs[pos * CT + col]; },
                [&](int pos, int, float2 v) { zs[pos * CT + col] = v; });
        } else {
            // X[k] = Ev + w^k Od and X[M-k] = conj(Ev - w^k Od) from the pair (Z[k], Z[M-k])
            float2* dst = p.s1 + (size_t)c * (M + 1) * NN + col0 + col;
            const float2* zc = zs + col;
            auto emit = [&](int k, int pos_k, int pos_mk, float2* lo, float2* hi) {
                const float2 zk = zc[pos_k * CT];
                const float2 zm = cconj(zc[pos_mk * CT]);
                const float2 ev = cscale(cadd(zk, zm), 0.5f);
                const float2 d = csub(zk, zm);
                const float2 od = make_float2(0.5f * d.y, -0.5f * d.x);
                const float2 t = TwS::mul(od, k * (kTwN / (2 * M)));
                *lo = cadd(ev, t);
                if (hi != lo) *hi = cconj(csub(ev, t));       // (a compile-time flag instead of the compare: 6 % slower at M = 128)
            };
            constexpr int kPairs = (M / 2) / P::TL, kBlocks = M / P::TL;        // k = tau + m*TL < M/2
            float2* lo = dst + (size_t)tau * NN;
            float2* hi = dst + (size_t)(M - tau) * NN;
            const size_t step = (size_t)P::TL * NN;
            const int pos_tau = P::freq_to_pos(tau);
            // position of frequency M - k: for tau > 0, M - k = (kBlocks - 1 - m) TL + (TL - tau), two disjoint
            // bit fields of the digit reversal (a bit permutation); for tau == 0 it is (kBlocks - m) TL mod M
            const int pos_neg = P::freq_to_pos((P::TL - tau) & (P::TL - 1));
            LCT_UNROLL
            for (int m = 0; m < kPairs; ++m) {
                const int pos_mk = tau != 0 ? pos_neg + P::freq_to_pos((kBlocks - 1 - m) * P::TL)
                                            : P::freq_to_pos(((kBlocks - m) * P::TL) & (M - 1));
                emit(tau + m * P::TL, pos_tau + P::freq_to_pos(m * P::TL), pos_mk, lo, hi);   // disjoint bit fields
                lo += step;
                hi -= step;
            }
            if (tau == 0) emit(M / 2, P::freq_to_pos(M / 2), P::freq_to_pos(M / 2), dst + (size_t)(M / 2) * NN, dst + (size_t)(M / 2) * NN);
        }
    }
};

// K1, persistent form (used at one block per SM, M = 512): same phases as TimeFwd, but min(tiles, resident)
// blocks walk the tiles and the next tile's rows are copied into xs (cp.async) during the in-place stages.
// Kept as a separate type so that the one-tile-per-block kernel above compiles exactly as before: at
// M <= 256 its code generation is sensitive enough that sharing the source cost it up to 10 %.
template <class P, int CT_> struct TimeFwdPersistent {
    using TwS = TwNone;                        // constant-bank twiddles (a block-local table measured slower here)
    static constexpr int M = P::L, CT = CT_, kThreads = P::TL * CT;
    // phases: load x tile | gather + stage 0 -> zs | stages 1.. in place | post-process
    static constexpr int kPhases = 2 + (P::S - 1) + 1;
    // x tile, FFT buffer and the operator's row records side by side: the register file already caps
    // the kernel at two blocks per SM, so nothing is gained by aliasing them (and a phase is saved)
    static constexpr size_t kXs = ((size_t)(M + 2) * CT * sizeof(float) + 15) / 16 * 16;
    static constexpr size_t kWork = kXs + (size_t)M * CT * sizeof(float2);
    // operator tables behind the work buffers: the pair records (M/2 x 32 B) and the row records of the first
    // kLongRows rows, the only ones still applied row by row
    static constexpr int kLongQ = TimeLongQ<M>::Q, kLongPairs = kLongQ * P::st(0), kLongRows = 2 * kLongPairs;
    static_assert(P::TL == P::st(0), "stage-0 butterfly q of line thread tau starts at pair tau + q * st(0)");
    static constexpr size_t kSmem = TwS::kBytes + (kWork + (size_t)(M + kLongRows) * sizeof(float4));
    static constexpr bool kWarpSync = false;
    // 1024 threads/SM at <= 64 regs; the 32-wide butterflies need 128 regs (512 threads/SM)
    static constexpr int kMinBlocks = (P::E >= 32) ? (512 / kThreads > 0 ? 512 / kThreads : 1) : ((kThreads >= 1024) ? 1 : 1024 / kThreads);
    struct Regs {};
    static constexpr bool kPersist = true;
    static void grid(const Params& p, int& gx, int& gy) {
        if (kPersist) { gx = TileWalk(p, p.N * p.N / CT).G; gy = 1; }
        else { gx = p.N * p.N / CT; gy = p.C; }              // one tile per block: (column tile, channel)
    }
    static int iterations(const Params& p) { return kPersist ? TileWalk(p, p.N * p.N / CT).iterations() : 1; }
    static constexpr bool kTileWalks = true;              // see TimeInv::walk_active
    static LCT_DEV bool walk_active(const Params& p, int bx, int, int it) {
        const TileWalk walk(p, p.N * p.N / CT, kPersist);
        return !kPersist || bx + it * walk.G < walk.total;
    }

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    // the operator tables and the block's first tile are requested from the kernel driver, right after the wait for the
    // previous kernel and before the prologue's barrier: the first phase is then the same code for every tile of the walk
    // (K1 229 -> 221 us at cfg3, 122 -> 119 at cfg5)
    static constexpr bool kPreload = LCT_TIME_PRELOAD && kPersist;
    static LCT_DEV void preload(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int) {
        if constexpr (kPreload) {
            unsigned char* smem = smem_base + TwS::kBytes;
            for (int j = tid; j < M + kLongRows; j += kThreads)
                reinterpret_cast<float4*>(smem + kWork)[j] = j < M ? LCT_LDG(p.pair + j) : LCT_LDG(p.ell + (j - M));
            issue_tile(p, smem, tid, bx);
        }
    }

    // x tile -> xs[(M+2)][CT] f32 by 16-byte asynchronous copies, zero-filled outside the window [be, en) and in
    // the two pad rows.  Slot i = tid + u * kThreads covers row i / V4, column quad i % V4: the quad is fixed per
    // thread and the row advances by kThreads / V4 per slot, so the source offset is stepped, not recomputed.
    static LCT_DEV void issue_tile(const Params& p, unsigned char* smem, int tid, int tile) {
        constexpr int V4 = CT / 4, kSlots = (M + 2) * V4;
        static_assert(kThreads % V4 == 0, "column quad must be fixed per thread");
        constexpr int kRowStep = kThreads / V4, kIters = (kSlots + kThreads - 1) / kThreads;
        const int NN = p.N * p.N, tpc = NN / CT, c = tile >> ilog2_pow2(tpc), col0 = (tile & (tpc - 1)) * CT;
        const int be = window_begin(p, c), en = be + p.in_T;
        const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)c * p.in_T * NN + col0);
        float4* xs4 = reinterpret_cast<float4*>(smem);
        const int t0 = tid / V4;
        ptrdiff_t off = (ptrdiff_t)(t0 - be) * (NN / 4) + tid % V4;
        const ptrdiff_t off_step = (ptrdiff_t)kRowStep * (NN / 4);
        LCT_UNROLL
        for (int u = 0; u < kIters; ++u, off += off_step) {
            const int i = tid + u * kThreads, t = t0 + u * kRowStep;
            const bool ok = t >= be && t < en;
            if (i < kSlots) cp_async16(xs4 + i, ok ? src + off : src, ok);
        }
        cp_async_commit();
    }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int it) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int col = tid % CT, tau = line_thread<CT>(tid);
        const int NN = p.N * p.N, tpc = NN / CT;
        const TileWalk walk(p, tpc, kPersist);
        int tile, col0, c;
        if constexpr (kPersist) {
            tile = bx + it * walk.G;
#ifdef LCT_EMULATE
            if (tile >= walk.total) return;                  // on the GPU the driver leaves the loop (walk_active)
#endif
            col0 = (tile & (tpc - 1)) * CT;
            c = tile >> ilog2_pow2(tpc);
        } else {
            tile = by * tpc + bx; col0 = bx * CT; c = by;
        }
        float* xs = reinterpret_cast<float*>(smem);
        float2* zs = reinterpret_cast<float2*>(smem + kXs);
        if constexpr (PH == 0 && kPersist) {
            if (!kPreload && it == 0) {
                for (int j = tid; j < M + kLongRows; j += kThreads)
                    reinterpret_cast<float4*>(smem + kWork)[j] = j < M ? LCT_LDG(p.pair + j) : LCT_LDG(p.ell + (j - M));
                issue_tile(p, smem, tid, tile);
            }
            cp_async_wait_all();                             // later tiles were issued during the previous tile's stages
        } else if constexpr (PH == 0) {
            // one tile per block: plain 128-bit loads through registers, same slot walk as issue_tile
            const int be = window_begin(p, c), en = be + p.in_T;
            constexpr int V4 = CT / 4, kSlots = (M + 2) * V4;
            const float4* src = reinterpret_cast<const float4*>(p.in + (size_t)c * p.in_T * NN + col0);
            float4* xs4 = reinterpret_cast<float4*>(smem);
            for (int j = tid; j < M + kLongRows; j += kThreads)
                reinterpret_cast<float4*>(smem + kWork)[j] = j < M ? LCT_LDG(p.pair + j) : LCT_LDG(p.ell + (j - M));
            constexpr int kRowStep = kThreads / V4, kIters = (kSlots + kThreads - 1) / kThreads;
            const int t0 = tid / V4;
            ptrdiff_t off = (ptrdiff_t)(t0 - be) * (NN / 4) + tid % V4;
            const ptrdiff_t off_step = (ptrdiff_t)kRowStep * (NN / 4);
            float4 v[kIters];
            LCT_UNROLL
            for (int u = 0; u < kIters; ++u, off += off_step) {
                const int t = t0 + u * kRowStep;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t >= be && t < en) v[u] = LCT_LDG(src + off);
            }
            if (p.ahead > 0 && tile + p.ahead < walk.total) {
                // own loads are in flight: ask L2 for the tile of the block that will run here about one block life later
                const int nt = tile + p.ahead;
                const float* ns = p.in + (size_t)(nt >> ilog2_pow2(tpc)) * p.in_T * NN + (size_t)(nt & (tpc - 1)) * CT;     // one 128-byte line per time bin
                for (int t = tid; t < p.in_T; t += kThreads) prefetch_l2(ns + (size_t)t * NN);
            }
            LCT_UNROLL
            for (int u = 0; u < kIters; ++u) {
                const int i = tid + u * kThreads;
                if (i < kSlots) xs4[i] = v[u];
            }
        } else if constexpr (PH == 1) {
            fwd_stage<P, 0, true, TwS>(tau,
                [&](int pos, int slot) {
                    // input q of the butterfly is pair tau + q * st(0): from q = kLongQ on, one pair record; below, row
                    // by row -- and rows longer than three taps sit below 2 * st(0): the first input of a butterfly only
                    const float4* pairs = reinterpret_cast<const float4*>(smem + kWork);
                    const float4* ell = pairs + M;
                    if (slot % P::radix(0) >= kLongQ) return pair_dot<CT>(pairs, pos, xs + col);
                    if (slot % P::radix(0) == 0)
                        return make_float2(band_dot<true, CT>(p, ell, 2 * pos, xs + col), band_dot<true, CT>(p, ell, 2 * pos + 1, xs + col));
                    return make_float2(band_dot<false, CT>(p, ell, 2 * pos, xs + col), band_dot<false, CT>(p, ell, 2 * pos + 1, xs + col));
                },
                [&](int pos, int, float2 v) { zs[pos * CT + col] = v; });
        } else if constexpr (PH < 1 + P::S) {
            constexpr int s = PH - 1;
            if constexpr (s == 1 && kPersist) {              // xs was consumed by the gather: start the next tile's copy
                if (tile + walk.G < walk.total) issue_tile(p, smem, tid, tile + walk.G);
            }
            fwd_stage<P, s, false, TwS>(tau,
                [&](int pos, int) { return zs[pos * CT + col]; },
                [&](int pos, int, float2 v) { zs[pos * CT + col] = v; });
        } else {
            // X[k] = Ev + w^k Od and X[M-k] = conj(Ev - w^k Od) from the pair (Z[k], Z[M-k])
            float2* dst = p.s1 + (size_t)c * (M + 1) * NN + col0 + col;
            const float2* zc = zs + col;
            auto emit = [&](int k, int pos_k, int pos_mk, float2* lo, float2* hi) {
                const float2 zk = zc[pos_k * CT];
                const float2 zm = cconj(zc[pos_mk * CT]);
                const float2 ev = cscale(cadd(zk, zm), 0.5f);
                const float2 d = csub(zk, zm);
                const float2 od = make_float2(0.5f * d.y, -0.5f * d.x);
                const float2 t = TwS::mul(od, k * (kTwN / (2 * M)));
                *lo = cadd(ev, t);
                if (hi != lo) *hi = cconj(csub(ev, t));       // (a compile-time flag instead of the compare: 6 % slower at M = 128)
            };
            constexpr int kPairs = (M / 2) / P::TL, kBlocks = M / P::TL;        // k = tau + m*TL < M/2
            float2* lo = dst + (size_t)tau * NN;
            float2* hi = dst + (size_t)(M - tau) * NN;
            const size_t step = (size_t)P::TL * NN;
            const int pos_tau = P::freq_to_pos(tau);
            // position of frequency M - k: for tau > 0, M - k = (kBlocks - 1 - m) TL + (TL - tau), two disjoint
            // bit fields of the digit reversal (a bit permutation); for tau == 0 it is (kBlocks - m) TL mod M
            const int pos_neg = P::freq_to_pos((P::TL - tau) & (P::TL - 1));
            LCT_UNROLL
            for (int m = 0; m < kPairs; ++m) {
                const int pos_mk = tau != 0 ? pos_neg + P::freq_to_pos((kBlocks - 1 - m) * P::TL)
                                            : P::freq_to_pos(((kBlocks - m) * P::TL) & (M - 1));
                emit(tau + m * P::TL, pos_tau + P::freq_to_pos(m * P::TL), pos_mk, lo, hi);   // disjoint bit fields
                lo += step;
                hi -= step;
            }
            if (tau == 0) emit(M / 2, P::freq_to_pos(M / 2), P::freq_to_pos(M / 2), dst + (size_t)(M / 2) * NN, dst + (size_t)(M / 2) * NN);
        }
    }
};

// ---------------------------------------------------------------------------
// K5: Hermitian inverse FFT along T (keep t < M), real part, mtxi gather.
// ---------------------------------------------------------------------------
template <class P, int CT_> struct TimeInv {
    using TwS = TwNone;                        // constant-bank twiddles (a block-local table measured slower here)
    static constexpr int M = P::L, CT = CT_, kThreads = P::TL * CT;
    static constexpr int kPhases = 2 + P::S + 1;          // load | stage S-1 -> regs | scatter | middle.. | stage 0 -> vol | gather
    // spectrum tile / FFT buffer (aliased: the first stage goes through registers) + the real volume
    // tile next to it (registers cap the kernel at two blocks per SM anyway; saves a phase)
    static constexpr size_t kZs = ((size_t)(M + 1) * CT * sizeof(float2) + 15) / 16 * 16;
    static constexpr size_t kWork = kZs + (size_t)(M + 2) * CT * sizeof(float);
    static constexpr size_t kBarRel = kWork + (size_t)M * sizeof(float4) + 32;    // behind the row records and the MinMaxSlot
    static constexpr size_t kSmem = TwS::kBytes + kBarRel + 16;
    static constexpr bool kWarpSync = false;
    // 1024 threads/SM at <= 64 regs; the 32-wide butterflies need 128 regs (512 threads/SM)
    static constexpr int kMinBlocks = (P::E >= 32) ? (512 / kThreads > 0 ? 512 / kThreads : 1) : ((kThreads >= 1024) ? 1 : (kThreads == 256 ? LCT_K5_BLOCKS_256 : 1024 / kThreads));
    struct Regs { float2 a[P::E]; };
    static void grid(const Params& p, int& gx, int& gy) { gx = TileWalk(p, p.N * p.N / CT).G; gy = 1; }
    static int iterations(const Params& p) { return TileWalk(p, p.N * p.N / CT).iterations(); }
#ifndef LCT_TIME_INV_DRIVER_EXIT
#define LCT_TIME_INV_DRIVER_EXIT 0
#endif
    // Block-uniform: false once the block's walk has run past the last tile.  When the driver tests it (once per
    // iteration) the phases carry no early return, the compiler sees that nothing in Regs lives from one tile to the
    // next and the three 64-bit spills and reloads of r.a per tile disappear -- and the kernel runs 6-12 % slower
    // (221 -> 234 us at 8 x 512x128^2, 100 -> 112 at 16 x 128^3, 37.0 -> 39.6 at 8 x 256x64^2): the reload only ever
    // showed up in the profile because it is the first instruction behind the wait for the tile.  Off for this kernel.
    static constexpr bool kDriverExit = LCT_TIME_INV_DRIVER_EXIT;
    static LCT_DEV bool walk_active(const Params& p, int bx, int, int it) {
        const TileWalk walk(p, p.N * p.N / CT);
        return !kDriverExit || bx + it * walk.G < walk.total;
    }
    static constexpr bool kTileWalks = true;

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) {
        TwS::fill(smem, tid, kThreads);
        if (tid == 0) minmax_slot_init(reinterpret_cast<MinMaxSlot*>(smem + TwS::kBytes + kWork + (size_t)M * sizeof(float4)));
        if (kBulk && tid == 0) mbar_init(reinterpret_cast<unsigned long long*>(smem + TwS::kBytes + kBarRel), 1);
#ifndef LCT_EMULATE
        if constexpr (kBulk) __syncthreads();               // the driver skips the barrier behind a table-less prologue
#endif
    }

#ifndef LCT_TIME_INV_BULK
#define LCT_TIME_INV_BULK 0
#endif
    // Bulk mode: every row of the tile is CT * 8 contiguous bytes of S1, so thread t asks the bulk-copy unit for row t
    // (one instruction where the 16-byte copies below take (M + 1) * CT / 2 / kThreads, with their addresses) and the
    // rows complete one mbarrier; thread 0 makes the single arrival and announces the whole tile's bytes.
    static constexpr bool kBulk = LCT_TIME_INV_BULK && (CT * sizeof(float2)) % 16 == 0;
    static LCT_DEV void wait_tile(unsigned char* smem, int it) {
        if constexpr (kBulk) mbar_wait(reinterpret_cast<unsigned long long*>(smem + kBarRel), (unsigned)(it & 1));
        else cp_async_wait_all();
    }
    static LCT_DEV void issue_tile_bulk(const Params& p, unsigned char* smem, int tid, int tile) {
        constexpr unsigned kRowBytes = CT * sizeof(float2);
        const int NN = p.N * p.N, tpc = NN / CT, c = tile >> ilog2_pow2(tpc), col0 = (tile & (tpc - 1)) * CT;
        const float2* src = p.s1 + (size_t)c * (M + 1) * NN + col0;
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + kBarRel);
        fence_proxy_async();                      // the stages' accesses to zs (ordered by the barrier) before the copies
        if (tid == 0) mbar_arrive_expect(bar, (unsigned)(M + 1) * kRowBytes);
        for (int k = tid; k <= M; k += kThreads) bulk_load(smem + (size_t)k * kRowBytes, src + (size_t)k * NN, kRowBytes, bar);
    }

    // (Requesting the row records and the first tile from the kernel driver instead of in the first phase, as the
    //  persistent K1 does, measured slower here: 37.3 -> 39.4 us at cfg2, 216 -> 219 at cfg3.)
    // spectrum tile -> zs[(M+1)][CT] c64 by 16-byte asynchronous copies (two columns each)
    static LCT_DEV void issue_tile(const Params& p, unsigned char* smem, int tid, int tile) {
        if constexpr (kBulk) { issue_tile_bulk(p, smem, tid, tile); return; }
        constexpr int V2 = CT / 2, kSlots = (M + 1) * V2;
        static_assert(kThreads % V2 == 0, "column pair must be fixed per thread");
        constexpr int kIters = (kSlots + kThreads - 1) / kThreads;
        const int NN = p.N * p.N, tpc = NN / CT, c = tile >> ilog2_pow2(tpc), col0 = (tile & (tpc - 1)) * CT;
        const float4* src = reinterpret_cast<const float4*>(p.s1 + (size_t)c * (M + 1) * NN + col0);
        float4* zs4 = reinterpret_cast<float4*>(smem);
        size_t off = (size_t)(tid / V2) * (NN / 2) + tid % V2;           // stepped per slot, not recomputed
        const size_t off_step = (size_t)(kThreads / V2) * (NN / 2);
        LCT_UNROLL
        for (int u = 0; u < kIters; ++u, off += off_step)
            if (tid + u * kThreads < kSlots) cp_async16(zs4 + tid + u * kThreads, src + off, true);
        cp_async_commit();
    }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs& r, unsigned char* smem_base, int tid, int bx, int, int it) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int col = tid % CT, tau = line_thread<CT>(tid);
        const int NN = p.N * p.N, tpc = NN / CT;
        const TileWalk walk(p, tpc);
        const int tile = bx + it * walk.G;
#ifndef LCT_EMULATE
        if constexpr (!kDriverExit)
#endif
        if (tile >= walk.total) return;                      // block-uniform: the barriers are in the driver
        const int col0 = (tile & (tpc - 1)) * CT, c = tile >> ilog2_pow2(tpc);
        float2* zs = reinterpret_cast<float2*>(smem);
        float* vol = reinterpret_cast<float*>(smem + kZs);
        constexpr int SL = P::S - 1;
        if constexpr (PH == 0) {
            if (it == 0) {
                for (int j = tid; j < M; j += kThreads) reinterpret_cast<float4*>(smem + kWork)[j] = LCT_LDG(p.ell + j);
                issue_tile(p, smem, tid, tile);
            }
            wait_tile(smem, it);                             // later tiles were issued during the previous tile's gather
        } else if constexpr (PH == 1) {
            // Z[k] = (X[k] + conj X[M-k]) + i conj(w^k) (X[k] - conj X[M-k])
            auto z_at = [&](int k) -> float2 {
                const float2 xk = zs[k * CT + col];
                const float2 xm = cconj(zs[(M - k) * CT + col]);
                const float2 d = TwS::mulc(csub(xk, xm), k * (kTwN / (2 * M)));
                return cadd(cadd(xk, xm), make_float2(-d.y, d.x));
            };
            if constexpr (SL == 0) {
                inv_stage<P, 0, true, TwS>(tau,
                    [&](int pos, int slot) { return z_at(P::template freq_of<0>(pos, slot)); },
                    [&](int pos, int, float2 v) { vol[(2 * pos) * CT + col] = v.x; vol[(2 * pos + 1) * CT + col] = v.y; });
                if (tau == 0) { vol[M * CT + col] = 0.f; vol[(M + 1) * CT + col] = 0.f; }
            } else {
                inv_stage<P, SL, false, TwS>(tau,
                    [&](int pos, int slot) { return z_at(P::template freq_of<SL>(pos, slot)); },
                    [&](int, int slot, float2 v) { r.a[slot] = v; });
            }
        } else if constexpr (PH == 2) {
            if constexpr (SL > 0) {
                for_each_slot<P, SL>(tau, [&](int pos, int slot) { zs[pos * CT + col] = r.a[slot]; });
            }
        } else if constexpr (PH < 2 + SL) {          // middle stages SL-1 .. 1, in place
            constexpr int s = SL - (PH - 2);
            inv_stage<P, s, false, TwS>(tau,
                [&](int pos, int) { return zs[pos * CT + col]; },
                [&](int pos, int, float2 v) { zs[pos * CT + col] = v; });
        } else if constexpr (PH == 2 + SL && SL > 0) {
            inv_stage<P, 0, true, TwS>(tau,
                [&](int pos, int) { return zs[pos * CT + col]; },
                [&](int pos, int, float2 v) { vol[(2 * pos) * CT + col] = v.x; vol[(2 * pos + 1) * CT + col] = v.y; });
            if (tau == 0) { vol[M * CT + col] = 0.f; vol[(M + 1) * CT + col] = 0.f; }   // pad rows for band_dot
        } else if constexpr (PH == kPhases - 1) {
            // zs was consumed by the last stage: the next tile's spectrum flies in while this one is gathered
            if (SL > 0 && tile + walk.G < walk.total) issue_tile(p, smem, tid, tile + walk.G);
            const int be = window_begin(p, c);
            float* dst = p.out + (size_t)c * p.out_T * NN + col0 + col;
            float* d = dst + (size_t)tau * NN;
            const size_t step = (size_t)P::TL * NN;
            const float* vc = vol + col;
            const float4* ell = reinterpret_cast<const float4*>(smem + kWork);
#ifndef LCT_K5_UNROLL
#define LCT_K5_UNROLL 8
#endif
            [[maybe_unused]] constexpr int kGatherUnroll = LCT_K5_UNROLL;
#ifndef LCT_TIME_INV_TRIP_COUNT
#define LCT_TIME_INV_TRIP_COUNT 1
#endif
            // outputs tau, tau + TL, ... below out_T: a warp-uniform trip count instead of a test per output -- the
            // uniform branch around every output kept the compiler from batching the loads of one unrolled group,
            // so each output waited out its own shared-memory round trip (wait + short_scoreboard were 53 % of this
            // phase's stall samples, and the phase is 39 % of the kernel)
            const int rows_out = LCT_TIME_INV_TRIP_COUNT ? (p.out_T - tau + P::TL - 1) / P::TL : M / P::TL;
            if (p.minmax_keys == nullptr) {
                [[maybe_unused]] const float4* er = ell + be + tau;         // mtxi rows have at most three entries: no tail
#ifndef LCT_EMULATE
#pragma unroll kGatherUnroll                        // 8: full unrolling (16) measured 2 % slower at M = 256
#endif
                for (int m = 0; m < rows_out; ++m, d += step)
                    if (LCT_TIME_INV_TRIP_COUNT || tau + m * P::TL < p.out_T) {
                        if constexpr (LCT_MTXI_CLOSED_FORM) *d = band_dot_sq<CT, P::ilog2(M)>(ell, be + tau + m * P::TL, vc);
                        else *d = band_dot<false, CT>(p, er, m * P::TL, vc);
                    }
            } else {
                // Values only (two FMNMX per output): the positions the backward of normalize_feature needs are found by
                // the normalisation pass itself, which reads every value anyway (lct_normalize.cuh) -- the keys leave
                // here with the position field at "unknown", which any real position beats.  Tracking the positions in
                // this loop (two compares, two value selects, two index selects per output) cost 19 % of the kernel.
                // FMNMX drops a NaN, so every value is also folded into `poison` (v * 0 is 0 for finite v, NaN for a NaN
                // or an infinity: one FFMA per output); a poisoned strip redoes its reduction on the integer keys, where
                // NaN wins both sides as it does in torch.min / torch.max.
                float mn = kInfinity, mx = -kInfinity, poison = 0.f;
                [[maybe_unused]] const float4* er = ell + be + tau;         // same loop shape as the plain path above
#ifndef LCT_EMULATE
#pragma unroll kGatherUnroll
#endif
                for (int m = 0; m < rows_out; ++m, d += step)
                    if (LCT_TIME_INV_TRIP_COUNT || tau + m * P::TL < p.out_T) {
                        float v;
                        if constexpr (LCT_MTXI_CLOSED_FORM) v = band_dot_sq<CT, P::ilog2(M)>(ell, be + tau + m * P::TL, vc);
                        else v = band_dot<false, CT>(p, er, m * P::TL, vc);
                        *d = v;
                        poison = fmaf(v, 0.f, poison);
                        mn = fminf(mn, v);
                        mx = fmaxf(mx, v);
                    }
                const unsigned int base = (unsigned int)(col0 + col);
                unsigned long long kmin = minmax_key(mn, 0xffffffffu, false);
                unsigned long long kmax = minmax_key(mx, 0xffffffffu, true);
                if (poison != 0.f) {                       // NaN or infinity somewhere in this thread's strip
                    kmin = ~0ull; kmax = ~0ull;
                    for (int j = tau; j < p.out_T; j += P::TL) {
                        const float v = band_dot<false, CT>(p, ell, be + j, vc);
                        const unsigned long long a = minmax_key(v, (unsigned int)j * NN + base, false);
                        const unsigned long long b = minmax_key(v, (unsigned int)j * NN + base, true);
                        kmin = a < kmin ? a : kmin;
                        kmax = b < kmax ? b : kmax;
                    }
                }
                minmax_commit(p.minmax_keys, p.c_base + c, kmin, kmax,
                              reinterpret_cast<MinMaxSlot*>(smem + kWork + (size_t)M * sizeof(float4)), kThreads / 32);
            }
        }
    }
};

// ---------------------------------------------------------------------------
// K2: zero-extended forward FFT along H.  One block = one (c, kt) plane x CT columns.
// ---------------------------------------------------------------------------
template <class P, int CT_> struct RowFwd {
    using TwS = TwNone;                        // constant-bank twiddles (a block-local table measured slower here)
    static constexpr int L = P::L, N = L / 2, CT = CT_, kThreads = P::TL * CT;
    static constexpr int kPhases = P::S;
    static constexpr bool kWarpSync = false;
    static constexpr int kMinBlocks = (P::E >= 32) ? (512 / kThreads > 0 ? 512 / kThreads : 1)
                                                   : ((kThreads >= 1024) ? 1 : ((1024 / kThreads) > 32 ? 32 : (1024 / kThreads)));   // <= 64 regs
    static constexpr size_t kSmem = TwS::kBytes + ((P::S > 1) ? (size_t)L * CT * sizeof(float2) : 0);
    struct Regs {};
    static void grid(const Params& p, int& gx, int& gy) { gx = p.N / CT; gy = p.C * (p.M + 1); }
    static int iterations(const Params&) { return 1; }
    static constexpr bool kSinglePass = true;       // the driver runs the phases once, without a loop (K2 295 -> 293 us at 8 x 512x128^2)

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int col = tid % CT, tau = line_thread<CT>(tid);
        float2* zs = reinterpret_cast<float2*>(smem);
        const float2* src = p.s1 + (size_t)by * N * N + bx * CT + col;
        float2* dst = p.s2 + (size_t)by * L * N + bx * CT + col;
        auto ld_g = [&](int pos, int) { return LCT_LDCG(src + (size_t)pos * N); };        // (K2 156 -> 155 us at cfg4)
        auto ld_s = [&](int pos, int) { return zs[pos * CT + col]; };
        auto st_s = [&](int pos, int, float2 v) { zs[pos * CT + col] = v; };
        auto st_g = [&](int pos, int slot, float2 v) { dst[(size_t)P::template freq_of<P::S - 1>(pos, slot) * N] = v; };
        if constexpr (PH == 0) {
            if (p.ahead > 0) {       // warm L2 with the (plane, column tile) of the block one residency ahead
                const int gx = N / CT;
                const long long next = (long long)by * gx + bx + p.ahead;
                if (next / gx < (long long)p.C * (p.M + 1)) {
                    const char* ns = reinterpret_cast<const char*>(p.s1 + (size_t)(next / gx) * N * N + (size_t)(next % gx) * CT);
                    constexpr int kLines = (CT * (int)sizeof(float2) + 127) / 128;
                    for (int i = tid; i < N * kLines; i += kThreads)
                        prefetch_l2(ns + (size_t)(i / kLines) * N * sizeof(float2) + (i % kLines) * 128);
                }
            }
        }
        if constexpr (P::S == 1) {
            fwd_stage<P, 0, true, TwS>(tau, ld_g, st_g);
        } else if constexpr (PH == 0) {
            fwd_stage<P, 0, true, TwS>(tau, ld_g, st_s);
        } else if constexpr (PH < P::S - 1) {
            fwd_stage<P, PH, false, TwS>(tau, ld_s, st_s);
        } else {
            fwd_stage<P, P::S - 1, false, TwS>(tau, ld_s, st_g);
        }
    }
};

// ---------------------------------------------------------------------------
// K4: inverse FFT along H, keep h < N.
// ---------------------------------------------------------------------------
template <class P, int CT_> struct RowInv {
    using TwS = TwNone;                        // constant-bank twiddles (a block-local table measured slower here)
    static constexpr int L = P::L, N = L / 2, CT = CT_, kThreads = P::TL * CT;
    static constexpr int kPhases = P::S;
    static constexpr bool kWarpSync = false;
    static constexpr int kMinBlocks = (P::E >= 32) ? (512 / kThreads > 0 ? 512 / kThreads : 1)
                                                   : ((kThreads >= 1024) ? 1 : ((1024 / kThreads) > 32 ? 32 : (1024 / kThreads)));   // <= 64 regs
    static constexpr size_t kSmem = TwS::kBytes + ((P::S > 1) ? (size_t)L * CT * sizeof(float2) : 0);
    struct Regs {};
    static void grid(const Params& p, int& gx, int& gy) { gx = p.N / CT; gy = p.C * (p.M + 1); }
    static int iterations(const Params&) { return 1; }
    static constexpr bool kSinglePass = true;       // the driver runs the phases once, without a loop: K4 273 -> 252 us at 8 x 512x128^2, 146 -> 135 at 16 x 128^3

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int col = tid % CT, tau = line_thread<CT>(tid);
        float2* zs = reinterpret_cast<float2*>(smem);
        const float2* src = p.s2 + (size_t)by * L * N + bx * CT + col;
        float2* dst = p.s1 + (size_t)by * N * N + bx * CT + col;
        // L2-only loads: K4 246 -> 242 us at cfg3, 133 -> 131 at cfg4
        auto ld_g = [&](int pos, int slot) { return LCT_LDCG(src + (size_t)P::template freq_of<P::S - 1>(pos, slot) * N); };
        auto ld_s = [&](int pos, int) { return zs[pos * CT + col]; };
        auto st_s = [&](int pos, int, float2 v) { zs[pos * CT + col] = v; };
        auto st_g = [&](int pos, int, float2 v) { dst[(size_t)pos * N] = v; };
        if constexpr (PH == 0) {
            if (p.ahead > 0) {
                const int gx = N / CT;
                const long long next = (long long)by * gx + bx + p.ahead;
                if (next / gx < (long long)p.C * (p.M + 1)) {
                    const char* ns = reinterpret_cast<const char*>(p.s2 + (size_t)(next / gx) * L * N + (size_t)(next % gx) * CT);
                    constexpr int kLines = (CT * (int)sizeof(float2) + 127) / 128;
                    for (int i = tid; i < L * kLines; i += kThreads)
                        prefetch_l2(ns + (size_t)(i / kLines) * N * sizeof(float2) + (i % kLines) * 128);
                }
            }
        }
        if constexpr (P::S == 1) {
            inv_stage<P, 0, true, TwS>(tau, ld_g, st_g);
        } else if constexpr (PH == 0) {
            inv_stage<P, P::S - 1, false, TwS>(tau, ld_g, st_s);
        } else if constexpr (PH < P::S - 1) {
            inv_stage<P, P::S - 1 - PH, false, TwS>(tau, ld_s, st_s);
        } else {
            inv_stage<P, 0, true, TwS>(tau, ld_s, st_g);
        }
    }
};

// ---------------------------------------------------------------------------
// K2 / K4 for long columns (N = 256): the zero-extended 2N-point transform along H written as two N-point
// transforms by output parity -- even rows FFT_N(x), odd rows FFT_N(x w_2N^n); on the way back
// y[n] = IFFT_N(X_even)[n] + conj(w_2N^n) IFFT_N(X_odd)[n].  The two parities run side by side: threads
// [0, TL*CT) own the even rows, threads [TL*CT, 2*TL*CT) the odd rows, each half with its own exchange buffer
// (the inverse adds the halves in a third phase).  16-wide butterflies at 64 registers and two 512-thread blocks
// per SM instead of 32-wide ones at 128 registers: cfg5 K2 240 -> 184 us, K4 225 -> 196 us.  (Running the parities
// one after the other through one buffer measured 196 us forward and, with the even half parked in the output
// rows, 251 us inverse.  Per-half named barriers between the stages, as in the plane kernel: K2 182 us, K4 218 us -- kept out.)
// P is the N-point plan; S2 rows stay in natural kh order, so K3 is unchanged.
// ---------------------------------------------------------------------------
template <class P, int CT_> struct RowFwdSplit {
    static_assert(P::S == 2, "RowFwdSplit needs a two-stage plan");
    using TwS = TwNone;
    static constexpr int N = P::L, L = 2 * N, CT = CT_, kHalf = P::TL * CT, kThreads = 2 * kHalf;
    static constexpr int kPhases = 2;
    static constexpr bool kWarpSync = false;
    static constexpr int kMinBlocks = (kThreads >= 1024) ? 1 : 1024 / kThreads;      // <= 64 regs
    static constexpr size_t kSmem = TwS::kBytes + (size_t)2 * N * CT * sizeof(float2);
    static_assert(kHalf % 32 == 0, "a warp must not straddle the two parities");
    struct Regs {};
    static void grid(const Params& p, int& gx, int& gy) { gx = p.N / CT; gy = p.C * (p.M + 1); }
    static int iterations(const Params&) { return 1; }
    static constexpr bool kSinglePass = false;      // (single pass measured slower here: 188 vs 178 us at 512x256^2)

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int) {
        unsigned char* smem = smem_base + TwS::kBytes;
        const int parity = tid / kHalf, t = tid % kHalf;
        const int col = t % CT, tau = line_thread<CT>(t);
        float2* z = reinterpret_cast<float2*>(smem) + parity * (N * CT);
        auto st_s = [&](int pos, int, float2 v) { z[pos * CT + col] = v; };
        if constexpr (PH == 0) {
            const float2* src = p.s1 + (size_t)by * N * N + bx * CT + col;
            if (p.ahead > 0) {
                const int gx = N / CT;
                const long long next = (long long)by * gx + bx + p.ahead;
                if (next / gx < (long long)p.C * (p.M + 1)) {
                    const char* ns = reinterpret_cast<const char*>(p.s1 + (size_t)(next / gx) * N * N + (size_t)(next % gx) * CT);
                    constexpr int kLines = (CT * (int)sizeof(float2) + 127) / 128;
                    for (int i = tid; i < N * kLines; i += kThreads)
                        prefetch_l2(ns + (size_t)(i / kLines) * N * sizeof(float2) + (i % kLines) * 128);
                }
            }
            if (parity == 0)
                fwd_stage<P, 0, false, TwS>(tau, [&](int pos, int) { return src[(size_t)pos * N]; }, st_s);
            else if constexpr (LCT_SPLIT_TW_CONST && P::TL == P::st(0)) {
                const float2 wt = TwGlobal::get(tau * (kTwN / L));
                fwd_stage<P, 0, false, TwS>(tau,
                    [&](int pos, int slot) { return split_twiddle<P::R0, false>(src[(size_t)pos * N], wt, slot % P::R0); }, st_s);
            } else
                fwd_stage<P, 0, false, TwS>(tau,
                    [&](int pos, int) { return TwGlobal::mul(src[(size_t)pos * N], pos * (kTwN / L)); }, st_s);
        } else {
            float2* dst = p.s2 + (size_t)by * L * N + bx * CT + col + (size_t)parity * N;
            fwd_stage<P, 1, false, TwS>(tau,
                [&](int pos, int) { return z[pos * CT + col]; },
                [&](int pos, int slot, float2 v) { dst[(size_t)P::template freq_of<1>(pos, slot) * (2 * N)] = v; });
        }
    }
};

template <class P, int CT_> struct RowInvSplit {
    static_assert(P::S == 2, "RowInvSplit needs a two-stage plan");
    using TwS = TwNone;
    static constexpr int N = P::L, L = 2 * N, CT = CT_, kHalf = P::TL * CT, kThreads = 2 * kHalf;
    static constexpr int kPhases = 3;
    static constexpr bool kWarpSync = false;
    static constexpr int kMinBlocks = (kThreads >= 1024) ? 1 : 1024 / kThreads;      // <= 64 regs
    static constexpr size_t kSmem = TwS::kBytes + (size_t)2 * N * CT * sizeof(float2);
    static_assert(kHalf % 32 == 0, "a warp must not straddle the two parities");
    struct Regs {};
    static void grid(const Params& p, int& gx, int& gy) { gx = p.N / CT; gy = p.C * (p.M + 1); }
    static int iterations(const Params&) { return 1; }
    static constexpr bool kSinglePass = false;      // (single pass: no difference)

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwS::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs&, unsigned char* smem_base, int tid, int bx, int by, int) {
        unsigned char* smem = smem_base + TwS::kBytes;
        float2* zs = reinterpret_cast<float2*>(smem);
        if constexpr (PH < 2) {
            const int parity = tid / kHalf, t = tid % kHalf;
            const int col = t % CT, tau = line_thread<CT>(t);
            float2* z = zs + parity * (N * CT);
            auto ld_s = [&](int pos, int) { return z[pos * CT + col]; };
            auto st_s = [&](int pos, int, float2 v) { z[pos * CT + col] = v; };
            if constexpr (PH == 0) {
                const float2* src = p.s2 + (size_t)by * L * N + bx * CT + col + (size_t)parity * N;
                if (p.ahead > 0) {
                    const int gx = N / CT;
                    const long long next = (long long)by * gx + bx + p.ahead;
                    if (next / gx < (long long)p.C * (p.M + 1)) {
                        const char* ns = reinterpret_cast<const char*>(p.s2 + (size_t)(next / gx) * L * N + (size_t)(next % gx) * CT);
                        constexpr int kLines = (CT * (int)sizeof(float2) + 127) / 128;
                        for (int i = tid; i < L * kLines; i += kThreads)
                            prefetch_l2(ns + (size_t)(i / kLines) * N * sizeof(float2) + (i % kLines) * 128);
                    }
                }
                inv_stage<P, 1, false, TwS>(tau,
                    [&](int pos, int slot) { return LCT_LDCG(src + (size_t)P::template freq_of<1>(pos, slot) * (2 * N)); }, st_s);
            } else {
                inv_stage<P, 0, false, TwS>(tau, ld_s, st_s);       // in place: a butterfly writes the positions it read
            }
        } else {
            // y[n] = y_even[n] + conj(w_2N^n) y_odd[n]
            float2* dst = p.s1 + (size_t)by * N * N + bx * CT;
            const float2* zo = zs + N * CT;
            constexpr int kIters = N * CT / kThreads;
            static_assert(N * CT % kThreads == 0 && kThreads % CT == 0, "the plane tile must divide among the threads");
            // n = tid / CT + u * kStep: conj(w_2N^n) = conj(w_2N^(tid / CT)) * conj of a compile-time root of unity
            constexpr int kStep = kThreads / CT;
            constexpr bool kConstTw = LCT_SPLIT_TW_CONST && (32 * kStep) % L == 0;
            const float2 wn0 = TwGlobal::get((tid / CT) * (kTwN / L));
            LCT_UNROLL
            for (int u = 0; u < kIters; ++u) {
                const int i = tid + u * kThreads, n = i / CT, col = i % CT;
                if constexpr (kConstTw) dst[(size_t)n * N + col] = cadd(zs[i], twmul32<true>(cmulc(zo[i], wn0), u * (32 * kStep / L)));
                else dst[(size_t)n * N + col] = cadd(zs[i], TwGlobal::mulc(zo[i], n * (kTwN / L)));
            }
        }
    }
};

enum FilterLayout { kFilterFull = 0, kFilterQuarter = 1, kFilterHalfRows = 2 };
// Symmetric-filter addressing (Params::filt_sym): index of frequency k in the stored range [0, N] and the exponent
// of the twiddle w_2N that maps the stored value onto W(k) (0 for k <= N).
template <int N> LCT_DEV int sym_index(int k) { return k <= N ? k : 2 * N - k; }
template <int N> LCT_DEV int sym_twist(int k) { return k <= N ? 0 : k; }

// ---------------------------------------------------------------------------
// K3: along W (the contiguous axis): zero-extended FFT, filter multiply, inverse
// FFT, crop -- in place on S2.  One block = RB consecutive kh rows of one kt
// plane; it loops over the channels so the filter row stays in registers.
// Two-stage plans only (lanes run along the line so global accesses coalesce).
// ---------------------------------------------------------------------------
// (An instantiation without the channel loop for single-channel launches, as the H-axis kernels have it, measured slower:
//  416 vs 408 us at 512x256^2, 123 vs 119 us at 1 x 512x128^2.  Channel 0's row requested before the prologue barrier, as
//  in the plane kernel: no difference here, 430 vs 408 us in the parity-split kernel at one channel.)
template <class P, int RB_, bool SYM = false> struct ColFilter {
    static_assert(P::S == 2, "ColFilter needs a two-stage plan");
    static_assert(32 % P::TL == 0, "the threads of one line must share a warp");
    static constexpr int L = P::L, N = L / 2, RB = RB_, kThreads = P::TL * RB;
    static constexpr int kPhases = 3;
    // block-local [k][lo] table for the 8/16-wide plans; the 32-wide plan (N = 256) measured faster on __ldg
    using TwL = typename std::conditional<(P::E <= 16), TwLine<P>, TwGlobal>::type;
    // all threads of a line sit in one warp (tid % TL), so the exchanges only need __syncwarp:
    // warps run through the channel loop independently of each other.
    static constexpr bool kWarpSync = true;
#ifndef LCT_COLFILTER_BLOCKS
#define LCT_COLFILTER_BLOCKS 2
#endif
    static constexpr int kMinBlocks = (P::E <= 16) ? LCT_COLFILTER_BLOCKS : 1;
    static constexpr int PAD = 1;
    static constexpr int RS = L + P::R0 * PAD + ((P::TL < 16) ? 8 : 0);    // row stride in float2
    static constexpr size_t kSmem = TwL::kBytes + (size_t)RB * RS * sizeof(float2);
    static constexpr int kIn = P::E / 2;                                   // non-zero inputs per thread
    static constexpr int kFiltPitch = SYM ? N + 1 : L;                     // stored filter row, in values (kFilterQuarter)
    struct Regs { float2 w[P::E]; float2 pre[kIn]; };
    static void grid(const Params& p, int& gx, int& gy) { gx = L / RB; gy = p.M + 1; }
    static int iterations(const Params& p) { return p.C; }
    static LCT_DEV int padpos(int pos) { return pos + (pos / P::st(0)) * PAD; }

    // loads the thread's share of one input row (positions < N) into r.pre
    static LCT_DEV void fetch(const float2* row, int tau, Regs& r) {
        constexpr int r0 = P::R0, str = P::st(0), NB = L / r0;
        LCT_UNROLL
        for (int m = 0; m < NB / P::TL; ++m) {
            const int lo = tau + m * P::TL;
            LCT_UNROLL
            for (int q = 0; q < r0 / 2; ++q) r.pre[m * (r0 / 2) + q] = row[lo + q * str];
        }
    }

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwL::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs& r, unsigned char* smem, int tid, int bx, int by, int it) {
        const int tau = tid % P::TL, rl = tid / P::TL;
        const int kh = bx * RB + rl, kt = by, c = it;
        float2* zs = reinterpret_cast<float2*>(smem + TwL::kBytes) + rl * RS;
        const size_t chan = (size_t)(p.M + 1) * L * N;
        float2* row = p.s2 + (size_t)c * chan + ((size_t)kt * L + kh) * N;
        if constexpr (PH == 0) {
            if (it == 0 && p.ahead > 0) {
                // warm L2 with channel 0's rows and the filter rows of the block one residency ahead
                const int gx = L / RB;
                const long long next = (long long)by * gx + bx + p.ahead;
                if (next / gx <= p.M) {
                    const size_t first = (size_t)(next / gx) * L + (size_t)(next % gx) * RB;          // kt * L + kh
                    const char* rows = reinterpret_cast<const char*>(p.s2 + first * N);
                    constexpr int kRowLines = RB * N * (int)sizeof(float2) / 128;
                    for (int i = tid; i < kRowLines; i += kThreads) prefetch_l2(rows + (size_t)i * 128);
                    if constexpr (SYM) {
                        // the RB stored rows this block will read: kh' descends when kh > N
                        const int nkh = (int)(next % gx) * RB, lo = nkh < N ? nkh : 2 * N - (nkh + RB - 1);
                        const int nrows = (lo + RB <= N + 1) ? RB : N + 1 - lo;
                        const char* filt = reinterpret_cast<const char*>(p.filt + ((size_t)(next / gx) * (N + 1) + lo) * kFiltPitch);
                        for (int i = tid; i < (nrows * kFiltPitch * (int)sizeof(float2) + 127) / 128; i += kThreads) prefetch_l2(filt + (size_t)i * 128);
                    } else {
                        const char* filt = reinterpret_cast<const char*>(p.filt + first * L);
                        constexpr int kFiltLines = RB * L * (int)sizeof(float2) / 128;
                        for (int i = tid; i < kFiltLines; i += kThreads) prefetch_l2(filt + (size_t)i * 128);
                    }
                }
            }
            if (it == 0) {
                fetch(row, tau, r);
                if constexpr (SYM) {
                    // quarter filter: stored row kh' = min(kh, 2N - kh), stored column kw' likewise; one twiddle
                    // w_2N^(twist(kh) + twist(kw)) per value, paid once per block (the row stays in registers)
                    const float2* f = p.filt + ((size_t)kt * (N + 1) + sym_index<N>(kh)) * (N + 1);
                    const int th = sym_twist<N>(kh);
                    for_each_slot<P, 1>(tau, [&](int pos, int slot) {
                        const int kw = P::template freq_of<1>(pos, slot);
                        float2 w = LCT_LDG(f + sym_index<N>(kw));
                        w = TwGlobal::mul(w, ((th + sym_twist<N>(kw)) & (L - 1)) * (kTwN / L));
                        if (p.conj_filter) w.y = -w.y;
                        r.w[slot] = w;
                    });
                } else {
                    const float2* f = p.filt + ((size_t)kt * L + kh) * L;
                    for_each_slot<P, 1>(tau, [&](int pos, int slot) {
                        float2 w = LCT_LDG(f + P::template freq_of<1>(pos, slot));
                        if (p.conj_filter) w.y = -w.y;
                        r.w[slot] = w;
                    });
                }
            }
            constexpr int r0 = P::R0;
            fwd_stage<P, 0, true, TwL>(tau,
                [&](int, int slot) { return r.pre[(slot / r0) * (r0 / 2) + (slot % r0)]; },
                [&](int pos, int, float2 v) { zs[padpos(pos)] = v; });
            if (it + 1 < p.C) fetch(row + chan, tau, r);        // next channel's row flies during this one's math
        } else if constexpr (PH == 1) {
            float2 a[P::E];
            fwd_stage<P, 1, false, TwL>(tau,
                [&](int pos, int) { return zs[padpos(pos)]; },
                [&](int, int slot, float2 v) { a[slot] = cmul(v, r.w[slot]); });
            inv_stage<P, 1, false, TwL>(tau,
                [&](int, int slot) { return a[slot]; },
                [&](int pos, int, float2 v) { zs[padpos(pos)] = v; });
        } else {
            inv_stage<P, 0, true, TwL>(tau,
                [&](int pos, int) { return zs[padpos(pos)]; },
                [&](int pos, int, float2 v) { row[pos] = v; });
        }
    }
};

// ---------------------------------------------------------------------------
// K3 for long lines (N = 256): the same pass with the zero-extended 2N-point transform written as two
// N-point transforms -- even output frequencies FFT_N(x), odd ones FFT_N(x w_2N^n) -- each filtered and
// inverted, y[n] = y_even[n] + conj(w_2N^n) y_odd[n].  Same arithmetic as ColFilter, but the butterflies
// are 16 wide instead of 32: a third of the registers and twice the resident warps.  P is the N-point plan.
// ---------------------------------------------------------------------------
template <class P, int RB_, bool SYM = false> struct ColFilterSplit {
    static_assert(P::S == 2, "ColFilterSplit needs a two-stage plan");
    static_assert(32 % P::TL == 0, "the threads of one line must share a warp");
    static constexpr int N = P::L, L = 2 * N, RB = RB_, kThreads = P::TL * RB;
    static constexpr int kPhases = 3;
    using TwL = TwLine<P>;
    static constexpr bool kWarpSync = true;
    static constexpr int kMinBlocks = 2;
    static constexpr int PAD = 1;
    static constexpr int RS = N + P::R0 * PAD + ((P::TL < 16) ? 8 : 0);    // row stride in float2
    static constexpr size_t kSmem = TwL::kBytes + (size_t)2 * RB * RS * sizeof(float2);     // even and odd exchange rows
    static constexpr int kFiltPitch = L;                                   // stored filter row, in values (kFilterHalfRows keeps whole rows)
    // Requesting the even parity's filter values a stage early (before the odd stage-0 transform) and the odd parity's
    // before the even inverse stage was measured at cfg5: 527 vs 496 us -- the 32 values held across the stage push the
    // kernel from 60 to 116 bytes of spills at its 128-register cap, which costs more than the hidden latency returns.
#ifndef LCT_K3S_EARLY_FILTER
#define LCT_K3S_EARLY_FILTER 0
#endif
    static constexpr bool kEarlyFilter = LCT_K3S_EARLY_FILTER != 0;
    struct Regs { float2 in[P::E]; float2 w[P::E]; };

    // r.w <- the filter values of this thread's stage-1 slots for one output parity (kw = 2 f + parity).
    // Natural layout [kt][kh][kw]; half-rows layout (SYM, kFilterHalfRows): stored row min(kh, 2N - kh), times w_2N^kh
    // when mirrored (one row-constant twiddle).
    static LCT_DEV void load_filter(const Params& p, Regs& r, int tau, int kt, int kh, int parity) {
        const float2* f = SYM ? p.filt + ((size_t)kt * (N + 1) + sym_index<N>(kh)) * L
                              : p.filt + ((size_t)kt * L + kh) * L;
        float2 rowtw = make_float2(1.f, 0.f);
        if constexpr (SYM) rowtw = TwGlobal::get(sym_twist<N>(kh) * (kTwN / L));
        for_each_slot<P, 1>(tau, [&](int pos, int slot) {
            float2 v = LCT_LDG(f + 2 * P::template freq_of<1>(pos, slot) + parity);
            if constexpr (SYM) v = cmul(v, rowtw);
            if (p.conj_filter) v.y = -v.y;
            r.w[slot] = v;
        });
    }
    static void grid(const Params& p, int& gx, int& gy) { gx = L / RB; gy = p.M + 1; }
    static int iterations(const Params& p) { return p.C; }
    static LCT_DEV int padpos(int pos) { return pos + (pos / P::st(0)) * PAD; }

    static LCT_DEV void fetch(const float2* row, int tau, Regs& r) {
        for_each_slot<P, 0>(tau, [&](int pos, int slot) { r.in[slot] = row[pos]; });
    }

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) { TwL::fill(smem, tid, kThreads); }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs& r, unsigned char* smem, int tid, int bx, int by, int it) {
        const int tau = tid % P::TL, rl = tid / P::TL;
        const int kh = bx * RB + rl, kt = by, c = it;
        float2* ze = reinterpret_cast<float2*>(smem + TwL::kBytes) + rl * RS;
        float2* zo = ze + RB * RS;
        const size_t chan = (size_t)(p.M + 1) * L * N;
        float2* row = p.s2 + (size_t)c * chan + ((size_t)kt * L + kh) * N;
        if constexpr (PH == 0) {
            if (it == 0 && p.ahead > 0) {
                // warm L2 with channel 0's rows and the filter rows of the block one residency ahead
                const int gx = L / RB;
                const long long next = (long long)by * gx + bx + p.ahead;
                if (next / gx <= p.M) {
                    const size_t first = (size_t)(next / gx) * L + (size_t)(next % gx) * RB;          // kt * L + kh
                    const char* rows = reinterpret_cast<const char*>(p.s2 + first * N);
                    constexpr int kRowLines = RB * N * (int)sizeof(float2) / 128;
                    for (int i = tid; i < kRowLines; i += kThreads) prefetch_l2(rows + (size_t)i * 128);
                    if constexpr (SYM) {
                        // the RB stored rows this block will read: kh' descends when kh > N
                        const int nkh = (int)(next % gx) * RB, lo = nkh < N ? nkh : 2 * N - (nkh + RB - 1);
                        const int nrows = (lo + RB <= N + 1) ? RB : N + 1 - lo;
                        const char* filt = reinterpret_cast<const char*>(p.filt + ((size_t)(next / gx) * (N + 1) + lo) * kFiltPitch);
                        for (int i = tid; i < (nrows * kFiltPitch * (int)sizeof(float2) + 127) / 128; i += kThreads) prefetch_l2(filt + (size_t)i * 128);
                    } else {
                        const char* filt = reinterpret_cast<const char*>(p.filt + first * L);
                        constexpr int kFiltLines = RB * L * (int)sizeof(float2) / 128;
                        for (int i = tid; i < kFiltLines; i += kThreads) prefetch_l2(filt + (size_t)i * 128);
                    }
                }
            }
            if (it == 0) fetch(row, tau, r);
            fwd_stage<P, 0, false, TwL>(tau,
                [&](int, int slot) { return r.in[slot]; },
                [&](int pos, int, float2 v) { ze[padpos(pos)] = v; });
            // the even parity's filter values are requested here, a whole stage before their first use: at one channel
            // per launch nothing else hides their latency (it was 22 % of the kernel's stall samples)
            if constexpr (kEarlyFilter) load_filter(p, r, tau, kt, kh, 0);
            if constexpr (LCT_SPLIT_TW_CONST && P::TL == P::st(0)) {
                const float2 wt = TwGlobal::get(tau * (kTwN / L));
                fwd_stage<P, 0, false, TwL>(tau,
                    [&](int, int slot) { return split_twiddle<P::R0, false>(r.in[slot], wt, slot % P::R0); },
                    [&](int pos, int, float2 v) { zo[padpos(pos)] = v; });
            } else
                fwd_stage<P, 0, false, TwL>(tau,
                    [&](int pos, int slot) { return TwGlobal::mul(r.in[slot], pos * (kTwN / L)); },
                    [&](int pos, int, float2 v) { zo[padpos(pos)] = v; });
            if (it + 1 < p.C) fetch(row + chan, tau, r);        // next channel's row flies during this one's math
        } else if constexpr (PH == 1) {
            float2 a[P::E];
            if constexpr (!kEarlyFilter) load_filter(p, r, tau, kt, kh, 0);
            fwd_stage<P, 1, false, TwL>(tau,
                [&](int pos, int) { return ze[padpos(pos)]; },
                [&](int, int slot, float2 v) { a[slot] = cmul(v, r.w[slot]); });
            if constexpr (kEarlyFilter) load_filter(p, r, tau, kt, kh, 1);          // odd parity: flies during the even inverse stage
            inv_stage<P, 1, false, TwL>(tau,
                [&](int, int slot) { return a[slot]; },
                [&](int pos, int, float2 v) { ze[padpos(pos)] = v; });
            if constexpr (!kEarlyFilter) load_filter(p, r, tau, kt, kh, 1);
            fwd_stage<P, 1, false, TwL>(tau,
                [&](int pos, int) { return zo[padpos(pos)]; },
                [&](int, int slot, float2 v) { a[slot] = cmul(v, r.w[slot]); });
            inv_stage<P, 1, false, TwL>(tau,
                [&](int, int slot) { return a[slot]; },
                [&](int pos, int, float2 v) { zo[padpos(pos)] = v; });
        } else {
            float2 ya[P::E];
            inv_stage<P, 0, false, TwL>(tau,
                [&](int pos, int) { return ze[padpos(pos)]; },
                [&](int, int slot, float2 v) { ya[slot] = v; });
            if constexpr (LCT_SPLIT_TW_CONST && P::TL == P::st(0)) {
                const float2 wt = TwGlobal::get(tau * (kTwN / L));
                inv_stage<P, 0, false, TwL>(tau,
                    [&](int pos, int) { return zo[padpos(pos)]; },
                    [&](int pos, int slot, float2 v) { row[pos] = cadd(ya[slot], split_twiddle<P::R0, true>(v, wt, slot % P::R0)); });
            } else
                inv_stage<P, 0, false, TwL>(tau,
                    [&](int pos, int) { return zo[padpos(pos)]; },
                    [&](int pos, int slot, float2 v) { row[pos] = cadd(ya[slot], TwGlobal::mulc(v, pos * (kTwN / L))); });
        }
    }
};

// ---------------------------------------------------------------------------
// Plane-resident fusion of K2 + K3 + K4 (N <= 64): one block owns one (c, kt) plane.
// The zero-extended 2N x N half-transformed plane lives in shared memory (row stride N+1
// float2, conflict-free for lanes along W and for lanes along H), so the plane is read
// from and written to HBM exactly once: 16 B per spectrum element instead of 80 B.
//   H forward  : K2's stages, lanes along W, output rows stay in plan position order.
//   W pass     : per row, the 2N-point zero-extended FFT is two N-point FFTs (even / odd
//                output frequencies, the odd one pre-rotated by w_2N^n); each is filtered and
//                inverted -- the even one in the row's own N slots, the odd one in a 1-batch side
//                buffer -- and the two results recombined (y = y_even + conj(w_2N^n) y_odd).  Lanes run along H here, so twiddles are
//                warp-uniform and the filter, stored as [kt][kw/2][row][kw&1], is read coalesced.
//   H inverse  : K4's stages, written back over the input plane in S1.
// ---------------------------------------------------------------------------
//
// PERSIST (N = 64): about as many blocks as are resident together walk the planes (channel fastest, so the blocks
// that are resident together share a filter plane), and while a block runs the two H-inverse phases of one plane, the bulk-copy
// unit brings the next plane -- 8 N^2 contiguous bytes of S1 -- into the side buffer X, which the W pass has just
// released; the first H stage then reads its inputs from shared memory instead of waiting for L2.
#ifndef LCT_PLANE_PERSIST
#define LCT_PLANE_PERSIST 0
#endif
template <class PHp, class PWp, int NT_, bool PERSIST = false> struct PlaneFilter {
    static constexpr int L = PHp::L, N = L / 2;
    static_assert(PWp::L == N && PWp::S == 2 && PHp::S == 2, "plane fusion needs two-stage plans");
    static constexpr int kThreads = NT_;
    static constexpr int CB = kThreads / PHp::TL;                // columns per H batch
    static_assert(CB >= 1 && N % CB == 0, "column batches must tile the plane");
    static constexpr int nHB = N / CB;
    static constexpr int RBt = (kThreads / PWp::TL) < L ? (kThreads / PWp::TL) : L;   // rows per W batch
    static_assert(L % RBt == 0, "row batches must tile the plane");
    static constexpr int nWB = L / RBt;
    static constexpr int RS = N + 1;
    using TwP = TwShared<L>;
    static constexpr size_t kTwBytes = TwP::kBytes;
    // plane T[L][RS] + side buffer X[RBt][RS]: the odd-parity transform of a row batch is exchanged
    // through X while the even one uses the rows' own slots, so both run in the same three phases
    static constexpr size_t kPlaneBytes = (size_t)N * N * sizeof(float2);
    static_assert(!PERSIST || (size_t)RBt * RS * sizeof(float2) >= kPlaneBytes, "the next plane is staged in X");
    static_assert(!PERSIST || (kTwBytes + (size_t)L * RS * sizeof(float2)) % 16 == 0, "bulk copies land on 16-byte boundaries");
    static constexpr size_t kBarOff = kTwBytes + (size_t)(L + RBt) * RS * sizeof(float2);
    static constexpr size_t kSmem = kBarOff + (PERSIST ? 16 : 0);
    static constexpr int kPhases = 2 + nWB * 3 + 2;
    static constexpr int EW = PWp::E;
    static constexpr bool kWarpSync = false;
    static constexpr int kMinBlocks = (2 * kSmem <= 220 * 1024) ? 2 : 1;      // two blocks per SM at <= 64 registers
#ifndef LCT_PLANE_PRELOAD
#define LCT_PLANE_PRELOAD 1
#endif
    // The first H stage's inputs are requested before the barrier that ends the prologue (the twiddle-table fill), so
    // the plane's L2 round trip and the table's overlap instead of following each other at the start of every block.
    static constexpr bool kPreload = LCT_PLANE_PRELOAD && !PERSIST && nHB == 1;
    struct Regs { float2 in[kPreload ? PHp::E : 1]; };
    static LCT_DEV void preload(const Params& p, Regs& r, unsigned char*, int tid, int bx, int by) {
        if constexpr (kPreload) {
            const float2* src = p.s1 + ((size_t)bx * (p.M + 1) + by) * N * N + tid % CB;
            for_each_slot_lower<PHp, 0>(line_thread<CB>(tid), [&](int pos, int slot) { r.in[slot] = src[(size_t)pos * N]; });
        }
    }
    // PERSIST: grid (C, R) with R = resident blocks / C rows of planes; block (c, r) walks kt = r, r + R, ... -- a
    // two-dimensional grid keeps the plane index free of divisions and in the uniform datapath (a flat walk that
    // split its index by C in every phase cost the kernel 7 % more instructions than the prefetch saved)
    static LCT_HD int walk_rows(const Params& p) {
        const int r = (PERSIST && p.ahead > 0) ? p.ahead / p.C : p.M + 1;
        return r < 1 ? 1 : (r > p.M + 1 ? p.M + 1 : r);
    }
    static void grid(const Params& p, int& gx, int& gy) { gx = p.C; gy = walk_rows(p); }   // c fastest: filter plane reused from L2
    static int iterations(const Params& p) { return (p.M + 1 + walk_rows(p) - 1) / walk_rows(p); }
    static LCT_DEV int rows_of_grid(const Params& p) {
#ifdef LCT_EMULATE
        return walk_rows(p);
#else
        return (int)gridDim.y;
#endif
    }
    static constexpr bool kTileWalks = true;              // see TimeInv::walk_active
    static constexpr bool kSinglePass = !PERSIST;         // iterations() == 1: the driver runs the phases without a loop
    static LCT_DEV bool walk_active(const Params& p, int, int by, int it) { return !PERSIST || by + it * rows_of_grid(p) <= p.M; }

    static constexpr bool kHasPrologue = true;
    static LCT_DEV void prologue(const Params&, Regs&, unsigned char* smem, int tid, int, int) {
        TwP::fill(smem, tid, kThreads);
        if (PERSIST && tid == 0) mbar_init(reinterpret_cast<unsigned long long*>(smem + kBarOff), 1);
    }
    // plane (c, kt) -> X, as N x N dense c64; one thread asks, the barrier collects
    static LCT_DEV void stage_plane(const Params& p, unsigned char* smem, int c, int kt) {
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + kBarOff);
        fence_proxy_async();
        mbar_arrive_expect(bar, (unsigned)kPlaneBytes);
        bulk_load(smem + kTwBytes + (size_t)L * RS * sizeof(float2), p.s1 + ((size_t)c * (p.M + 1) + kt) * N * N,
                  (unsigned)kPlaneBytes, bar);
    }
#ifndef LCT_PLANE_GROUPSYNC
#define LCT_PLANE_GROUPSYNC 2
#endif
    // During the W pass a row is shared by the TL warps that hold its line threads: with 64-row batches and 32-lane
    // warps those are the even warps (rows 0..31 of a batch) or the odd warps (rows 32..63).  The two sets touch
    // disjoint rows of T and X, so between W phases each set synchronises on its own named barrier and the sets drift
    // apart -- one can be in its shared-memory burst while the other is in its butterflies.
    static constexpr bool kGroupSync = LCT_PLANE_GROUPSYNC && RBt == 64 && RBt * PWp::TL == kThreads;
    static constexpr int kGroupThreads = kThreads / 2;
    static LCT_DEV int group_of(int tid) { return (tid >> 5) & 1; }
    // The same two warp sets own disjoint column halves in the H passes (64-column batches), so the barrier between
    // the two H stages is a set barrier as well; only the H <-> W turns need the whole block.
    static constexpr bool kGroupSyncH = (LCT_PLANE_GROUPSYNC >= 2) && kGroupSync && CB == 64 && nHB == 1;
    static constexpr bool group_phase(int ph) {
        return (kGroupSync && ph >= 2 && ph < 2 + 3 * nWB - 1) || (kGroupSyncH && (ph == 0 || ph == 2 + 3 * nWB));
    }

    template <int PH> static LCT_DEV void phase(const Params& p, Regs& r, unsigned char* smem, int tid, int bx, int by, int it) {
        float2* T = reinterpret_cast<float2*>(smem + kTwBytes);
        float2* X = T + L * RS;
        const int rows = PERSIST ? rows_of_grid(p) : 0, kt = by + it * rows;
#ifdef LCT_EMULATE
        if (kt > p.M) return;                                // on the GPU the driver leaves the loop (walk_active)
#endif
        const size_t plane = (size_t)bx * (p.M + 1) + kt;
        constexpr int kW0 = 2, kW1 = 2 + 3 * nWB;
        if constexpr (PH == 0 && PERSIST) {
#ifdef LCT_EMULATE
            if (it == 0) stage_plane(p, smem, bx, kt);       // every emulated thread copies before it reads (any thread order)
#else
            if (it == 0 && tid == 0) stage_plane(p, smem, bx, kt);        // later planes were asked for during the H inverse
#endif
            mbar_wait(reinterpret_cast<unsigned long long*>(smem + kBarOff), (unsigned)(it & 1));
        }
        if constexpr (PH == kW1 && PERSIST) {
            // the W pass is over (whole-block barrier): X is free until the next plane's W pass
            if (tid == 0 && kt + rows <= p.M) stage_plane(p, smem, bx, kt + rows);
        }
        if constexpr (PH < kW0) {
            // H forward; column batches touch disjoint columns, so they share a phase
            const int tau = line_thread<CB>(tid);
            LCT_UNROLL
            for (int hb = 0; hb < nHB; ++hb) {
                const int col = hb * CB + tid % CB;
                const float2* src = PERSIST ? X + col : p.s1 + plane * N * N + col;
                if constexpr (PH == 0 && kPreload) {
                    fwd_stage<PHp, 0, true, TwP>(tau,
                        [&](int, int slot) { return r.in[slot]; },
                        [&](int pos, int, float2 v) { T[pos * RS + col] = v; });
                } else if constexpr (PH == 0) {
                    fwd_stage<PHp, 0, true, TwP>(tau,
                        [&](int pos, int) { return src[(size_t)pos * N]; },
                        [&](int pos, int, float2 v) { T[pos * RS + col] = v; });
                } else {
                    fwd_stage<PHp, 1, false, TwP>(tau,
                        [&](int pos, int) { return T[pos * RS + col]; },
                        [&](int pos, int, float2 v) { T[pos * RS + col] = v; });
                }
            }
        } else if constexpr (PH < kW1) {
            // W pass of one row batch: even (a) and odd (b) output parities side by side
            constexpr int wb = (PH - kW0) / 3, st3 = (PH - kW0) % 3;
            const int rl = tid % RBt, tau = line_thread<RBt>(tid), row = wb * RBt + rl;
            if (tid >= RBt * PWp::TL) return;
            float2* Tr = T + row * RS;
            float2* Xr = X + rl * RS;
            if constexpr (st3 == 0) {
                float2 in[EW], tw[StageTwiddles<PWp, 0>::kCount];
                load_stage_twiddles<PWp, 0, TwP>(tau, tw);              // shared by both parities
                for_each_slot<PWp, 0>(tau, [&](int pos, int slot) { in[slot] = Tr[pos]; });
                fwd_stage_tw<PWp, 0>(tau, tw,
                    [&](int, int slot) { return in[slot]; },
                    [&](int pos, int, float2 v) { Tr[pos] = v; });
#ifndef LCT_PLANE_TW_CONST
#define LCT_PLANE_TW_CONST 1
#endif
                if constexpr (LCT_PLANE_TW_CONST && PWp::TL == PWp::st(0)) {
                    const float2 wt = TwP::get(tau * (kTwN / L));          // w_2N^n = w_2N^tau * (compile-time root)
                    fwd_stage_tw<PWp, 0>(tau, tw,
                        [&](int, int slot) { return split_twiddle<PWp::R0, false>(in[slot], wt, slot % PWp::R0); },
                        [&](int pos, int, float2 v) { Xr[pos] = v; });
                } else
                    fwd_stage_tw<PWp, 0>(tau, tw,
                        [&](int pos, int slot) { return TwP::mul(in[slot], pos * (kTwN / L)); },
                        [&](int pos, int, float2 v) { Xr[pos] = v; });
#ifndef LCT_EMULATE
                {   // pull next phase's filter values from L2 towards L1 while the exchange settles
                    const float4* f = reinterpret_cast<const float4*>(p.filt) + (size_t)kt * N * L + row;
                    for_each_slot<PWp, 1>(tau, [&](int pos, int) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(f + (size_t)PWp::pos_to_freq(pos) * L));
                    });
                }
#endif
            } else if constexpr (st3 == 1) {
                // filter stored [kt][kw/2][row][kw&1]: one 128-bit load brings both parities' values
                const float4* f = reinterpret_cast<const float4*>(p.filt) + (size_t)kt * N * L + row;
                float2 wa[EW], wb_[EW];
                for_each_slot<PWp, 1>(tau, [&](int pos, int slot) {
                    const float4 w = LCT_LDG(f + (size_t)PWp::template freq_of<1>(pos, slot) * L);
                    wa[slot] = make_float2(w.x, w.y);
                    wb_[slot] = make_float2(w.z, w.w);
                });
                float2 b[EW];
                fwd_stage<PWp, 1, false, TwP>(tau,
                    [&](int pos, int) { return Tr[pos]; },
                    [&](int, int slot, float2 v) { b[slot] = p.conj_filter ? cmulc(v, wa[slot]) : cmul(v, wa[slot]); });
                inv_stage<PWp, 1, false, TwP>(tau,
                    [&](int, int slot) { return b[slot]; },
                    [&](int pos, int, float2 v) { Tr[pos] = v; });
                fwd_stage<PWp, 1, false, TwP>(tau,
                    [&](int pos, int) { return Xr[pos]; },
                    [&](int, int slot, float2 v) { b[slot] = p.conj_filter ? cmulc(v, wb_[slot]) : cmul(v, wb_[slot]); });
                inv_stage<PWp, 1, false, TwP>(tau,
                    [&](int, int slot) { return b[slot]; },
                    [&](int pos, int, float2 v) { Xr[pos] = v; });
            } else {
                float2 ya[EW], tw[StageTwiddles<PWp, 0>::kCount];
                load_stage_twiddles<PWp, 0, TwP>(tau, tw);
                inv_stage_tw<PWp, 0>(tau, tw,
                    [&](int pos, int) { return Tr[pos]; },
                    [&](int, int slot, float2 v) { ya[slot] = v; });
                if constexpr (LCT_PLANE_TW_CONST && PWp::TL == PWp::st(0)) {
                    const float2 wt = TwP::get(tau * (kTwN / L));
                    inv_stage_tw<PWp, 0>(tau, tw,
                        [&](int pos, int) { return Xr[pos]; },
                        [&](int pos, int slot, float2 v) {
                            Tr[pos] = cadd(ya[slot], split_twiddle<PWp::R0, true>(v, wt, slot % PWp::R0));
                        });
                } else
                    inv_stage_tw<PWp, 0>(tau, tw,
                        [&](int pos, int) { return Xr[pos]; },
                        [&](int pos, int slot, float2 v) {
                            Tr[pos] = cadd(ya[slot], TwP::mulc(v, pos * (kTwN / L)));
                        });
            }
        } else {
            constexpr int s = PH - kW1;
            const int tau = line_thread<CB>(tid);
            LCT_UNROLL
            for (int hb = 0; hb < nHB; ++hb) {
                const int col = hb * CB + tid % CB;
                float2* dst = p.s1 + plane * N * N + col;
                if constexpr (s == 0) {
                    inv_stage<PHp, 1, false, TwP>(tau,
                        [&](int pos, int) { return T[pos * RS + col]; },
                        [&](int pos, int, float2 v) { T[pos * RS + col] = v; });
                } else {
                    inv_stage<PHp, 0, true, TwP>(tau,
                        [&](int pos, int) { return T[pos * RS + col]; },
                        [&](int pos, int, float2 v) { dst[(size_t)pos * N] = v; });
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------
// generic kernel driver
// ---------------------------------------------------------------------------
template <class K, class = void> struct has_prologue { static constexpr bool value = false; };
template <class K> struct has_prologue<K, decltype((void)K::kHasPrologue)> { static constexpr bool value = true; };

template <class K, class = void> struct prologue_is_empty { static constexpr bool value = false; };
template <class K> struct prologue_is_empty<K, std::enable_if_t<(K::TwS::kBytes == 0)>> { static constexpr bool value = true; };

template <class K, class = void> struct has_preload { static constexpr bool value = false; };
template <class K> struct has_preload<K, decltype((void)K::kPreload)> { static constexpr bool value = true; };

template <class K, class = void> struct has_single_pass { static constexpr bool value = false; };
template <class K> struct has_single_pass<K, std::enable_if_t<K::kSinglePass>> { static constexpr bool value = true; };

template <class K, class = void> struct has_tile_walk { static constexpr bool value = false; };
template <class K> struct has_tile_walk<K, decltype((void)K::kTileWalks)> { static constexpr bool value = true; };

template <class K, class = void> struct has_group_sync { static constexpr bool value = false; };
template <class K> struct has_group_sync<K, decltype((void)K::kGroupSync)> { static constexpr bool value = true; };

#ifndef LCT_EMULATE
template <class K, int PH> struct PhaseLoop {
    static LCT_DEV void run(const Params& p, typename K::Regs& r, unsigned char* smem, int it) {
        K::template phase<PH>(p, r, smem, threadIdx.x, blockIdx.x, blockIdx.y, it);
        if constexpr (K::kWarpSync) __syncwarp();
        else if constexpr (has_group_sync<K>::value) {
            if constexpr (K::group_phase(PH))
                if (K::group_of(threadIdx.x)) asm volatile("bar.sync 2, %0;" ::"n"(K::kGroupThreads) : "memory");
                else asm volatile("bar.sync 1, %0;" ::"n"(K::kGroupThreads) : "memory");
            else __syncthreads();
        }
        else __syncthreads();
        if constexpr (PH + 1 < K::kPhases) PhaseLoop<K, PH + 1>::run(p, r, smem, it);
    }
};

template <class K> __global__ void __launch_bounds__(K::kThreads, K::kMinBlocks) lct_kernel(const __grid_constant__ Params p, const int iters) {
    extern __shared__ __align__(128) unsigned char smem[];
    typename K::Regs r;
    // the prologue only touches plan constants (twiddle and operator tables) and shared memory: it may run while the
    // previous kernel of the stream is still draining
    if constexpr (has_prologue<K>::value) K::prologue(p, r, smem, threadIdx.x, blockIdx.x, blockIdx.y);
    asm volatile("griddepcontrol.wait;" ::: "memory");                   // no-op for a plain launch
    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (has_preload<K>::value) K::preload(p, r, smem, threadIdx.x, blockIdx.x, blockIdx.y);    // loads fly across the barrier
#ifndef LCT_SKIP_EMPTY_PROLOGUE_BARRIER
#define LCT_SKIP_EMPTY_PROLOGUE_BARRIER 1
#endif
    // kernels whose prologue fills no table (constant-bank twiddles) have nothing to publish before phase 0; whatever
    // else the prologue initialises is first used behind a later phase barrier
    if constexpr (has_prologue<K>::value && !(LCT_SKIP_EMPTY_PROLOGUE_BARRIER && prologue_is_empty<K>::value)) __syncthreads();
    if constexpr (has_single_pass<K>::value) {
        PhaseLoop<K, 0>::run(p, r, smem, 0);                 // no loop: nothing in Regs outlives its last use
    } else {
        for (int it = 0; it < iters; ++it) {
            if constexpr (has_tile_walk<K>::value)
                if (!K::walk_active(p, blockIdx.x, blockIdx.y, it)) break;
            PhaseLoop<K, 0>::run(p, r, smem, it);
        }
    }
}
#endif

}  // namespace lct
