// Register-resident FFT building blocks for the LCT kernels (sm_100a).
//
// Everything here is header-only device code.  It also compiles as plain host
// C++ when LCT_EMULATE is defined; that build exists ONLY for tests/emu (a
// thread-by-thread CPU emulation of the kernels used to debug index math in a
// container without a GPU) and is never linked into the product library.
//
// Conventions: complex = float2 (re, im); forward DFT uses exp(-2*pi*i*k*n/L);
// "inverse" is the un-normalised conjugate transform (the 1/(8V) factor of the
// reference's torch.ifft, tflct.py:151, is folded into the filter table).
#pragma once

#include <stdint.h>

#ifdef LCT_EMULATE
#include <cmath>
#include <cstring>
#include <vector_types.h>
#define LCT_DEV inline
#define LCT_HD inline
#define LCT_UNROLL
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#else
#include <cuda_runtime.h>
#define LCT_DEV __device__ __forceinline__
#define LCT_HD __host__ __device__ __forceinline__
#define LCT_UNROLL _Pragma("unroll")
#endif

namespace lct {

constexpr int kTwN = 1024;   // size of the master twiddle table exp(-2*pi*i*j/1024)

// ---- complex helpers -------------------------------------------------------
// Blackwell packs two fp32 operations into one issue slot (FADD2 / FMUL2 / FFMA2, sm_100 only), with
// free operand swizzles (scalar broadcast, lo/hi swap).  A complex add or subtract is one instruction
// instead of two, a complex multiply three instead of four.  The kernels are bound by instruction
// issue inside the SM, not by DRAM, so this is where the slots are saved.
#if !defined(LCT_EMULATE) && !defined(LCT_NO_PACKED_FP32)
#define LCT_PACKED_FP32 1
LCT_DEV float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
LCT_DEV float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
// a*b = a * b.x + (-a.y*b.y, a.x*b.y): FMUL2 on the swapped pair with b.y broadcast, then FFMA2 with b.x
// broadcast and a per-half sign on the addend -- two instructions, all swizzles/signs are operand modifiers
LCT_DEV float2 cmul(float2 a, float2 b) {
    const float2 u = __fmul2_rn(make_float2(a.y, a.x), make_float2(b.y, b.y));
    return __ffma2_rn(a, make_float2(b.x, b.x), make_float2(-u.x, u.y));
}
// a * conj(b)
LCT_DEV float2 cmulc(float2 a, float2 b) {
    const float2 u = __fmul2_rn(make_float2(a.y, a.x), make_float2(b.y, b.y));
    return __ffma2_rn(a, make_float2(b.x, b.x), make_float2(u.x, -u.y));
}
LCT_DEV float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
#else
LCT_DEV float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LCT_DEV float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LCT_DEV float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
LCT_DEV float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
LCT_DEV float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
#endif
// scalar forms (2 FMUL + 2 FFMA), for code where the packed form's register pairing costs more than it saves
LCT_DEV float2 cmul_s(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
LCT_DEV float2 cmulc_s(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
LCT_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) / +i (inverse)
template <bool INV> LCT_DEV float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// cos/sin(2*pi*i/32), i in the first octant..quadrant; everything else by symmetry.
LCT_DEV float cos32_q(int i) {   // i in [0, 8]
    switch (i) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323043f;
        case 2: return 0.92387953251128674f;
        case 3: return 0.83146961230254524f;
        case 4: return 0.70710678118654752f;
        case 5: return 0.55557023301960218f;
        case 6: return 0.38268343236508978f;
        case 7: return 0.19509032201612825f;
        default: return 0.0f;
    }
}
// a * exp(-/+ 2*pi*i * idx/32); idx must be a compile-time constant after unrolling
// so that the branches below fold away.
template <bool INV> LCT_DEV float2 twmul32(float2 a, int idx) {
    idx &= 31;
    if (idx == 0) return a;
    if (idx == 16) return make_float2(-a.x, -a.y);
    if (idx == 8) return mul_mi<INV>(a);
    if (idx == 24) return mul_mi<!INV>(a);
    const int quad = idx >> 3, r = idx & 7;          // angle = quad*90deg + r*11.25deg
    float c, s;                                      // cos/sin of the forward angle magnitude
    if (r == 0) { c = 1.f; s = 0.f; } else { c = cos32_q(r); s = cos32_q(8 - r); }
    // rotate by quad quarter turns: (c,s) -> cos,sin of full angle
    float cc, ss;
    if (quad == 0) { cc = c; ss = s; }
    else if (quad == 1) { cc = -s; ss = c; }
    else if (quad == 2) { cc = -c; ss = -s; }
    else { cc = s; ss = -c; }
    // forward: multiply by (cc - i ss); inverse: (cc + i ss)
    if (r == 4) {   // |cc| == |ss| == 1/sqrt2 : one add network + one scale
        const float h = 0.70710678118654752f;
        const float sx = (cc > 0.f) ? 1.f : -1.f, sy = (ss > 0.f) ? 1.f : -1.f;
        // (x + i y) * (sx - i*sgn*sy) * h, sgn = +1 forward
        const float sg = INV ? -sy : sy;
        // real: sx*x + sg*y ; imag: sx*y - sg*x
#ifdef LCT_PACKED_FP32
        return __fmul2_rn(__fadd2_rn(make_float2(sx * a.x, sx * a.y), make_float2(sg * a.y, -sg * a.x)), make_float2(h, h));
#else
        return make_float2(h * (sx * a.x + sg * a.y), h * (sx * a.y - sg * a.x));
#endif
    }
    if (INV) ss = -ss;
#ifdef LCT_PACKED_FP32
    return cmul(a, make_float2(cc, -ss));            // a * (cc - i ss)
#else
    return make_float2(a.x * cc + a.y * ss, a.y * cc - a.x * ss);
#endif
}

// ---- in-register DFTs, natural-order output -------------------------------
template <int R, bool INV> LCT_DEV void dft(float2* a);

template <int R1, int R2, bool INV> LCT_DEV void dft_composite(float2* a) {
    // n = n2 + R2*n1, k = k1 + R1*k2:
    // X[k1 + R1 k2] = sum_n2 w_R2^(n2 k2) * w_R^(n2 k1) * sum_n1 w_R1^(n1 k1) a[n2 + R2 n1]
    constexpr int R = R1 * R2;
    float2 A[R1][R2];
    LCT_UNROLL
    for (int n2 = 0; n2 < R2; ++n2) {
        float2 col[R1];
        LCT_UNROLL
        for (int n1 = 0; n1 < R1; ++n1) col[n1] = a[n2 + R2 * n1];
        dft<R1, INV>(col);
        LCT_UNROLL
        for (int k1 = 0; k1 < R1; ++k1) A[k1][n2] = twmul32<INV>(col[k1], n2 * k1 * (32 / R));
    }
    LCT_UNROLL
    for (int k1 = 0; k1 < R1; ++k1) {
        float2 row[R2];
        LCT_UNROLL
        for (int n2 = 0; n2 < R2; ++n2) row[n2] = A[k1][n2];
        dft<R2, INV>(row);
        LCT_UNROLL
        for (int k2 = 0; k2 < R2; ++k2) a[k1 + R1 * k2] = row[k2];
    }
}

template <int R, bool INV> LCT_DEV void dft(float2* a) {
    static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "unsupported radix");
    if constexpr (R == 1) {
    } else if constexpr (R == 2) {
        const float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else if constexpr (R == 4) {
        const float2 t0 = cadd(a[0], a[2]), t1 = csub(a[0], a[2]);
        const float2 t2 = cadd(a[1], a[3]), t3 = mul_mi<INV>(csub(a[1], a[3]));
        a[0] = cadd(t0, t2); a[2] = csub(t0, t2);
        a[1] = cadd(t1, t3); a[3] = csub(t1, t3);
    } else if constexpr (R == 8) {
        dft_composite<2, 4, INV>(a);
    } else if constexpr (R == 16) {
        dft_composite<4, 4, INV>(a);
    } else {
        dft_composite<4, 8, INV>(a);
    }
}

// Forward-style transform whose inputs a[R/2..R) are known to be zero
// (zero-padded line): X[2k] = DFT_{R/2}(a)[k], X[2k+1] = DFT_{R/2}(a * w_R^n)[k].
template <int R, bool INV> LCT_DEV void dft_zero_upper(float2* a) {
    constexpr int H = R / 2;
    float2 e[H], o[H];
    LCT_UNROLL
    for (int n = 0; n < H; ++n) { e[n] = a[n]; o[n] = twmul32<INV>(a[n], n * (32 / R)); }
    dft<H, INV>(e);
    dft<H, INV>(o);
    LCT_UNROLL
    for (int k = 0; k < H; ++k) { a[2 * k] = e[k]; a[2 * k + 1] = o[k]; }
}

// Transform of which only outputs [0, R/2) are wanted (cropped line):
// y[n] = T_{R/2}(X_even)[n] + w_R^n * T_{R/2}(X_odd)[n], result left in a[0..R/2).
template <int R, bool INV> LCT_DEV void dft_lower_only(float2* a) {
    constexpr int H = R / 2;
    float2 e[H], o[H];
    LCT_UNROLL
    for (int k = 0; k < H; ++k) { e[k] = a[2 * k]; o[k] = a[2 * k + 1]; }
    dft<H, INV>(e);
    dft<H, INV>(o);
    LCT_UNROLL
    for (int n = 0; n < H; ++n) a[n] = cadd(e[n], twmul32<INV>(o[n], n * (32 / R)));
}

// ---- multi-stage plan over a line of length L = R0*R1*R2 --------------------
// In-place decimation-in-frequency: stage s works on sub-blocks of length
// Ls(s) = L/(R0..R(s-1)); a butterfly touches positions hi*Ls + q*st + lo with
// st = Ls/R(s).  After the last stage position p holds frequency pos_to_freq(p)
// (mixed-radix digit reversal); the inverse runs the adjoint stages in reverse.
template <int R0_, int R1_, int R2_ = 1> struct Plan {
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
    static constexpr int S = (R2 > 1) ? 3 : ((R1 > 1) ? 2 : 1);
    static constexpr int L = R0 * R1 * R2;
    static constexpr int E = (R0 > R1 ? (R0 > R2 ? R0 : R2) : (R1 > R2 ? R1 : R2));   // elements per thread
    static constexpr int TL = L / E;                                                  // threads per line
    static constexpr int radix(int s) { return s == 0 ? R0 : (s == 1 ? R1 : R2); }
    static constexpr int Ls(int s) { return s == 0 ? L : (s == 1 ? L / R0 : L / (R0 * R1)); }
    static constexpr int st(int s) { return Ls(s) / radix(s); }
    static constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }
    static constexpr int B0 = ilog2(R0), B1 = ilog2(R1), B2 = ilog2(R2);
    // positions and frequencies are the same digits in opposite order; all radices are powers of two,
    // so both maps are bit-field permutations (linear over disjoint bit sets)
    LCT_HD static int pos_to_freq(int p) {
        return (p >> (B1 + B2)) | (((p >> B2) & (R1 - 1)) << B0) | ((p & (R2 - 1)) << (B0 + B1));
    }
    LCT_HD static int freq_to_pos(int f) {
        return ((f & (R0 - 1)) << (B1 + B2)) | (((f >> B0) & (R1 - 1)) << B2) | (f >> (B0 + B1));
    }
    // frequency of the element a stage-s thread addresses as (pos, slot): the butterfly base is mapped
    // once at run time, the element's own digit (slot % radix, known at compile time) is a constant
    template <int s> LCT_HD static int freq_of(int pos, int slot) {
        const int off = (slot % radix(s)) * st(s);
        return pos_to_freq(pos - off) + pos_to_freq(off);
    }
};

// Callbacks: ld(pos, slot) -> float2 and stf(pos, slot, v), where pos is the line
// position and slot = m*r + q numbers the thread's elements (compile-time after
// unrolling, so callers may keep them in registers).
//
// Forward stage s for line-thread `tau` (of TL).
// ZERO_UPPER (stage 0 only): inputs with q >= r/2 are zero and are not loaded.
template <class P, int s, bool ZERO_UPPER, class TW, class LD, class ST>
LCT_DEV void fwd_stage(int tau, LD ld, ST stf) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    static_assert(NB % P::TL == 0, "threads per line must divide butterflies per stage");
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int hi = b / str, lo = b % str;
        const int base = hi * Ls + lo;
        float2 a[r];
        if constexpr (ZERO_UPPER) {
            LCT_UNROLL
            for (int q = 0; q < r / 2; ++q) a[q] = ld(base + q * str, m * r + q);
            dft_zero_upper<r, false>(a);
        } else {
            LCT_UNROLL
            for (int q = 0; q < r; ++q) a[q] = ld(base + q * str, m * r + q);
            dft<r, false>(a);
        }
        if constexpr (s < P::S - 1) {
            LCT_UNROLL
            for (int k = 1; k < r; ++k) a[k] = TW::template stage_mul<Ls, str>(a[k], k, lo);
        }
        LCT_UNROLL
        for (int k = 0; k < r; ++k) stf(base + k * str, m * r + k, a[k]);
    }
}

// Inverse (adjoint) stage s.  LOWER_ONLY (stage 0 only): only outputs q < r/2 are produced.
template <class P, int s, bool LOWER_ONLY, class TW, class LD, class ST>
LCT_DEV void inv_stage(int tau, LD ld, ST stf) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    static_assert(NB % P::TL == 0, "threads per line must divide butterflies per stage");
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int hi = b / str, lo = b % str;
        const int base = hi * Ls + lo;
        float2 a[r];
        LCT_UNROLL
        for (int k = 0; k < r; ++k) a[k] = ld(base + k * str, m * r + k);
        if constexpr (s < P::S - 1) {
            LCT_UNROLL
            for (int k = 1; k < r; ++k) a[k] = TW::template stage_mulc<Ls, str>(a[k], k, lo);
        }
        if constexpr (LOWER_ONLY) {
            dft_lower_only<r, true>(a);
            LCT_UNROLL
            for (int q = 0; q < r / 2; ++q) stf(base + q * str, m * r + q, a[q]);
        } else {
            dft<r, true>(a);
            LCT_UNROLL
            for (int q = 0; q < r; ++q) stf(base + q * str, m * r + q, a[q]);
        }
    }
}

// Variants of the two stage functions that take the inter-stage twiddles from a caller-held array
// (tw[m*(r-1) + k-1] = w_Ls^(k lo) for butterfly m, filled once by load_stage_twiddles) so that
// several transforms with the same thread mapping share one set of table reads.
template <class P, int s, class TW> LCT_DEV void load_stage_twiddles(int tau, float2* tw) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int lo = (tau + m * P::TL) % str;
        LCT_UNROLL
        for (int k = 1; k < r; ++k) tw[m * (r - 1) + k - 1] = TW::get((k * lo) * (kTwN / Ls));
    }
}
template <class P, int s> struct StageTwiddles { static constexpr int kCount = (P::L / P::radix(s) / P::TL) * (P::radix(s) - 1); };

template <class P, int s, class LD, class ST>
LCT_DEV void fwd_stage_tw(int tau, const float2* tw, LD ld, ST stf) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    static_assert(s < P::S - 1, "the last stage has no twiddles");
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int base = (b / str) * Ls + (b % str);
        float2 a[r];
        LCT_UNROLL
        for (int q = 0; q < r; ++q) a[q] = ld(base + q * str, m * r + q);
        dft<r, false>(a);
        LCT_UNROLL
        for (int k = 1; k < r; ++k) a[k] = cmul(a[k], tw[m * (r - 1) + k - 1]);
        LCT_UNROLL
        for (int k = 0; k < r; ++k) stf(base + k * str, m * r + k, a[k]);
    }
}
template <class P, int s, class LD, class ST>
LCT_DEV void inv_stage_tw(int tau, const float2* tw, LD ld, ST stf) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    static_assert(s < P::S - 1, "the last stage has no twiddles");
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int base = (b / str) * Ls + (b % str);
        float2 a[r];
        LCT_UNROLL
        for (int k = 0; k < r; ++k) a[k] = ld(base + k * str, m * r + k);
        LCT_UNROLL
        for (int k = 1; k < r; ++k) a[k] = cmulc(a[k], tw[m * (r - 1) + k - 1]);
        dft<r, true>(a);
        LCT_UNROLL
        for (int q = 0; q < r; ++q) stf(base + q * str, m * r + q, a[q]);
    }
}

// Visit (pos, slot) of every element this thread owns in stage s
// (all r per butterfly, or only the lower half).
template <class P, int s, class FN> LCT_DEV void for_each_slot(int tau, FN fn) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int base = (b / str) * Ls + (b % str);
        LCT_UNROLL
        for (int k = 0; k < r; ++k) fn(base + k * str, m * r + k);
    }
}
template <class P, int s, class FN> LCT_DEV void for_each_slot_lower(int tau, FN fn) {
    constexpr int r = P::radix(s), Ls = P::Ls(s), str = P::st(s), NB = P::L / r;
    LCT_UNROLL
    for (int m = 0; m < NB / P::TL; ++m) {
        const int b = tau + m * P::TL;
        const int base = (b / str) * Ls + (b % str);
        LCT_UNROLL
        for (int k = 0; k < r / 2; ++k) fn(base + k * str, m * r + k);
    }
}

}  // namespace lct
