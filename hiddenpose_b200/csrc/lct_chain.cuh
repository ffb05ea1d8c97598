// Plan selection and the five-kernel chain, shared by the CUDA library
// (lct_api.cu) and by the CPU thread emulator used in tests (tests/emu).
#pragma once

#include "lct_kernels.cuh"

namespace lct {

// Column-batched plans (lanes run across columns; used along T with L = M and
// along H with L = 2N).
template <int L> struct ColPlan;
template <> struct ColPlan<8>   { using type = Plan<4, 2>; };
template <> struct ColPlan<16>  { using type = Plan<4, 4>; };
template <> struct ColPlan<32>  { using type = Plan<8, 4>; };
template <> struct ColPlan<64>  { using type = Plan<8, 8>; };
template <> struct ColPlan<128> { using type = Plan<16, 8>; };
template <> struct ColPlan<256> { using type = Plan<16, 16>; };
template <> struct ColPlan<512> { using type = Plan<16, 32>; };    // measured on cfg3: (8,8,8) 590 us, (16,32) 355-447 us

// The two time kernels may prefer different factorizations of the same length.
template <int M> struct TimeFwdPlan { using type = typename ColPlan<M>::type; };
template <int M> struct TimeInvPlan { using type = typename ColPlan<M>::type; };
// cfg3, one 512-thread block of 32 columns per SM: K1 395 us, K5 337 us; (16,16,2) / (16,32) on
// 16-column tiles at two blocks per SM: 419 / 358 us
template <> struct TimeFwdPlan<512> { using type = Plan<32, 16>; };
template <> struct TimeInvPlan<512> { using type = Plan<32, 16>; };
template <int L> struct RowFwdPlan { using type = typename ColPlan<L>::type; };
template <int L> struct RowInvPlan { using type = typename ColPlan<L>::type; };
template <> struct RowFwdPlan<512> { using type = Plan<32, 16>; };         // cfg5: 248 us; (8,8,8) 385 us, (16,32) 412 us

// Line plans for K3 (lanes run along the contiguous line): two stages only.
template <int L> struct LinePlan;
template <> struct LinePlan<16>  { using type = Plan<4, 4>; };
template <> struct LinePlan<32>  { using type = Plan<8, 4>; };
template <> struct LinePlan<64>  { using type = Plan<8, 8>; };
template <> struct LinePlan<128> { using type = Plan<16, 8>; };
template <> struct LinePlan<256> { using type = Plan<16, 16>; };
template <> struct LinePlan<512> { using type = Plan<32, 16>; };        // (16,32) measured the same

// column tile (in columns of the flattened H*W axis) for the T-axis kernels
#ifndef LCT_TIME_TILE_512
#define LCT_TIME_TILE_512 32
#endif
template <int M> struct TimeTile { static constexpr int CT = (M >= 512) ? LCT_TIME_TILE_512 : 32; };
inline int time_tile_columns(int M) { return M >= 512 ? LCT_TIME_TILE_512 : 32; }
// column tile along W for the H-axis kernels
template <int N> struct RowTile { static constexpr int CT = (N >= 256) ? 16 : (N < 32 ? N : 32); };
// rows per block for K3
template <int N> struct LineRows {
    using P = typename LinePlan<2 * N>::type;
    static constexpr int RB = (256 / P::TL) < 2 * N ? (256 / P::TL) : 2 * N;
};

// plane-resident fusion of K2+K3+K4: the 2N x (N+1) c64 plane must fit in shared memory
constexpr bool plane_fusable(int N) { return N <= 64; }
constexpr int kPlaneThreads = 512;          // 256-thread blocks (3 per SM) measured slower
template <int N> struct PlaneKernel {
    using PHp = typename ColPlan<2 * N>::type;
    static constexpr int kFull = N * PHp::TL;                               // one column batch
    static constexpr int NT = kFull < kPlaneThreads ? kFull : kPlaneThreads;
    // persistent walk with the next plane staged by the bulk-copy unit: at N = 64 (measured); the smaller planes
    // are a few KB and their kernels are launch-bound
    using type = PlaneFilter<PHp, typename ColPlan<N>::type, NT, (LCT_PLANE_PERSIST && N == 64)>;
};
// H-frequency held by plane row r after the forward H stages (the fused filter is stored in this order)
template <int N> int plane_row_freq(int r) { return ColPlan<2 * N>::type::pos_to_freq(r); }

constexpr bool supported_M(int M) { return M == 32 || M == 64 || M == 128 || M == 256 || M == 512; }
constexpr bool supported_N(int N) { return N == 8 || N == 16 || N == 32 || N == 64 || N == 128 || N == 256; }

enum ChainStage { kStageTimeFwd = 1, kStageRowFwd = 2, kStageColFilter = 4, kStageRowInv = 8, kStageTimeInv = 16, kStageAll = 31 };

template <int M, class Launcher> int launch_time_fwd(const Params& p, Launcher& l) {
    using P = typename TimeFwdPlan<M>::type;
    if constexpr (P::E >= 32)           // one block per SM: persistent blocks with the next tile copied ahead
        return l.template launch<TimeFwdPersistent<P, TimeTile<M>::CT>>(p);
    else
        return l.template launch<TimeFwd<P, TimeTile<M>::CT>>(p);
}
// (M = 512 as two 256-point inverses by input parity, side by side in a 1024-thread block like the H-axis kernels
//  at N = 256, measured slower: 268 vs 230 us at cfg3, 143 vs 124 us at cfg5)
template <int M, class Launcher> int launch_time_inv(const Params& p, Launcher& l) {
    return l.template launch<TimeInv<typename TimeInvPlan<M>::type, TimeTile<M>::CT>>(p);
}
// 512-point columns (N = 256) run as two 256-point transforms by parity (RowFwdSplit / RowInvSplit)
#ifndef LCT_ROW_SPLIT
#define LCT_ROW_SPLIT 1
#endif
template <int N, class Launcher> int launch_row_fwd(const Params& p, Launcher& l) {
    if constexpr (N >= 256 && LCT_ROW_SPLIT)
        return l.template launch<RowFwdSplit<typename ColPlan<N>::type, 16>>(p);
    else
        return l.template launch<RowFwd<typename RowFwdPlan<2 * N>::type, RowTile<N>::CT>>(p);
}
template <int N, class Launcher> int launch_row_inv(const Params& p, Launcher& l) {
    if constexpr (N >= 256 && LCT_ROW_SPLIT)
        return l.template launch<RowInvSplit<typename ColPlan<N>::type, 16>>(p);
    else
        return l.template launch<RowInv<typename RowInvPlan<2 * N>::type, RowTile<N>::CT>>(p);
}
template <int N, class Launcher> int launch_col_filter(const Params& p, Launcher& l) {
    // 512-point lines: two 256-point transforms by output parity (16-wide butterflies): 564 vs 682 us at cfg5, 5 % ahead
    // with four channels.  At N = 128 the register-cached filter of ColFilter wins clearly (288 vs 460 us at cfg4).
    // (The two parities side by side in one warp, as the H-axis kernels do it, measured slower here: 673-795 vs 560 us.)
    if constexpr (N >= 256) {
        using PL = typename LinePlan<N>::type;
        if (p.filt_sym) return l.template launch<ColFilterSplit<PL, 256 / PL::TL, true>>(p);
        return l.template launch<ColFilterSplit<PL, 256 / PL::TL, false>>(p);
    }
    else {
        if (p.filt_sym) return l.template launch<ColFilter<typename LinePlan<2 * N>::type, LineRows<N>::RB, true>>(p);
        return l.template launch<ColFilter<typename LinePlan<2 * N>::type, LineRows<N>::RB, false>>(p);
    }
}
template <int N, class Launcher> int launch_plane(const Params& p, Launcher& l) {
    if constexpr (plane_fusable(N)) return l.template launch<typename PlaneKernel<N>::type>(p);
    else return -1;
}

#define LCT_SWITCH_M(M, CALL)                                   \
    switch (M) {                                                \
        case 32:  { constexpr int kM = 32;  rc = CALL; } break; \
        case 64:  { constexpr int kM = 64;  rc = CALL; } break; \
        case 128: { constexpr int kM = 128; rc = CALL; } break; \
        case 256: { constexpr int kM = 256; rc = CALL; } break; \
        case 512: { constexpr int kM = 512; rc = CALL; } break; \
        default: rc = -1;                                       \
    }
#define LCT_SWITCH_N(N, CALL)                                   \
    switch (N) {                                                \
        case 8:   { constexpr int kN = 8;   rc = CALL; } break; \
        case 16:  { constexpr int kN = 16;  rc = CALL; } break; \
        case 32:  { constexpr int kN = 32;  rc = CALL; } break; \
        case 64:  { constexpr int kN = 64;  rc = CALL; } break; \
        case 128: { constexpr int kN = 128; rc = CALL; } break; \
        case 256: { constexpr int kN = 256; rc = CALL; } break; \
        default: rc = -1;                                       \
    }

// Rows of mtx below this index may carry more than three entries: the time-forward kernel looks for a
// CSR tail only at the first input of each stage-0 butterfly (positions < st(0), i.e. rows < 2 st(0)).
template <int M> constexpr int time_tail_rows_for() { return 2 * TimeFwdPlan<M>::type::st(0); }
inline int time_tail_rows(int M) {
    int rc = 0;
    LCT_SWITCH_M(M, (time_tail_rows_for<kM>()));
    return rc < 0 ? 0 : rc;
}
// Pairs (2p, 2p+1) of mtx rows from this index on are applied through pair records (at most three columns together).
template <int M> constexpr int time_long_pairs_for() { return TimeLongQ<M>::Q * TimeFwdPlan<M>::type::st(0); }
inline int time_long_pairs(int M) {
    int rc = 0;
    LCT_SWITCH_M(M, (time_long_pairs_for<kM>()));
    return rc < 0 ? 0 : rc;
}

// Resampling-operator tables the chain needs (device pointers on the GPU; see lct_tables.h).
struct BandTable { const float4* ell; const int* rowptr; const float* vals; const float4* pair; };
struct ChainTables {
    BandTable mtx_falloff;      // mtx[i][j] * falloff[j]     forward  K1
    BandTable mtx;              // mtx[i][j]                  backward K1
    BandTable mtxi;             // mtxi[j][i]                 forward  K5
    BandTable mtxi_falloff;     // mtxi[j][i] * falloff[j]    backward K5
    const float2* filt;         // (M+1, 2N, 2N) natural order -- or its (M+1, N+1, N+1) quarter -- for K3 (null if fused)
    const float2* filt_plane;   // (M+1, 2N kw, 2N plane rows), for PlaneFilter   (null if not fused)
    int filt_sym;               // FilterLayout of filt (Params::filt_sym)
};

// Runs the stages selected in `mask` (all five for a real call).
//   forward : in = x (C,Tin,N,N) placed at [be, be+Tin);   out = y  (C,M,N,N)
//   backward: in = gy (C,M,N,N) placed at [0, M);          out = gx (C,Tin,N,N) = rows [be, be+Tin)
// `l.mark(i)` is called before stage i (i = 0..4) and once more (i = 5) at the end.
template <class Launcher>
int run_chain(Launcher& l, const ChainTables& t, int M, int N, int C, int D, int Tin,
              int be_uniform, const int* be_dev, int c_base, const float* in, float* out,
              float2* s1, float2* s2, bool backward, int mask = kStageAll,
              unsigned long long* minmax_keys = nullptr) {
    Params p{};
    p.M = M; p.N = N; p.C = C; p.D = D;
    p.s1 = s1; p.s2 = s2; p.filt = t.filt; p.conj_filter = backward ? 1 : 0; p.filt_sym = t.filt_sym;
    p.in = in; p.out = out; p.c_base = c_base;
    int rc = 0;
    l.mark(0);
    if (mask & kStageTimeFwd) {
        p.in_T = backward ? M : Tin;
        p.be_uniform = backward ? 0 : be_uniform;
        p.be_dev = backward ? nullptr : be_dev;
        const BandTable& bt = backward ? t.mtx : t.mtx_falloff;
        p.ell = bt.ell; p.rowptr = bt.rowptr; p.vals = bt.vals; p.pair = bt.pair;
        LCT_SWITCH_M(M, (launch_time_fwd<kM>(p, l)));
        if (rc) return rc;
    }
    l.mark(1);
    constexpr int kMiddle = kStageRowFwd | kStageColFilter | kStageRowInv;
    if ((mask & kMiddle) == kMiddle && t.filt_plane) {
        p.filt = t.filt_plane;
        LCT_SWITCH_N(N, (launch_plane<kN>(p, l)));
        if (rc) return rc;
        l.mark(2); l.mark(3);
        mask &= ~kMiddle;
    }
    if (mask & kStageRowFwd) {
        LCT_SWITCH_N(N, (launch_row_fwd<kN>(p, l)));
        if (rc) return rc;
    }
    l.mark(2);
    if (mask & kStageColFilter) {
        LCT_SWITCH_N(N, (launch_col_filter<kN>(p, l)));
        if (rc) return rc;
    }
    l.mark(3);
    if (mask & kStageRowInv) {
        LCT_SWITCH_N(N, (launch_row_inv<kN>(p, l)));
        if (rc) return rc;
    }
    l.mark(4);
    if (mask & kStageTimeInv) {
        p.out_T = backward ? Tin : M;
        p.be_uniform = backward ? be_uniform : 0;
        p.be_dev = backward ? be_dev : nullptr;
        const BandTable& bt = backward ? t.mtxi_falloff : t.mtxi;
        p.minmax_keys = backward ? nullptr : minmax_keys;
        p.ell = bt.ell; p.rowptr = bt.rowptr; p.vals = bt.vals;
        LCT_SWITCH_M(M, (launch_time_inv<kM>(p, l)));
        if (rc) return rc;
    }
    l.mark(5);
    return 0;
}

}  // namespace lct
