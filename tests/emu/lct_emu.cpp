// CPU thread emulator for the LCT kernels -- TEST INFRASTRUCTURE ONLY.
//
// Compiles hiddenpose_b200/csrc/lct_kernels.cuh with LCT_EMULATE and steps every
// block phase by phase, thread by thread, with the same index math the GPU runs.
// It exists because the build container has no GPU: kernel logic is debugged here,
// then confirmed on a B200.  It is never loaded by the hiddenpose_b200 package.
#define LCT_EMULATE 1
#include <cmath>
#include <cstring>
#include <array>
#include <limits>
#include <utility>
#include <vector>

#include "lct_chain.cuh"
#include "lct_tables.h"

namespace lct {
float2 h_tw[kTwN];
}

namespace {

bool g_reverse_threads = false;   // run threads in reverse order inside a phase: exposes intra-phase races

template <class K, int PH> struct EmuPhases {
    static void run(const lct::Params& p, std::vector<typename K::Regs>& regs, unsigned char* smem, int bx, int by, int it) {
        if (!g_reverse_threads)
            for (int tid = 0; tid < K::kThreads; ++tid) K::template phase<PH>(p, regs[tid], smem, tid, bx, by, it);
        else
            for (int tid = K::kThreads - 1; tid >= 0; --tid) K::template phase<PH>(p, regs[tid], smem, tid, bx, by, it);
        if constexpr (PH + 1 < K::kPhases) EmuPhases<K, PH + 1>::run(p, regs, smem, bx, by, it);
    }
};

// Set-barrier drift: kernels that synchronise warp sets on named barriers between some phases (K::group_phase)
// claim that the sets are independent until the next whole-block barrier.  In drift mode one set runs through the
// whole run of such phases before the other starts -- the largest skew the hardware could produce.
int g_drift = 0;                  // 0 = off, 1 = set 0 first, 2 = set 1 first

template <class K, int PH> void run_phase_of_set(const lct::Params& p, std::vector<typename K::Regs>& regs, unsigned char* smem,
                                                 int bx, int by, int it, int set) {
    for (int i = 0; i < K::kThreads; ++i) {
        const int tid = g_reverse_threads ? K::kThreads - 1 - i : i;
        if constexpr (lct::has_group_sync<K>::value) {
            if (set >= 0 && K::group_of(tid) != set) continue;
        }
        K::template phase<PH>(p, regs[tid], smem, tid, bx, by, it);
    }
}
template <class K> using PhaseFn = void (*)(const lct::Params&, std::vector<typename K::Regs>&, unsigned char*, int, int, int, int);
template <class K, int... I> std::array<PhaseFn<K>, sizeof...(I)> phase_table(std::integer_sequence<int, I...>) {
    return {{&run_phase_of_set<K, I>...}};
}
template <class K> void run_phases_drifting(const lct::Params& p, std::vector<typename K::Regs>& regs, unsigned char* smem,
                                            int bx, int by, int it) {
    if constexpr (lct::has_group_sync<K>::value) {
        const auto table = phase_table<K>(std::make_integer_sequence<int, K::kPhases>{});
        int ph = 0;
        while (ph < K::kPhases) {
            if (!K::group_phase(ph)) { table[ph](p, regs, smem, bx, by, it, -1); ++ph; continue; }
            int e = ph;
            while (e < K::kPhases - 1 && K::group_phase(e)) ++e;          // phase e ends with a whole-block barrier
            for (int k = 0; k < 2; ++k) {
                const int set = g_drift == 1 ? k : 1 - k;
                for (int q = ph; q <= e; ++q) table[q](p, regs, smem, bx, by, it, set);
            }
            ph = e + 1;
        }
    }
}

struct EmuLauncher {
    void mark(int) {}
    template <class K> int launch(const lct::Params& p0) {
        lct::Params p = p0;
        p.ahead = 3;                                         // three "resident" blocks: persistent kernels walk several tiles each
        int gx, gy;
        K::grid(p, gx, gy);
        const int iters = K::iterations(p);
        std::vector<unsigned char> smem(K::kSmem + 64);     // 64 guard bytes: a write past kSmem trips the canary
        std::vector<typename K::Regs> regs(K::kThreads);
        for (int by = 0; by < gy; ++by)
            for (int bx = 0; bx < gx; ++bx) {
                // poison shared memory so reads of never-written slots surface as NaN
                float* f = reinterpret_cast<float*>(smem.data());
                for (size_t i = 0; i < smem.size() / 4; ++i) f[i] = std::numeric_limits<float>::quiet_NaN();
                if constexpr (lct::has_prologue<K>::value) {
                    if (!g_reverse_threads)
                        for (int tid = 0; tid < K::kThreads; ++tid) K::prologue(p, regs[tid], smem.data(), tid, bx, by);
                    else
                        for (int tid = K::kThreads - 1; tid >= 0; --tid) K::prologue(p, regs[tid], smem.data(), tid, bx, by);
                }
                if constexpr (lct::has_preload<K>::value)
                    for (int tid = 0; tid < K::kThreads; ++tid) K::preload(p, regs[tid], smem.data(), tid, bx, by);
                std::memset(smem.data() + K::kSmem, 0xA5, 64);
                for (int it = 0; it < iters; ++it) {
                    if (g_drift && lct::has_group_sync<K>::value) run_phases_drifting<K>(p, regs, smem.data(), bx, by, it);
                    else EmuPhases<K, 0>::run(p, regs, smem.data(), bx, by, it);
                }
                for (int g = 0; g < 64; ++g)
                    if (smem[K::kSmem + g] != 0xA5) return 200;      // shared-memory overrun
            }
        return 0;
    }
};

}  // namespace

extern "C" {

void lct_emu_init(int reverse_threads) {
    g_reverse_threads = reverse_threads != 0;
    for (int j = 0; j < lct::kTwN; ++j) {
        const double a = -2.0 * M_PI * j / lct::kTwN;
        lct::h_tw[j].x = (float)std::cos(a);
        lct::h_tw[j].y = (float)std::sin(a);
    }
}

void lct_emu_set_drift(int mode) { g_drift = mode; }

// All pointers are host pointers.  `filt` must already carry the 1/(8 M N N) scale; with filt_sym != 0 it is the
// quarter layout (M+1, N+1, N+1) of Params::filt_sym.  The operator
// arrives as the CSR of mtx plus the falloff vector, exactly as lct_plan_create receives it.
int lct_emu_run(int M, int N, int C, int D, int Tin, int be_uniform, const int* be,
                const float* in, float* out, float* s1, float* s2,
                const int* mtx_rowptr, const int* mtx_colidx, const float* mtx_vals, const float* falloff,
                const float* filt, const float* filt_plane, int backward, int mask, int filt_sym) {
    lct::HostTables ht;
    if (!lct::build_tables(M, mtx_rowptr, mtx_colidx, mtx_vals, falloff, lct::time_tail_rows(M), ht, lct::time_long_pairs(M), lct::time_tile_columns(M)).empty()) return 100;
    auto band = [](const std::vector<lct::EllRow>& e, const std::vector<int32_t>& rp, const std::vector<float>& v,
                   const std::vector<lct::PairRow>* pr = nullptr) {
        return lct::BandTable{reinterpret_cast<const float4*>(e.data()), rp.data(), v.data(),
                              pr ? reinterpret_cast<const float4*>(pr->data()) : nullptr};
    };
    lct::ChainTables t{band(ht.mtx_ell_falloff, ht.mtx_rowptr, ht.mtx_vals_falloff, &ht.mtx_pair_falloff),
                       band(ht.mtx_ell, ht.mtx_rowptr, ht.mtx_vals, &ht.mtx_pair),
                       band(ht.mtxi_ell, ht.mtxi_rowptr, ht.mtxi_vals), band(ht.mtxi_ell_falloff, ht.mtxi_rowptr, ht.mtxi_vals_falloff),
                       reinterpret_cast<const float2*>(filt), reinterpret_cast<const float2*>(filt_plane), filt_sym};
    EmuLauncher l;
    return lct::run_chain(l, t, M, N, C, D, Tin, be_uniform, be, 0, in, out,
                          reinterpret_cast<float2*>(s1), reinterpret_cast<float2*>(s2), backward != 0, mask);
}

// plane row -> H frequency map used to lay out the fused filter (same function lct_api.cu uses)
int lct_emu_plane_row_freq(int N, int r) {
    int rc = -1;
    LCT_SWITCH_N(N, (lct::plane_row_freq<kN>(r)));
    return rc;
}

}  // extern "C"
