// Host-side construction of the resampling tables the kernels read.
// Shared by the CUDA library (lct_api.cu) and the CPU thread emulator (tests/emu).
//
// The reference applies mtx / mtxi = mtx^T as dense M x M matmuls
// (/root/reference/models/tflct.py:135-138,156-159).  Both are staircase band matrices
// (utils/helper.py:35-69): every row is one contiguous run of columns, nearly always of
// length <= 3.  Each row is therefore stored as one 16-byte record
//     { start * kEllStride,  w0, w1, w2 }
// read with a single 128-bit load and applied branch-free (kEllStride = columns per time tile,
// so the first field is directly the element offset into the tile).  Only the few longer rows
// (the first ~sqrt(M) rows of mtx) continue into the CSR arrays; build_tables checks that they
// all lie below `tail_rows`, the only rows for which the time-forward kernel looks for a tail,
// and that mtxi = mtx^T has none.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace lct {

constexpr int kEllStride = 32;                        // default tile width; build_tables takes the kernel's own (TimeTile<M>::CT)

struct EllRow { int32_t offset; float w[3]; };       // 16 bytes, loaded as float4
// The time-forward kernel consumes rows in pairs (2p, 2p+1) -- the real and imaginary input of the packed-real
// FFT -- and the two bands overlap: outside the first few pairs their union spans at most three columns.  Such a
// pair is one 32-byte record: the union's offset, row 2p's three weights, row 2p+1's three weights (zero where a
// row does not reach), so three tile loads feed six multiply-adds.
struct PairRow { int32_t offset; float a[3]; float b[3]; float pad; };     // 32 bytes, loaded as two float4

struct HostTables {
    int M = 0;
    // mtx rows (K1): band start/len + values with and without the falloff folded in
    std::vector<int32_t> mtx_rowptr;
    std::vector<float> mtx_vals_falloff, mtx_vals;
    std::vector<EllRow> mtx_ell_falloff, mtx_ell;
    std::vector<PairRow> mtx_pair_falloff, mtx_pair;
    // mtxi rows (K5)
    std::vector<int32_t> mtxi_rowptr;
    std::vector<float> mtxi_vals, mtxi_vals_falloff;
    std::vector<EllRow> mtxi_ell, mtxi_ell_falloff;
};

inline std::vector<EllRow> make_ell(int M, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& start,
                                    const std::vector<float>& vals, int stride = kEllStride) {
    std::vector<EllRow> ell(M);
    for (int i = 0; i < M; ++i) {
        const int len = rowptr[i + 1] - rowptr[i];
        // the low bits of the offset (free: it is a multiple of the tile width) carry the number of entries past the
        // third, saturating at stride - 1 (the kernel then takes the true length from the CSR row pointers)
        const int extra = len > 3 ? len - 3 : 0;
        ell[i].offset = (len > 0 ? start[i] : 0) * stride + (extra < stride - 1 ? extra : stride - 1);
        for (int e = 0; e < 3; ++e) ell[i].w[e] = (e < len) ? vals[rowptr[i] + e] : 0.0f;
    }
    return ell;
}

// Pair records for pairs >= long_pairs (earlier pairs are zero-filled: the kernel takes the row records for them).
inline std::vector<PairRow> make_pairs(int M, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& start,
                                       const std::vector<float>& vals, int long_pairs, int stride = kEllStride) {
    std::vector<PairRow> pr(M / 2);
    std::memset(pr.data(), 0, sizeof(PairRow) * pr.size());
    for (int p = long_pairs; p < M / 2; ++p) {
        const int r0 = 2 * p, r1 = 2 * p + 1;
        const int l0 = rowptr[r0 + 1] - rowptr[r0], l1 = rowptr[r1 + 1] - rowptr[r1];
        const int u = l0 > 0 ? (l1 > 0 ? (start[r0] < start[r1] ? start[r0] : start[r1]) : start[r0]) : (l1 > 0 ? start[r1] : 0);
        pr[p].offset = u * stride;
        for (int e = 0; e < l0; ++e) pr[p].a[start[r0] + e - u] = vals[rowptr[r0] + e];
        for (int e = 0; e < l1; ++e) pr[p].b[start[r1] + e - u] = vals[rowptr[r1] + e];
    }
    return pr;
}

// Returns "" on success, otherwise a description of what is wrong with the operator.
inline std::string build_tables(int M, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                                const float* falloff /* M or null */, int tail_rows, HostTables& t, int long_pairs = -1,
                                int stride = kEllStride /* columns per time tile: TimeTile<M>::CT */) {
    if (long_pairs < 0) long_pairs = M / 2;               // no pair records wanted
    t.M = M;
    if (rowptr[0] != 0) return "CSR row pointers must start at 0";
    const int nnz = rowptr[M];
    if (nnz <= 0) return "empty operator";
    std::vector<float> fall(M, 1.0f);
    if (falloff) std::memcpy(fall.data(), falloff, sizeof(float) * M);
    std::vector<int32_t> start(M, 0), t_start(M, M), t_count(M, 0), t_last(M, -1);
    for (int i = 0; i < M; ++i) {
        if (rowptr[i + 1] < rowptr[i]) return "CSR row pointers not monotone";
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int j = colidx[e];
            if (j < 0 || j >= M) return "CSR column index out of range";
            if (e == rowptr[i]) start[i] = j;
            else if (j != colidx[e - 1] + 1) return "operator rows must be contiguous column runs";
            if (t_count[j] == 0) t_start[j] = i;
            else if (i != t_last[j] + 1) return "operator columns must be contiguous row runs";
            t_last[j] = i;
            t_count[j]++;
        }
    }
    for (int i = 0; i < M; ++i) {
        if (rowptr[i + 1] - rowptr[i] > 3 && i >= tail_rows)
            return "operator rows with more than 3 entries must be among the first " + std::to_string(tail_rows) + " rows";
        if (t_count[i] > 3) return "operator columns must have at most 3 entries";
    }
    for (int p = long_pairs; p < M / 2; ++p) {
        const int r0 = 2 * p, r1 = 2 * p + 1;
        const int l0 = rowptr[r0 + 1] - rowptr[r0], l1 = rowptr[r1 + 1] - rowptr[r1];
        if (l0 == 0 || l1 == 0) continue;
        const int lo = start[r0] < start[r1] ? start[r0] : start[r1];
        const int e0 = start[r0] + l0, e1 = start[r1] + l1;
        if ((e0 > e1 ? e0 : e1) - lo > 3)
            return "operator row pairs (2p, 2p+1) beyond the first " + std::to_string(long_pairs) + " must span at most 3 columns together";
    }
    t.mtx_rowptr.assign(rowptr, rowptr + M + 1);
    t.mtx_vals.assign(vals, vals + nnz);
    t.mtx_vals_falloff.resize(nnz);
    for (int e = 0; e < nnz; ++e) t.mtx_vals_falloff[e] = vals[e] * fall[colidx[e]];   // x*gridz^p (tflct.py:123-127)
    t.mtx_ell = make_ell(M, t.mtx_rowptr, start, t.mtx_vals, stride);
    t.mtx_ell_falloff = make_ell(M, t.mtx_rowptr, start, t.mtx_vals_falloff, stride);
    t.mtx_pair = make_pairs(M, t.mtx_rowptr, start, t.mtx_vals, long_pairs, stride);
    t.mtx_pair_falloff = make_pairs(M, t.mtx_rowptr, start, t.mtx_vals_falloff, long_pairs, stride);

    // transpose: mtxi = mtx^T (helper.py:61)
    t.mtxi_rowptr.assign(M + 1, 0);
    for (int j = 0; j < M; ++j) t.mtxi_rowptr[j + 1] = t.mtxi_rowptr[j] + t_count[j];
    t.mtxi_vals.assign(nnz, 0.f);
    t.mtxi_vals_falloff.assign(nnz, 0.f);
    std::vector<int32_t> cursor(t.mtxi_rowptr.begin(), t.mtxi_rowptr.end() - 1);
    for (int i = 0; i < M; ++i)
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int j = colidx[e], dst = cursor[j]++;
            t.mtxi_vals[dst] = vals[e];
            t.mtxi_vals_falloff[dst] = vals[e] * fall[j];      // backward: falloff applied to the output bin
        }
    // the inverse gather computes a column's first row instead of reading it (band_dot_sq in lct_kernels.cuh)
    for (int j = 0; j < M; ++j)
        if (t_count[j] > 0 && t_start[j] != (int)(((long long)j * j) / M))
            return "operator column " + std::to_string(j) + " must start at row floor(j^2 / M) (the staircase of helper.py:35-69)";
    for (int j = 0; j < M; ++j) if (t_count[j] == 0) t_start[j] = (int)(((long long)j * j) / M);
    t.mtxi_ell = make_ell(M, t.mtxi_rowptr, t_start, t.mtxi_vals, stride);
    t.mtxi_ell_falloff = make_ell(M, t.mtxi_rowptr, t_start, t.mtxi_vals_falloff, stride);
    return "";
}

}  // namespace lct
