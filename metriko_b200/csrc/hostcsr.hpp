// Host-side CSR helpers used at set-up time (transpose, SpGEMM, Hermitian products), templated on real / complex.
// Set-up runs once per matrix; the per-iteration work is all on the device.
#pragma once
#include <omp.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

namespace mk {

// NVTX ranges (header-only nvtx3: no-ops unless a profiler injected its library) around every stage a timeline should show: set-up stages,
// each PCG solve, each inverse-iteration step, each rounding re-solve (SURVEY.md section 5 "tracing")
#ifndef MK_NVTX_RANGE_DEFINED
#define MK_NVTX_RANGE_DEFINED
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#endif

// MK_VERBOSE=1: set-up stage timers on stderr; always: one NVTX range per timer scope, one mark per lap
struct StageTimer {
    bool on;
    double t0;
    const char* what;
    NvtxRange range;
    explicit StageTimer(const char* w) : on(getenv("MK_VERBOSE") != nullptr), t0(omp_get_wtime()), what(w), range(w) {}
    void lap(const char* stage) {
        nvtxMarkA(stage);
        if (!on) return;
        const double t = omp_get_wtime();
        fprintf(stderr, "[mk] %s: %-28s %8.1f ms\n", what, stage, (t - t0) * 1e3);
        t0 = t;
    }
};

template <class S>
struct HostCsr {
    int rows = 0, cols = 0;
    std::vector<int> ptr, col;
    std::vector<S> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};

template <class S> inline S hconj(S v) { return v; }
template <> inline std::complex<double> hconj(std::complex<double> v) { return std::conj(v); }
template <class S> inline double habs(S v) { return std::abs(v); }

template <class S>
inline HostCsr<S> transpose(const HostCsr<S>& A, bool conjugate) {
    HostCsr<S> B;
    B.rows = A.cols;
    B.cols = A.rows;
    B.ptr.assign(B.rows + 1, 0);
    for (int c : A.col) B.ptr[c + 1]++;
    for (int i = 0; i < B.rows; i++) B.ptr[i + 1] += B.ptr[i];
    B.col.resize(A.col.size());
    B.val.resize(A.val.size());
    std::vector<int> pos(B.ptr.begin(), B.ptr.end() - 1);
    for (int r = 0; r < A.rows; r++)
        for (int k = A.ptr[r]; k < A.ptr[r + 1]; k++) {
            int p = pos[A.col[k]]++;
            B.col[p] = r;
            B.val[p] = conjugate ? hconj(A.val[k]) : A.val[k];
        }
    return B;
}

// C = A * B, OpenMP over row blocks (two passes: count, fill), columns sorted within each row
template <class S>
inline HostCsr<S> matmul(const HostCsr<S>& A, const HostCsr<S>& B) {
    HostCsr<S> C;
    C.rows = A.rows;
    C.cols = B.cols;
    C.ptr.assign(A.rows + 1, 0);
    const int nt = omp_get_max_threads();
    std::vector<std::vector<int>> tcol(nt);
    std::vector<std::vector<S>> tval(nt);
    std::vector<int> tstart(nt + 1, 0);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((int64_t)A.rows * t / nt), r1 = (int)((int64_t)A.rows * (t + 1) / nt);
        std::vector<S> acc(B.cols, S(0));
        std::vector<int> mark(B.cols, -1), touched;
        auto& oc = tcol[t];
        auto& ov = tval[t];
        for (int i = r0; i < r1; i++) {
            touched.clear();
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) {
                const int j = A.col[k];
                const S a = A.val[k];
                for (int l = B.ptr[j]; l < B.ptr[j + 1]; l++) {
                    const int c = B.col[l];
                    if (mark[c] != i) { mark[c] = i; acc[c] = S(0); touched.push_back(c); }
                    acc[c] += a * B.val[l];
                }
            }
            std::sort(touched.begin(), touched.end());
            for (int c : touched) { oc.push_back(c); ov.push_back(acc[c]); }
            C.ptr[i + 1] = (int)touched.size();
        }
    }
    for (int i = 0; i < A.rows; i++) C.ptr[i + 1] += C.ptr[i];
    C.col.resize(C.ptr[A.rows]);
    C.val.resize(C.ptr[A.rows]);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((int64_t)A.rows * t / nt);
        std::copy(tcol[t].begin(), tcol[t].end(), C.col.begin() + C.ptr[r0]);
        std::copy(tval[t].begin(), tval[t].end(), C.val.begin() + C.ptr[r0]);
    }
    return C;
}

// Transpose of a CSR matrix given by raw arrays (int64 row pointers, as the C-ABI passes the large IGM inputs), OpenMP:
// per-thread column histograms give every entry its slot without atomics and keep rows in ascending order inside each
// column (deterministic).
template <class S>
inline HostCsr<S> transpose_raw(int rows, int cols, const int64_t* ptr, const int* col, const S* val) {
    HostCsr<S> B;
    B.rows = cols;
    B.cols = rows;
    const int64_t nnz = ptr[rows];
    B.ptr.assign((size_t)cols + 1, 0);
    B.col.resize((size_t)nnz);
    B.val.resize((size_t)nnz);
    const int nt = omp_get_max_threads();
    std::vector<std::vector<int>> hist(nt);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((int64_t)rows * t / nt), r1 = (int)((int64_t)rows * (t + 1) / nt);
        std::vector<int>& h = hist[t];
        h.assign(cols, 0);
        for (int64_t k = ptr[r0]; k < ptr[r1]; k++) h[col[k]]++;
#pragma omp barrier
        // exclusive scan over threads per column -> h[c] becomes this thread's offset inside column c; totals into B.ptr
#pragma omp for schedule(static)
        for (int c = 0; c < cols; c++) {
            int run = 0;
            for (int tt = 0; tt < nt; tt++) { const int v = hist[tt][c]; hist[tt][c] = run; run += v; }
            B.ptr[c + 1] = run;
        }
#pragma omp single
        for (int c = 0; c < cols; c++) B.ptr[c + 1] += B.ptr[c];
        for (int r = r0; r < r1; r++)
            for (int64_t k = ptr[r]; k < ptr[r + 1]; k++) {
                const int c = col[k];
                const int p = B.ptr[c] + h[c]++;
                B.col[p] = r;
                B.val[p] = val[k];
            }
    }
    return B;
}

// C = A * B with B given by raw CSR arrays (int64 row pointers); same scheme as matmul()
template <class S>
inline HostCsr<S> matmul_raw(const HostCsr<S>& A, int b_cols, const int64_t* bptr, const int* bcol, const S* bval) {
    HostCsr<S> C;
    C.rows = A.rows;
    C.cols = b_cols;
    C.ptr.assign(A.rows + 1, 0);
    const int nt = omp_get_max_threads();
    std::vector<std::vector<int>> tcol(nt);
    std::vector<std::vector<S>> tval(nt);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((int64_t)A.rows * t / nt), r1 = (int)((int64_t)A.rows * (t + 1) / nt);
        std::vector<S> acc(b_cols, S(0));
        std::vector<int> mark(b_cols, -1), touched;
        auto& oc = tcol[t];
        auto& ov = tval[t];
        for (int i = r0; i < r1; i++) {
            touched.clear();
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) {
                const int j = A.col[k];
                const S a = A.val[k];
                for (int64_t l = bptr[j]; l < bptr[j + 1]; l++) {
                    const int c = bcol[l];
                    if (mark[c] != i) { mark[c] = i; acc[c] = S(0); touched.push_back(c); }
                    acc[c] += a * bval[l];
                }
            }
            std::sort(touched.begin(), touched.end());
            for (int c : touched) { oc.push_back(c); ov.push_back(acc[c]); }
            C.ptr[i + 1] = (int)touched.size();
        }
    }
    for (int i = 0; i < A.rows; i++) C.ptr[i + 1] += C.ptr[i];
    C.col.resize(C.ptr[A.rows]);
    C.val.resize(C.ptr[A.rows]);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int r0 = (int)((int64_t)A.rows * t / nt);
        std::copy(tcol[t].begin(), tcol[t].end(), C.col.begin() + C.ptr[r0]);
        std::copy(tval[t].begin(), tval[t].end(), C.val.begin() + C.ptr[r0]);
    }
    return C;
}

template <class S>
inline void matvec(const HostCsr<S>& A, const std::vector<S>& x, std::vector<S>& y) {
    y.assign(A.rows, S(0));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < A.rows; i++) {
        S s(0);
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) s += A.val[k] * x[A.col[k]];
        y[i] = s;
    }
}

template <class S>
inline std::vector<S> diagonal(const HostCsr<S>& A) {
    std::vector<S> d(A.rows, S(0));
    for (int i = 0; i < A.rows; i++)
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (A.col[k] == i) d[i] += A.val[k];
    return d;
}

}  // namespace mk
