// The coarse tail of the multigrid V-cycle as ONE kernel: every level with at most MK_TAIL_ROWS rows (pre-smoothing, restriction, ...,
// the dense coarsest solve, ..., prolongation, post-smoothing) runs inside a single thread-block cluster, phase after phase, separated
// by the hardware cluster barrier instead of kernel boundaries.
//
// MEASURED NEGATIVE RESULT -- OFF BY DEFAULT (MK_TAIL_ROWS=32768 switches it on). The round-1 launch lists (profiles/launches_r01g_*)
// showed the levels below ~25k rows as 17-25 launches of 3-18 us under ncu, and a single persistent kernel looked like the fix. Timed
// live as graph replays on the B200 (profiles/cycle_probe_r02_*.json, 1M vertices) the separate launches cost far less than ncu's
// serialised figures suggest -- the whole cycle from the 22k-row level downwards (13 launches) takes 43 us for the face hierarchy and
// 70 us for the IGM one, 3.4 us per launch including its gap -- because each of them spreads over 148 SMs, while one cluster is
// confined to 16 SMs of one GPC: every phase is a chain of 3-4 dependent L2 round trips whichever way it is launched, and 16k threads
// need several such chains per phase where the full grid needs one. This kernel takes 105 us (face) / 172 us (IGM) for the same
// levels with 16 CTAs and 137 / 252 us with 8; restricted to the levels below 16k rows it still loses (27 against 22.5 us).
// It stays in the tree as the measured alternative; the defaults run one launch per phase.
//
// Layout: one cluster of 16 CTAs (8 if the device refuses the non-portable size) x 1024 threads on one GPC; vectors and operators stay
// in L2 (a few MB per level), every phase is a lanes-per-row CSR SpMV with the smoother update as its epilogue, exactly the forms of
// SpmvMode (csr.cuh). Rows much longer than the mean get a warp, rows above kTailHugeRow entries (the dense dofs of the IGM hierarchy,
// coupled to every aggregate along their seam) a whole CTA, before the regular rows. Vectors written in one phase and gathered in the
// next are accessed with ld.global.cg / st.global.cg (L2 only): the barrier orders the accesses, L2 is the point of coherence between
// the SMs of the cluster.
#pragma once
#include <cooperative_groups.h>

#include "csr.cuh"

namespace mk {

namespace cg = cooperative_groups;

constexpr int kTailMaxLevels = 10;
constexpr int kTailThreads = 1024;
constexpr int kTailHugeRow = 2048;   // rows longer than this are summed by a whole CTA
constexpr int kTailMedRowMin = 64;   // rows longer than max(this, 4 x the mean row length) get a warp of their own
constexpr int kTailRowsDefault = 0;  // off: see the header comment

// one CSR operand of the tail with its row classes: regular rows (<= cut entries: G lanes each), medium rows (a warp each), huge rows (a CTA each)
template <class T>
struct TailMat {
    const int *ptr, *col;
    const T* val;
    const int *med, *huge;
    int n_med, n_huge, cut, gmax;
};
template <class T>
struct TailLevel {
    int n, deg;
    TailMat<T> A, P, R;  // A_l ; P_l : n_l x n_{l+1} ; R_l : n_{l+1} x n_l
    const T* dinv;
    T *x, *xo, *b, *r, *d, *q;
    double inv_theta, delta, sigma;
};
template <class T>
struct TailDesc {
    int nl, coarse_n;
    const T* coarse_inv;
    TailLevel<T> lv[kTailMaxLevels];
};

template <class T> __device__ __forceinline__ T ld_cg(const T* p);
template <> __device__ __forceinline__ float ld_cg<float>(const float* p) { return __ldcg(p); }
template <> __device__ __forceinline__ double ld_cg<double>(const double* p) { return __ldcg(p); }
template <> __device__ __forceinline__ cfloat ld_cg<cfloat>(const cfloat* p) { const float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return make_cf(v.x, v.y); }
template <> __device__ __forceinline__ cdouble ld_cg<cdouble>(const cdouble* p) { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return make_c(v.x, v.y); }
template <class T> __device__ __forceinline__ void st_cg(T* p, T v);
template <> __device__ __forceinline__ void st_cg<float>(float* p, float v) { __stcg(p, v); }
template <> __device__ __forceinline__ void st_cg<double>(double* p, double v) { __stcg(p, v); }
template <> __device__ __forceinline__ void st_cg<cfloat>(cfloat* p, cfloat v) { __stcg(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
template <> __device__ __forceinline__ void st_cg<cdouble>(cdouble* p, cdouble v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

// y_row = sum_k val[k] x[col[k]] over the whole cluster, `epi(row, sum)` called once per row by one thread.
// The cluster has 16k threads and every load is an L2 round trip (~250 cycles), so the shape is chosen for memory-level parallelism, not
// for coalescing: as few lanes per row as keeps all threads busy (G = threads / rows, a thread per row on the 22k-row level -- the first
// version used 8-32 lanes per row there: 22 dependent rounds per thread and 200 us per cycle), four entries in flight per lane.
template <class T, class Epi>
__device__ __forceinline__ void tail_spmv(int n, const TailMat<T>& M, const T* x, int brank, int nblocks, double* red, Epi epi) {
    const int* __restrict__ ptr = M.ptr;
    const int* __restrict__ col = M.col;
    const T* __restrict__ val = M.val;
    for (int i = brank; i < M.n_huge; i += nblocks) {  // one CTA per huge row (uniform per CTA: block_sum synchronises)
        const int row = __ldg(M.huge + i);
        const int b = __ldg(ptr + row), en = __ldg(ptr + row + 1);
        T a0 = ST<T>::zero(), a1 = ST<T>::zero();
        int k = b + threadIdx.x;
        for (; k + kTailThreads < en; k += 2 * kTailThreads) {
            const int c0 = __ldg(col + k), c1 = __ldg(col + k + kTailThreads);
            a0 = ST<T>::fma_(ld_ro(val + k), ld_cg(x + c0), a0);
            a1 = ST<T>::fma_(ld_ro(val + k + kTailThreads), ld_cg(x + c1), a1);
        }
        if (k < en) a0 = ST<T>::fma_(ld_ro(val + k), ld_cg(x + __ldg(col + k)), a0);
        const T sum = block_sum_t(a0 + a1, red);
        if (threadIdx.x == 0) epi(row, sum);
    }
    {   // one warp per medium row
        const int lane = threadIdx.x & 31;
        const int nwarps = nblocks * (kTailThreads / 32);
        for (int i = brank * (kTailThreads / 32) + (threadIdx.x >> 5); i < M.n_med; i += nwarps) {
            const int row = __ldg(M.med + i);
            const int b = __ldg(ptr + row), en = __ldg(ptr + row + 1);
            T a0 = ST<T>::zero(), a1 = ST<T>::zero();
            int k = b + lane;
            for (; k + 96 < en; k += 128) {
                const int c0 = __ldg(col + k), c1 = __ldg(col + k + 32), c2 = __ldg(col + k + 64), c3 = __ldg(col + k + 96);
                const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + 32), v2 = ld_ro(val + k + 64), v3 = ld_ro(val + k + 96);
                const T x0 = ld_cg(x + c0), x1 = ld_cg(x + c1), x2 = ld_cg(x + c2), x3 = ld_cg(x + c3);
                a0 = ST<T>::fma_(v0, x0, a0);
                a1 = ST<T>::fma_(v1, x1, a1);
                a0 = ST<T>::fma_(v2, x2, a0);
                a1 = ST<T>::fma_(v3, x3, a1);
            }
            for (; k < en; k += 32) a0 = ST<T>::fma_(ld_ro(val + k), ld_cg(x + __ldg(col + k)), a0);
            const T sum = group_sum<T>(a0 + a1, 32);
            if (lane == 0) epi(row, sum);
        }
    }
    const int nthreads = nblocks * kTailThreads;
    int G = 1;
    while (G < M.gmax && 2 * G * n <= nthreads) G <<= 1;  // lanes per row: only as many as there are idle threads
    const int tid = brank * kTailThreads + threadIdx.x;
    const int sub = tid & (G - 1);
    const int ngroups = nthreads / G;
    const int nloop = (n + ngroups - 1) / ngroups;  // same trip count for every thread: the shuffles need whole warps
    int row = tid / G;
    for (int it = 0; it < nloop; it++, row += ngroups) {
        bool live = row < n;
        int b = 0, en = 0;
        if (live) {
            b = __ldg(ptr + row);
            en = __ldg(ptr + row + 1);
            if (en - b > M.cut) live = false;  // medium / huge row: served above
        }
        T a0 = ST<T>::zero(), a1 = ST<T>::zero();
        if (live) {
            int k = b + sub;
            for (; k + 3 * G < en; k += 4 * G) {
                const int c0 = __ldg(col + k), c1 = __ldg(col + k + G), c2 = __ldg(col + k + 2 * G), c3 = __ldg(col + k + 3 * G);
                const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + G), v2 = ld_ro(val + k + 2 * G), v3 = ld_ro(val + k + 3 * G);
                const T x0 = ld_cg(x + c0), x1 = ld_cg(x + c1), x2 = ld_cg(x + c2), x3 = ld_cg(x + c3);
                a0 = ST<T>::fma_(v0, x0, a0);
                a1 = ST<T>::fma_(v1, x1, a1);
                a0 = ST<T>::fma_(v2, x2, a0);
                a1 = ST<T>::fma_(v3, x3, a1);
            }
            for (; k < en; k += G) a0 = ST<T>::fma_(ld_ro(val + k), ld_cg(x + __ldg(col + k)), a0);
        }
        const T sum = group_sum<T>(a0 + a1, G);
        if (live && sub == 0) epi(row, sum);
    }
}

// x = Ainv b, one warp per row over the cluster
template <class T>
__device__ __forceinline__ void tail_dense(int n, const T* __restrict__ Ainv, const T* b, T* x, int brank, int nblocks) {
    const int lane = threadIdx.x & 31;
    const int nwarps = nblocks * (kTailThreads / 32);
    for (int row = brank * (kTailThreads / 32) + (threadIdx.x >> 5); row < n; row += nwarps) {
        T a0 = ST<T>::zero(), a1 = ST<T>::zero();
        int j = lane;
        for (; j + 32 < n; j += 64) {
            a0 = ST<T>::fma_(ld_ro(Ainv + (size_t)row * n + j), ld_cg(b + j), a0);
            a1 = ST<T>::fma_(ld_ro(Ainv + (size_t)row * n + j + 32), ld_cg(b + j + 32), a1);
        }
        if (j < n) a0 = ST<T>::fma_(ld_ro(Ainv + (size_t)row * n + j), ld_cg(b + j), a0);
        const T sum = group_sum<T>(a0 + a1, 32);
        if (lane == 0) st_cg(x + row, sum);
    }
}

// Levels [0, nl) of `D` (= levels [tail_start, last] of the hierarchy). On entry the parent's restriction has written lv[0].b and the
// start direction dinv b / theta into lv[0].x (degree 1) or lv[0].d; on exit xo_top holds the cycle's result for lv[0].
template <class T>
__global__ void __launch_bounds__(kTailThreads, 1) k_vcycle_tail(const __grid_constant__ TailDesc<T> D, T* xo_top, const PcgState* st) {
    __shared__ double red[32];
    if (st && st->done) return;  // uniform over the cluster
    cg::cluster_group cl = cg::this_cluster();
    const int brank = (int)cl.block_rank(), nb = (int)cl.num_blocks();
    const int m = D.nl;
    // ---- down: pre-smoothing (zero guess) + restriction
    for (int j = 0; j + 1 < m; j++) {
        const TailLevel<T>& L = D.lv[j];
        const TailLevel<T>& C = D.lv[j + 1];
        T* d_cur = L.deg == 1 ? L.x : L.d;
        double rho = 1.0 / L.sigma;
        for (int k = 0; k < L.deg; k++) {
            const double rho1 = 1.0 / (2.0 * L.sigma - rho);
            const double c1 = rho1 * rho, c2 = 2.0 * rho1 / L.delta;
            const bool first = k == 0, last = k == L.deg - 1;
            T* d_next = last ? nullptr : (d_cur == L.d ? L.q : L.d);
            T* xv = L.x;
            const T* dc = d_cur;
            tail_spmv<T>(L.n, L.A, dc, brank, nb, red, [&](int row, T acc) {
                const T di = ld_cg(dc + row);
                const T ri = (first ? ld_cg(L.b + row) : ld_cg(L.r + row)) - acc;
                if (!first) st_cg(xv + row, ld_cg(xv + row) + di);
                else if (xv != dc) st_cg(xv + row, di);
                st_cg(L.r + row, ri);
                if (d_next) st_cg(d_next + row, ST<T>::scale(c1, di) + ST<T>::scale(c2, ld_ro(L.dinv + row) * ri));
            });
            cl.sync();
            d_cur = d_next;
            rho = rho1;
        }
        const bool coarsest = j + 2 == m;
        T* cstart = C.deg == 1 ? C.x : C.d;
        tail_spmv<T>(C.n, L.R, L.r, brank, nb, red, [&](int row, T acc) {
            st_cg(C.b + row, acc);
            if (!coarsest) st_cg(cstart + row, ST<T>::scale(C.inv_theta, ld_ro(C.dinv + row) * acc));
        });
        cl.sync();
    }
    // ---- coarsest level: dense inverse
    {
        const TailLevel<T>& C = D.lv[m - 1];
        tail_dense<T>(C.n, D.coarse_inv, C.b, m == 1 ? xo_top : C.xo, brank, nb);
        cl.sync();
    }
    // ---- up: prolongation + post-smoothing
    for (int j = m - 2; j >= 0; j--) {
        const TailLevel<T>& L = D.lv[j];
        const TailLevel<T>& C = D.lv[j + 1];
        T* xo = j == 0 ? xo_top : L.xo;
        tail_spmv<T>(L.n, L.P, C.xo, brank, nb, red,
                     [&](int row, T acc) { st_cg(L.x + row, ld_cg(L.x + row) + acc); });
        cl.sync();
        if (L.deg == 1) {
            tail_spmv<T>(L.n, L.A, L.x, brank, nb, red, [&](int row, T acc) {
                const T ri = ld_cg(L.b + row) - acc;
                st_cg(xo + row, ld_cg(L.x + row) + ST<T>::scale(L.inv_theta, ld_ro(L.dinv + row) * ri));
            });
            cl.sync();
        } else {
            tail_spmv<T>(L.n, L.A, L.x, brank, nb, red, [&](int row, T acc) {
                const T ri = ld_cg(L.b + row) - acc;
                st_cg(L.r + row, ri);
                st_cg(L.d + row, ST<T>::scale(L.inv_theta, ld_ro(L.dinv + row) * ri));
            });
            cl.sync();
            T* d_cur = L.d;
            double rho = 1.0 / L.sigma;
            for (int k = 0; k + 1 < L.deg; k++) {
                const double rho1 = 1.0 / (2.0 * L.sigma - rho);
                const double c1 = rho1 * rho, c2 = 2.0 * rho1 / L.delta;
                const bool lastk = k + 2 == L.deg;
                T* d_next = d_cur == L.d ? L.q : L.d;
                const T* dc = d_cur;
                tail_spmv<T>(L.n, L.A, dc, brank, nb, red, [&](int row, T acc) {
                    const T di = ld_cg(dc + row);
                    const T ri = ld_cg(L.r + row) - acc;
                    const T dn = ST<T>::scale(c1, di) + ST<T>::scale(c2, ld_ro(L.dinv + row) * ri);
                    if (lastk) {
                        st_cg(xo + row, ld_cg(L.x + row) + di + dn);
                    } else {
                        st_cg(L.x + row, ld_cg(L.x + row) + di);
                        st_cg(L.r + row, ri);
                        st_cg(d_next + row, dn);
                    }
                });
                cl.sync();
                d_cur = d_next;
                rho = rho1;
            }
        }
    }
}

// cluster size the device accepts for the tail kernel: 16 (non-portable, one GPC) if possible, else 8
template <class T>
inline int tail_cluster_size() {
    static int cached = 0;
    if (cached) return cached;
    int best = 8;
    if (const char* e = getenv("MK_TAIL_CLUSTER")) {
        const int c = atoi(e);
        if (c == 1 || c == 2 || c == 4 || c == 8 || c == 16) { best = c; }
        if (best > 8) cudaFuncSetAttribute(k_vcycle_tail<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cached = best;
        return cached;
    }
    if (cudaFuncSetAttribute(k_vcycle_tail<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(kTailThreads);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, k_vcycle_tail<T>, &cfg) == cudaSuccess && nclusters >= 1) best = 16;
    }
    cudaGetLastError();
    cached = best;
    return cached;
}

template <class T>
inline void launch_vcycle_tail(const TailDesc<T>& D, T* xo_top, const PcgState* st, cudaStream_t s) {
    const int C = tail_cluster_size<T>();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C);
    cfg.blockDim = dim3(kTailThreads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    MK_CUDA(cudaLaunchKernelEx(&cfg, k_vcycle_tail<T>, D, xo_top, st));
}

inline int tail_rows_threshold() {
    if (const char* e = getenv("MK_TAIL_ROWS")) return atoi(e);
    return kTailRowsDefault;
}

}  // namespace mk
