// Multigrid hierarchy set-up on the device. Same algorithm as the host build in amg.cuh (which stays as its CPU mirror
// for the tests): MIS(2) aggregation with transported tentative values, P = (I - omega D^{-1} A) T, R = P^H,
// A_c = R (A P), Chebyshev bounds by power iteration. Every step is a kernel over rows or non-zeros; the Galerkin products are
// the atomic-free symbolic/numeric SpGEMM of devcsr.cuh. Nothing but the (<= 1500 row) coarsest operator returns to the host.
#pragma once
#include "amg.cuh"
#include "devcsr.cuh"

namespace mk {

template <class T> __device__ __forceinline__ double dev_abs(T v);
template <> __device__ __forceinline__ double dev_abs<double>(double v) { return fabs(v); }
template <> __device__ __forceinline__ double dev_abs<cdouble>(cdouble v) { return hypot(v.x, v.y); }  // = std::abs(std::complex)

// phase(-v) = -v / |v| (1 for v = 0), the transport factor of the tentative prolongator (amg.cuh unit_phase)
template <class T> __device__ __forceinline__ T dev_neg_phase(T v);
template <> __device__ __forceinline__ double dev_neg_phase<double>(double v) {
    const double a = fabs(v);
    return a > 0 ? (-v) / a : 1.0;
}
template <> __device__ __forceinline__ cdouble dev_neg_phase<cdouble>(cdouble v) {
    const double a = hypot(v.x, v.y);
    return a > 0 ? make_c(-v.x / a, -v.y / a) : make_c(1.0, 0.0);
}
template <class T> __device__ __forceinline__ T dev_one();
template <> __device__ __forceinline__ double dev_one<double>() { return 1.0; }
template <> __device__ __forceinline__ cdouble dev_one<cdouble>() { return make_c(1.0, 0.0); }
template <class T> __device__ __forceinline__ T dev_recip(T v);
template <> __device__ __forceinline__ double dev_recip<double>(double v) { return 1.0 / v; }
template <> __device__ __forceinline__ cdouble dev_recip<cdouble>(cdouble v) {
    const double d = v.x * v.x + v.y * v.y;
    return make_c(v.x / d, -v.y / d);
}

// dense[i] = row longer than `limit`
static __global__ void k_flag_dense_rows(int n, const int* __restrict__ ptr, int limit, char* __restrict__ dense, int* __restrict__ n_dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const char d = (ptr[i + 1] - ptr[i] > limit) ? 1 : 0;
    dense[i] = d;
    if (d) atomicAdd(n_dense, 1);
}

// dabs[i] = |a_ii| ; dinv[i] = 1 / a_ii, or 1 / (l1 norm of the row) for dense dofs (amg.cuh build_hierarchy)
template <class T>
static __global__ void k_diag_dinv(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, const char* __restrict__ dense,
                                   double* __restrict__ dabs, T* __restrict__ dinv) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        const int b = ptr[i], e = ptr[i + 1];
        if (dense && dense[i]) {
            double l1 = 0.0;
            for (int k = b + lane; k < e; k += 32) l1 += dev_abs<T>(ld_ro(val + k));
            l1 = warp_sum(l1);
            const int p = find_col(col, b, e, i);
            if (lane == 0) {
                dabs[i] = p >= 0 ? dev_abs<T>(ld_ro(val + p)) : 0.0;
                T r = dev_one<T>();
                if (l1 > 0) r = ST<T>::scale(1.0 / l1, r);
                st_rw(dinv + i, r);
            }
        } else if (lane == 0) {
            const int p = find_col(col, b, e, i);
            T d = p >= 0 ? ld_ro(val + p) : ST<T>::zero();
            const double a = dev_abs<T>(d);
            dabs[i] = a;
            st_rw(dinv + i, a > 0 ? dev_recip<T>(d) : dev_one<T>());
        }
    }
}

// ---- MIS(2): the synchronous rounds of aggregate_mis2 (amg.cuh), one thread per node ----
template <class T>
__device__ __forceinline__ bool dev_strong(int i, int j, T v, const char* __restrict__ dense, const double* __restrict__ dabs, double theta2,
                                           const int* __restrict__ type) {
    if (j == i || (dense && (dense[j] || dense[i]))) return false;
    if (type && type[i] != type[j]) return false;
    const double a = dev_abs<T>(v);
    return a > 0 && a * a >= theta2 * dabs[i] * dabs[j];
}

template <class T>
static __global__ void k_mis_t1(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, const char* __restrict__ dense,
                                const double* __restrict__ dabs, double theta2, const int* __restrict__ type, const int* __restrict__ state, int hashed,
                                uint64_t* __restrict__ t1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t m = mis_key(state[i], i, hashed != 0);
    if (!(dense && dense[i]))
        for (int k = ptr[i]; k < ptr[i + 1]; k++) {
            const int j = col[k];
            if (dev_strong<T>(i, j, ld_ro(val + k), dense, dabs, theta2, type)) {
                const uint64_t kj = mis_key(state[j], j, hashed != 0);
                m = kj > m ? kj : m;
            }
        }
    t1[i] = m;
}

template <class T>
static __global__ void k_mis_t2_update(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, const char* __restrict__ dense,
                                       const double* __restrict__ dabs, double theta2, const int* __restrict__ type, const uint64_t* __restrict__ t1,
                                       const int* __restrict__ state,
                                       int* __restrict__ state_out, int* __restrict__ undecided) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = state[i];
    int out = st;
    if (st == 1) {
        uint64_t m = t1[i];
        for (int k = ptr[i]; k < ptr[i + 1]; k++) {
            const int j = col[k];
            if (dev_strong<T>(i, j, ld_ro(val + k), dense, dabs, theta2, type)) {
                const uint64_t kj = t1[j];
                m = kj > m ? kj : m;
            }
        }
        if (mis_key_index(m) == i) out = 2;
        else if ((m >> 62) == 2) out = 0;
        else atomicAdd(undecided, 1);
    }
    state_out[i] = out;
}

static __global__ void k_flag_eq(int n, const int* __restrict__ v, int what, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = v[i] == what ? 1 : 0;
}

// roots get their scanned id and tentative value 1
template <class T>
static __global__ void k_agg_roots(int n, const int* __restrict__ state, const int* __restrict__ id, int* __restrict__ agg, T* __restrict__ tv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (state[i] == 2) { agg[i] = id[i]; st_rw(tv + i, dev_one<T>()); }
    else { agg[i] = -1; st_rw(tv + i, ST<T>::zero()); }
}

// rounds 1 and 2 of the aggregate assignment: join through the strongest coupling to an already assigned neighbour
// (round 1: roots only). Reads agg_in / tv_in, writes agg_out / tv_out.
template <class T>
static __global__ void k_agg_join(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, const char* __restrict__ dense,
                                  const double* __restrict__ dabs, double theta2, const int* __restrict__ type, const int* __restrict__ state, int roots_only,
                                  const int* __restrict__ agg_in, const T* __restrict__ tv_in, int* __restrict__ agg_out, T* __restrict__ tv_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = agg_in[i];
    T t = ld_ro(tv_in + i);
    if (a == -1 && !(dense && dense[i])) {
        double best = -1.0;
        int bk = -1;
        for (int k = ptr[i]; k < ptr[i + 1]; k++) {
            const int j = col[k];
            const T v = ld_ro(val + k);
            if (!dev_strong<T>(i, j, v, dense, dabs, theta2, type)) continue;
            if (roots_only ? (state[j] != 2) : (agg_in[j] == -1)) continue;
            const double w = dev_abs<T>(v);
            if (w > best * kStrongerBy) { best = w; bk = k; }
        }
        if (bk >= 0) {
            const int j = col[bk];
            a = agg_in[j];
            t = dev_neg_phase<T>(ld_ro(val + bk)) * ld_ro(tv_in + j);
        }
    }
    agg_out[i] = a;
    st_rw(tv_out + i, t);
}

// flag = 1 for nodes that still need an aggregate of their own: pass 0 = leftovers (unassigned, not dense), pass 1 = dense dofs
static __global__ void k_flag_singletons(int n, const int* __restrict__ agg, const char* __restrict__ dense, int pass, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool d = dense && dense[i];
    flag[i] = pass == 0 ? ((agg[i] == -1 && !d) ? 1 : 0) : (d ? 1 : 0);
}

template <class T>
static __global__ void k_assign_singletons(int n, const int* __restrict__ flag, const int* __restrict__ id, int base, int* __restrict__ agg, T* __restrict__ tv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) { agg[i] = base + id[i]; st_rw(tv + i, dev_one<T>()); }
}

// size[a] = sum of the members' weights^2. `cnt` (null on the finest level: every dof counts 1) = how many finest-level dofs a dof of this
// level stands for = the SQUARE of its weight in the tentative prolongator (AmgParams::weighted), so the sums stay integers on every level:
// exact in any order, the hierarchy stays bit-reproducible.
static __global__ void k_agg_sizes(int n, const int* __restrict__ agg, const int* __restrict__ cnt, int* __restrict__ size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(size + agg[i], cnt ? cnt[i] : 1);
}

// T: one entry per row, (i, agg[i]) = tv[i] * w_i / sqrt(sum of w^2 over the aggregate), |tv| = 1, w_i = sqrt(cnt[i]) (1 on the finest level)
template <class T>
static __global__ void k_build_tentative(int n, const int* __restrict__ agg, const T* __restrict__ tv, const int* __restrict__ cnt, const int* __restrict__ size,
                                         int* __restrict__ tptr, int* __restrict__ tcol, T* __restrict__ tval) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    tptr[i] = i;
    if (i == n) return;
    const int a = agg[i];
    tcol[i] = a;
    const double nrm = sqrt((double)size[a]);
    const double w = cnt ? sqrt((double)cnt[i]) : 1.0;
    const T t = ld_ro(tv + i);
    st_rw(tval + i, ST<T>::scale(w / nrm, t));
}

// S = I - omega D^{-1} A on the pattern of A -- except the rows of dense dofs, which are the identity (one entry): dense dofs are
// interpolated by injection from their own singleton aggregate. Smoothing their row would spread them over every neighbouring aggregate
// and couple all of those pairwise in the Galerkin product (a clique per seam: 35 instead of 13 entries per row on level 1 of the IGM
// operator, 89 instead of 23 on level 2, at unchanged iteration counts).
static __global__ void k_smoother_lengths(int n, const int* __restrict__ ptr, const char* __restrict__ dense, int* __restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = (dense && dense[i]) ? 1 : ptr[i + 1] - ptr[i];
}
template <class T>
static __global__ void k_smoother_matrix(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, const T* __restrict__ dinv,
                                         double omega, const char* __restrict__ dense, const int* __restrict__ sptr, int* __restrict__ scol,
                                         T* __restrict__ sval) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        const int s0 = sptr[i];
        if (dense && dense[i]) {
            if (lane == 0) { scol[s0] = i; st_rw(sval + s0, dev_one<T>()); }
            continue;
        }
        const T d = ST<T>::scale(-omega, ld_ro(dinv + i));
        const int b = ptr[i];
        for (int k = b + lane; k < ptr[i + 1]; k += 32) {
            const int c = col[k];
            T v = d * ld_ro(val + k);
            if (c == i) v = v + dev_one<T>();
            scol[s0 + k - b] = c;
            st_rw(sval + s0 + k - b, v);
        }
    }
}

// ---- power iteration for rho(D^{-1} A) (amg.cuh estimate_rho_dinvA) ----
template <class T>
static __global__ void k_rho_init(int n, T* __restrict__ v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T t = ST<T>::zero();
        *reinterpret_cast<double*>(&t) = sin(1.0 + 0.7 * (double)i) + 0.3;
        st_rw(v + i, t);
    }
}
// w = dinv .* w ; out[0] = |w|^2, out[1] = |v|^2
template <class T>
static __global__ void __launch_bounds__(kThreads) k_rho_norms(int n, const T* __restrict__ dinv, T* __restrict__ w, const T* __restrict__ v, double* __restrict__ out,
                                                               ReduceWs ws) {
    __shared__ double red[32];
    double a0 = 0.0, a1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const T wi = ld_ro(dinv + i) * ld_rw(w + i);
        st_rw(w + i, wi);
        a0 += ST<T>::abs2(wi);
        a1 += ST<T>::abs2(ld_ro(v + i));
    }
    double r[2] = {a0, a1};
    grid_reduce<2>(r, ws.partials, ws.ticket, red, [&](const double* s) { out[0] = s[0]; out[1] = s[1]; });
}
// v = w / sqrt(out[0])
template <class T>
static __global__ void __launch_bounds__(kThreads) k_rho_scale(int n, const T* __restrict__ w, T* __restrict__ v, const double* __restrict__ out) {
    const double sc = 1.0 / sqrt(out[0]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_rw(v + i, ST<T>::scale(sc, ld_rw(w + i)));
}

template <class T>
inline double estimate_rho_dev(const DeviceCsr<T>& A, const T* dinv, cudaStream_t s, LinOp<T>* fast_op = nullptr, int iters = 12) {
    const int n = A.rows;
    DevBuf<T> v(n), w(n);
    DevBuf<double> out(2), partials((size_t)kMaxPartials * 4);
    DevBuf<unsigned int> ticket(1);
    ticket.zero(s);
    const int grid = stream_grid(n);
    k_rho_init<T><<<grid, kThreads, 0, s>>>(n, v.p);
    for (int it = 0; it < iters; it++) {
        // level 0 behind the solver's own operator (face stencil / SELL): 32-45 us per product instead of ~350 for the set-up's plain CSR kernel
        if (fast_op) fast_op->apply(v.p, w.p, nullptr, ReduceWs{partials.p, ticket.p}, s, false);
        else csr_matvec<T>(A, v.p, w.p, 1.0, s);
        k_rho_norms<T><<<grid, kThreads, 0, s>>>(n, dinv, w.p, v.p, out.p, ReduceWs{partials.p, ticket.p});
        k_rho_scale<T><<<grid, kThreads, 0, s>>>(n, w.p, v.p, out.p);
    }
    MK_CHECK_LAUNCH();
    double h[2] = {1.0, 1.0};
    MK_CUDA(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    return std::sqrt(h[0] / h[1]);
}

// a coarse dof inherits the type of its members (all equal: typed aggregation never mixes them; the racing writes store the same value)
static __global__ void k_coarse_types(int n, const int* __restrict__ agg, const int* __restrict__ type, int* __restrict__ ctype) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ctype[agg[i]] = type[i];
}

// Aggregation of one level on the device. Outputs the tentative prolongator T (n x nc, one entry per row); returns nc.
// dense (may be null): dense dofs of this level; dense_next (size >= n): flags of the coarse level (its last n_dense ids).
template <class T>
inline int aggregate_dev(const DeviceCsr<T>& A, double theta, const char* dense, int n_dense, const double* dabs, DeviceCsr<T>& Tm, cudaStream_t s,
                         const int* type = nullptr, DevBuf<int>* type_next = nullptr, const int* cnt = nullptr, DevBuf<int>* cnt_next = nullptr) {
    const int n = A.rows;
    const int g = (n + 255) / 256;
    const double theta2 = theta * theta;
    DevBuf<int> state(n), state2(n), counter(1), flag((size_t)n + 1), id((size_t)n + 1), agg(n), agg2(n);
    DevBuf<uint64_t> t1(n);
    DevBuf<T> tv(n), tv2(n);
    // state: dense dofs removed from the start, everything else undecided
    k_flag_singletons<<<g, 256, 0, s>>>(n, nullptr, dense, 1, flag.p);   // flag = dense
    k_flag_eq<<<g, 256, 0, s>>>(n, flag.p, 0, state.p);                 // state = dense ? 0 : 1
    MK_CHECK_LAUNCH();
    int* cur = state.p;
    int* nxt = state2.p;
    for (int round = 0; round < kMisMaxRounds; round++) {
        const int hashed = round >= kMisIndexRounds ? 1 : 0;
        counter.zero(s);
        k_mis_t1<T><<<g, 256, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, dense, dabs, theta2, type, cur, hashed, t1.p);
        k_mis_t2_update<T><<<g, 256, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, dense, dabs, theta2, type, t1.p, cur, nxt, counter.p);
        MK_CHECK_LAUNCH();
        std::swap(cur, nxt);
        if (dev_read_int(counter.p, s) == 0) break;
    }
    // roots numbered in index order
    k_flag_eq<<<g, 256, 0, s>>>(n, cur, 2, flag.p);
    dev_exclusive_scan(flag.p, id.p, n, s);
    int nagg = dev_read_int(id.p + n, s);
    k_agg_roots<T><<<g, 256, 0, s>>>(n, cur, id.p, agg.p, tv.p);
    k_agg_join<T><<<g, 256, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, dense, dabs, theta2, type, cur, 1, agg.p, tv.p, agg2.p, tv2.p);
    k_agg_join<T><<<g, 256, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, dense, dabs, theta2, type, cur, 0, agg2.p, tv2.p, agg.p, tv.p);
    // leftovers, then dense dofs, as singletons in index order
    k_flag_singletons<<<g, 256, 0, s>>>(n, agg.p, dense, 0, flag.p);
    dev_exclusive_scan(flag.p, id.p, n, s);
    const int n_left = dev_read_int(id.p + n, s);
    if (n_left) k_assign_singletons<T><<<g, 256, 0, s>>>(n, flag.p, id.p, nagg, agg.p, tv.p);
    nagg += n_left;
    if (n_dense) {
        k_flag_singletons<<<g, 256, 0, s>>>(n, agg.p, dense, 1, flag.p);
        dev_exclusive_scan(flag.p, id.p, n, s);
        k_assign_singletons<T><<<g, 256, 0, s>>>(n, flag.p, id.p, nagg, agg.p, tv.p);
        nagg += n_dense;
    }
    MK_CHECK_LAUNCH();
    DevBuf<int> size((size_t)std::max(nagg, 1));
    size.zero(s);
    k_agg_sizes<<<g, 256, 0, s>>>(n, agg.p, cnt, size.p);
    Tm.rows = n;
    Tm.cols = nagg;
    Tm.nnz = n;
    Tm.ptr.alloc((size_t)n + 1);
    Tm.col.alloc(n);
    Tm.val.alloc(n);
    k_build_tentative<T><<<(n + 256) / 256, 256, 0, s>>>(n, agg.p, tv.p, cnt, size.p, Tm.ptr.p, Tm.col.p, Tm.val.p);
    if (type && type_next) {
        type_next->alloc((size_t)std::max(nagg, 1));
        type_next->zero(s);
        k_coarse_types<<<g, 256, 0, s>>>(n, agg.p, type, type_next->p);
    }
    MK_CHECK_LAUNCH();
    MK_CUDA(cudaStreamSynchronize(s));
    if (cnt_next) *cnt_next = std::move(size);  // the coarse dofs' counts (weights^2)
    return nagg;
}

// ---- dense inverse of the coarsest operator on the device: in-place Gauss-Jordan without pivoting (the Galerkin operator is Hermitian
// positive definite), two launches per pivot, the n x n matrix stays in L2. 942 complex rows: ~10 ms against 137 ms for the host's
// Gauss-Jordan, which is what allowed the coarsest level to stop at <= 1500 rows instead of being coarsened once more.
__device__ __forceinline__ double dev_recip(double a) { return 1.0 / a; }
__device__ __forceinline__ cdouble dev_recip(cdouble a) {
    const double d = a.x * a.x + a.y * a.y;
    return make_c(a.x / d, -a.y / d);
}
__device__ __forceinline__ double dev_neg(double a) { return -a; }
__device__ __forceinline__ cdouble dev_neg(cdouble a) { return make_c(-a.x, -a.y); }

template <class T>
static __global__ void k_dense_from_csr(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val, T* __restrict__ M) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    for (int j = lane; j < n; j += 32) st_rw(M + (size_t)i * n + j, ST<T>::zero());
    __syncwarp();
    for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) st_rw(M + (size_t)i * n + col[k], ld_ro(val + k));  // sorted rows, distinct columns
}
// pivot k: rowk[j] = M[k][j] / M[k][k] (rowk[k] = 1 / M[k][k]), colk[i] = M[i][k]
template <class T>
static __global__ void k_gj_prep(int n, int k, const T* __restrict__ M, T* __restrict__ rowk, T* __restrict__ colk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const T p = dev_recip(M[(size_t)k * n + k]);
    rowk[j] = j == k ? p : ST<T>::fma_(M[(size_t)k * n + j], p, ST<T>::zero());
    colk[j] = M[(size_t)j * n + k];
}
template <class T>
static __global__ void k_gj_update(int n, int k, T* __restrict__ M, const T* __restrict__ rowk, const T* __restrict__ colk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    const T r = rowk[j];
    T out;
    if (i == k) out = r;
    else {
        const T c = dev_neg(colk[i]);
        out = j == k ? ST<T>::fma_(c, r, ST<T>::zero()) : ST<T>::fma_(c, r, M[(size_t)i * n + j]);
    }
    M[(size_t)i * n + j] = out;
}
template <class T>
inline void dense_inverse_dev(const DeviceCsr<T>& A, DevBuf<T>& inv, cudaStream_t s) {
    const int n = A.rows;
    inv.alloc((size_t)n * n);
    DevBuf<T> rowk(n), colk(n);
    k_dense_from_csr<T><<<(n * 32 + 255) / 256, 256, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, inv.p);
    const dim3 grid((n + 255) / 256, n);
    for (int k = 0; k < n; k++) {
        k_gj_prep<T><<<(n + 255) / 256, 256, 0, s>>>(n, k, inv.p, rowk.p, colk.p);
        k_gj_update<T><<<grid, 256, 0, s>>>(n, k, inv.p, rowk.p, colk.p);
    }
    MK_CHECK_LAUNCH();
}

static __global__ void k_flag_tail(int n, int first, char* __restrict__ dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dense[i] = i >= first ? 1 : 0;
}

// Builds the hierarchy of A (device CSR, sorted rows, explicit diagonal) on the device. A is consumed.
// `fine_op` non-null: level 0 applies that operator in the V-cycle, the CSR copy of level 0 is released after the set-up.
// keep_csr: keep every level's plain CSR arrays (tests read them back).
template <class T>
inline std::unique_ptr<AmgPrecond<T>> build_amg_device(cudaStream_t s, DeviceCsr<T>&& A0, LinOp<T>* fine_op, const AmgParams& prm,
                                                       DevBuf<int> types = DevBuf<int>()) {
    using S = typename HostScalar<T>::type;
    auto amg = std::make_unique<AmgPrecond<T>>();
    amg->prm = prm;
    amg->fine_op = fine_op;
    StageTimer tm("build_amg_device");
    const double nnz0 = (double)A0.nnz, n0 = (double)A0.rows;
    double nnz_sum = 0, n_sum = 0;
    DeviceCsr<T> A = std::move(A0);
    DevBuf<char> dense;
    DevBuf<int> counts;  // finest-level dofs per dof of the current level (null on the finest level); AmgParams::weighted
    int n_dense = 0;
    if (prm.dense_row > 0 && A.rows > prm.coarse_max) {
        dense.alloc(A.rows);
        DevBuf<int> cnt(1);
        cnt.zero(s);
        k_flag_dense_rows<<<(A.rows + 255) / 256, 256, 0, s>>>(A.rows, A.ptr.p, prm.dense_row, dense.p, cnt.p);
        n_dense = dev_read_int(cnt.p, s);
        if (!n_dense) dense.release();
    }
    while (true) {
        amg->lv.emplace_back();
        AmgLevelDev<T>& L = amg->lv.back();
        const int l = (int)amg->lv.size() - 1;
        const int n = A.rows;
        L.n = n;
        nnz_sum += (double)A.nnz;
        n_sum += n;
        L.dinv.alloc(n);
        DevBuf<double> dabs(n);
        k_diag_dinv<T><<<stream_grid((int64_t)n * 32), kThreads, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, n_dense ? dense.p : nullptr, dabs.p, L.dinv.p);
        MK_CHECK_LAUNCH();
        const double rho = estimate_rho_dev<T>(A, L.dinv.p, s, (l == 0 && fine_op && fine_op->applies()) ? fine_op : nullptr);
        L.lmax = 1.1 * rho;
        L.r.alloc(n); L.d.alloc(n); L.q.alloc(n);
        if (l > 0) { L.x.alloc(n); L.b.alloc(n); }
        tm.lap("dinv + rho");
        bool last = (n - n_dense <= prm.coarse_max) || (l + 1 >= prm.max_levels);
        DeviceCsr<T> Tm;
        DevBuf<int> types_next, counts_next;
        int nc = 0;
        if (!last) {
            nc = aggregate_dev<T>(A, prm.theta, n_dense ? dense.p : nullptr, n_dense, dabs.p, Tm, s, types.p, &types_next,
                                  prm.weighted ? counts.p : nullptr, prm.weighted ? &counts_next : nullptr);
            tm.lap("aggregate");
            if (nc >= n * 0.8 || nc == 0) last = true;
        }
        if (last) {
            L.A = std::move(A);
            break;
        }
        // P = (I - omega D^{-1} A) T
        const double omega = prm.omega_scale / rho;
        {
            DeviceCsr<T> Sm;
            const char* inj = (n_dense && prm.inject_dense) ? dense.p : nullptr;
            Sm.rows = Sm.cols = n;
            Sm.ptr.alloc((size_t)n + 1);
            {
                DevBuf<int> len((size_t)n + 1);
                len.zero(s);
                k_smoother_lengths<<<(n + 255) / 256, 256, 0, s>>>(n, A.ptr.p, inj, len.p);
                dev_exclusive_scan(len.p, Sm.ptr.p, n, s);
                Sm.nnz = dev_read_int(Sm.ptr.p + n, s);
            }
            Sm.col.alloc((size_t)std::max<int64_t>(Sm.nnz, 1));
            Sm.val.alloc((size_t)std::max<int64_t>(Sm.nnz, 1));
            k_smoother_matrix<T><<<stream_grid((int64_t)n * 32), kThreads, 0, s>>>(n, A.ptr.p, A.col.p, A.val.p, L.dinv.p, omega, inj, Sm.ptr.p, Sm.col.p,
                                                                                  Sm.val.p);
            MK_CHECK_LAUNCH();
            spgemm<T>(Sm, Tm, L.P, s);
        }
        tm.lap("P = S T");
        transpose_dev<T>(L.P, L.R, true, s);
        tm.lap("R = P^H");
        DeviceCsr<T> AP, Ac;
        spgemm<T>(A, L.P, AP, s);
        tm.lap("A P");
        spgemm<T>(L.R, AP, Ac, s);
        tm.lap("R (A P)");
        build_sell(L.P, s);
        build_sell(L.R, s);
        L.A = std::move(A);
        if (l == 0 && fine_op) {
            L.A.ptr.release(); L.A.col.release(); L.A.val.release();  // the V-cycle applies fine_op on level 0
        } else {
            build_sell(L.A, s);
        }
        tm.lap("SELL layouts");
        if (n_dense) {
            dense.alloc(nc);
            k_flag_tail<<<(nc + 255) / 256, 256, 0, s>>>(nc, nc - n_dense, dense.p);
            MK_CHECK_LAUNCH();
        }
        A = std::move(Ac);
        if (types.p) types = std::move(types_next);
        counts = std::move(counts_next);
    }
    // coarsest level: dense inverse on the host (<= 1500 rows) or Chebyshev
    AmgLevelDev<T>& C = amg->lv.back();
    amg->coarse_n = C.n;
    if (C.n <= 1500) {
        if (C.n > 128) {
            dense_inverse_dev<T>(C.A, amg->coarse_inv, s);
        } else {  // tiny: the host's pivoted Gauss-Jordan is as fast as the launches
            HostCsr<S> H = download_csr<T, S>(C.A, s);
            std::vector<S> inv = dense_inverse(H);
            amg->coarse_inv.alloc(inv.size());
            amg->coarse_inv.upload((const T*)inv.data(), inv.size(), s);
        }
        MK_CUDA(cudaStreamSynchronize(s));
        amg->coarse_dense = true;
    }
    if (!(amg->lv.size() == 1 && fine_op)) build_sell(C.A, s);
    tm.lap("coarsest level");
    amg->op_complexity = nnz_sum / nnz0;
    amg->grid_complexity = n_sum / n0;
    if (tm.on) {
        fprintf(stderr, "[mk] build_amg_device: levels");
        for (auto& lv : amg->lv) fprintf(stderr, " %d", lv.n);
        fprintf(stderr, " op_cx %.2f grid_cx %.2f\n", amg->op_complexity, amg->grid_complexity);
    }
    MK_CUDA(cudaStreamSynchronize(s));
    return amg;
}

}  // namespace mk
