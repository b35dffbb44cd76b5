// Device CSR matrix + vectorised (G lanes per row) SpMV with an optional fused <x, A x> reduction.
// Used for the IGM normal matrix A (real fp64, ~7 nnz/row plus a few very long kernel-column rows), for
// generic Hermitian inputs arriving as Eigen CSC through the drop-in seam, and for all multigrid levels.
#pragma once
#include <algorithm>
#include <vector>

#include "solver.cuh"

namespace mk {

template <class T>
struct DeviceCsr {
    int rows = 0, cols = 0;
    int64_t nnz = 0;
    DevBuf<int> ptr, col;   // CSR as given (long rows are served from it; also what the get_matrix entry points return)
    DevBuf<T> val;
    // SELL-32: rows in slices of one warp; slice s stores width_s x 32 entries column-major (entry k of lane l at
    // sl_ptr[s] + 32 k + l), padded with (a valid column, 0). One thread per row, every load coalesced across the warp.
    int n_slices = 0;
    DevBuf<int> sl_ptr, sl_col;
    DevBuf<T> sl_val;
    double padding = 1.0;   // stored / true entries
    // rows with more than kLongRow entries (dense dofs): empty in the SELL part, one CTA each, results parked in ylong[row]
    DevBuf<int> long_rows;
    DevBuf<unsigned int> long_bits;  // bit (row & 31) of word (row >> 5)
    DevBuf<T> ylong;
    int n_long = 0;
};

constexpr int kLongRow = 256;

// ---- CSR -> SELL-32 conversion on the device (once per operator) ----
// one warp per slice: width = longest non-long row, long rows flagged in a bit word and appended to the long-row list
static __global__ void k_sell_width(int rows, int n_slices, const int* __restrict__ ptr, int* __restrict__ width32, unsigned int* __restrict__ long_bits,
                                    int* __restrict__ long_rows, int* __restrict__ n_long) {
    const int lane = threadIdx.x & 31;
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sl >= n_slices) return;
    const int row = sl * 32 + lane;
    int len = 0;
    if (row < rows) len = ptr[row + 1] - ptr[row];
    const bool is_long = len > kLongRow;
    const unsigned int bits = __ballot_sync(0xffffffffu, is_long);
    if (is_long) {
        long_rows[atomicAdd(n_long, 1)] = row;
        len = 0;
    }
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) { width32[sl] = len * 32; long_bits[sl] = bits; }
}

// exclusive scan of n ints by one CTA (n = number of slices: 65k at 2M rows); out[n] = total
static __global__ void __launch_bounds__(1024) k_scan_exclusive(int n, const int* __restrict__ in, int* __restrict__ out) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = carry + x + (wid ? warp_tot[wid - 1] : 0);
        if (i < n) out[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

template <class T>
static __global__ void k_sell_fill(int rows, int n_slices, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val,
                                   const int* __restrict__ sl_ptr, int* __restrict__ sl_col, T* __restrict__ sl_val) {
    const int lane = threadIdx.x & 31;
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sl >= n_slices) return;
    const int row = sl * 32 + lane;
    const int base = sl_ptr[sl], w = (sl_ptr[sl + 1] - base) >> 5;
    int b = 0, len = 0;
    if (row < rows) { b = ptr[row]; len = ptr[row + 1] - b; if (len > kLongRow) len = 0; }
    const int padc = len > 0 ? col[b] : 0;
    for (int k = 0; k < w; k++) {
        const size_t o = (size_t)base + (size_t)k * 32 + lane;
        if (k < len) { sl_col[o] = col[b + k]; st_rw(sl_val + o, ld_ro(val + b + k)); }
        else { sl_col[o] = padc; st_rw(sl_val + o, ST<T>::zero()); }
    }
}

// SELL-32 layout (+ long-row list) from the plain CSR arrays already on the device (rows, cols, nnz, ptr, col, val set)
template <class T>
inline void build_sell(DeviceCsr<T>& A, cudaStream_t s) {
    const int rows = A.rows;
    const int ns = (rows + 31) / 32;
    A.n_slices = ns;
    DevBuf<int> width(std::max(ns, 1)), counter(1);
    A.sl_ptr.alloc((size_t)ns + 1);
    A.long_bits.alloc(std::max(ns, 1));
    A.long_rows.alloc(std::max(1, std::min(rows, 1 << 16)));  // more than 65k dense rows would not be a sparse operator
    counter.zero(s);
    const int wg = std::max(1, (ns * 32 + kThreads - 1) / kThreads);
    k_sell_width<<<wg, kThreads, 0, s>>>(rows, ns, A.ptr.p, width.p, A.long_bits.p, A.long_rows.p, counter.p);
    k_scan_exclusive<<<1, 1024, 0, s>>>(ns, width.p, A.sl_ptr.p);
    MK_CHECK_LAUNCH();
    int tot = 0, nl = 0;
    MK_CUDA(cudaMemcpyAsync(&tot, A.sl_ptr.p + ns, sizeof(int), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaMemcpyAsync(&nl, counter.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    if (tot < 0 || nl > (1 << 16)) throw std::invalid_argument("build_sell: matrix too large / too many dense rows for the SELL layout");
    A.n_long = nl;
    A.sl_col.alloc((size_t)std::max(tot, 1));
    A.sl_val.alloc((size_t)std::max(tot, 1));
    k_sell_fill<T><<<wg, kThreads, 0, s>>>(rows, ns, A.ptr.p, A.col.p, A.val.p, A.sl_ptr.p, A.sl_col.p, A.sl_val.p);
    MK_CHECK_LAUNCH();
    A.padding = A.nnz ? (double)tot / (double)A.nnz : 1.0;
    if (A.n_long) {
        A.ylong.alloc(rows);
        A.ylong.zero(s);
    }
}

// plain CSR arrays only (operands of the device set-up algebra, devcsr.cuh)
template <class T>
inline void upload_csr_plain(DeviceCsr<T>& A, int rows, int cols, const int* ptr, const int* col, const T* val, int64_t nnz, cudaStream_t s) {
    A.rows = rows;
    A.cols = cols;
    A.nnz = nnz;
    A.ptr.alloc((size_t)rows + 1);
    A.col.alloc((size_t)std::max<int64_t>(nnz, 1));
    A.val.alloc((size_t)std::max<int64_t>(nnz, 1));
    A.ptr.upload(ptr, (size_t)rows + 1, s);
    A.col.upload(col, (size_t)nnz, s);
    A.val.upload(val, (size_t)nnz, s);
}

template <class T>
inline void upload_csr(DeviceCsr<T>& A, int rows, int cols, const std::vector<int>& ptr, const std::vector<int>& col,
                       const std::vector<T>& val, cudaStream_t s) {
    upload_csr_plain<T>(A, rows, cols, ptr.data(), col.data(), val.data(), (int64_t)col.size(), s);
    build_sell(A, s);  // synchronises the stream: the host vectors may die after return
}

// one CTA per long row: ylong[row] = sum_k val_k x[col_k]
template <class T>
__global__ void __launch_bounds__(kThreads) k_spmv_long(const int* __restrict__ rows, const int* __restrict__ ptr,
                                                        const int* __restrict__ col, const T* __restrict__ val,
                                                        const T* __restrict__ x, T* __restrict__ ylong, const PcgState* st) {
    __shared__ double red[32];
    if (st && st->done) return;
    const int row = rows[blockIdx.x];
    const int b = ptr[row], e = ptr[row + 1];
    T acc = ST<T>::zero();
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) acc = ST<T>::fma_(ld_ro(val + k), ld_ro(x + col[k]), acc);
    if constexpr (sizeof(T) == 16) {
        const cdouble* a = reinterpret_cast<const cdouble*>(&acc);
        double re = block_sum(a->x, red), im = block_sum(a->y, red);
        if (threadIdx.x == 0) { cdouble o = make_c(re, im); st_rw(reinterpret_cast<cdouble*>(ylong) + row, o); }
    } else {
        const double* a = reinterpret_cast<const double*>(&acc);
        double v = block_sum(*a, red);
        if (threadIdx.x == 0) reinterpret_cast<double*>(ylong)[row] = v;
    }
}

template <class T>
__device__ __forceinline__ T group_sum(T v, int G);
template <>
__device__ __forceinline__ double group_sum<double>(double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <>
__device__ __forceinline__ cdouble group_sum<cdouble>(cdouble v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// y (+)= A x (+ shift x) on the SELL-32 layout, one thread per row, one warp per slice.
// DOT: pq = <x, y>, alpha = rho / pq finished by the last CTA (the PCG step). ACC: y += (prolongation).
template <class T, bool DOT, bool ACC>
__global__ void __launch_bounds__(kThreads, 4) k_spmv_sell(int n, int n_slices, const int* __restrict__ sl_ptr, const int* __restrict__ sl_col,
                                                        const T* __restrict__ sl_val, const unsigned int* __restrict__ long_bits,
                                                        const T* __restrict__ ylong, const T* __restrict__ x, T* __restrict__ y, double shift,
                                                        PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sl < n_slices; sl += nwarps) {
        const int base = __ldg(sl_ptr + sl);
        const int w = (__ldg(sl_ptr + sl + 1) - base) >> 5;
        const int* cp = sl_col + base + lane;
        const T* vp = sl_val + base + lane;
        T acc = ST<T>::zero();
        int k = 0;
        for (; k + 4 <= w; k += 4) {
            const int c0 = __ldg(cp + 32 * k), c1 = __ldg(cp + 32 * (k + 1)), c2 = __ldg(cp + 32 * (k + 2)), c3 = __ldg(cp + 32 * (k + 3));
            const T v0 = ld_ro(vp + 32 * k), v1 = ld_ro(vp + 32 * (k + 1)), v2 = ld_ro(vp + 32 * (k + 2)), v3 = ld_ro(vp + 32 * (k + 3));
            const T x0 = ld_ro(x + c0), x1 = ld_ro(x + c1), x2 = ld_ro(x + c2), x3 = ld_ro(x + c3);
            acc = ST<T>::fma_(v0, x0, acc);
            acc = ST<T>::fma_(v1, x1, acc);
            acc = ST<T>::fma_(v2, x2, acc);
            acc = ST<T>::fma_(v3, x3, acc);
        }
        for (; k < w; k++) acc = ST<T>::fma_(ld_ro(vp + 32 * k), ld_ro(x + __ldg(cp + 32 * k)), acc);
        const int row = sl * 32 + lane;
        if (row < n) {
            if (long_bits != nullptr && ((__ldg(long_bits + sl) >> lane) & 1u)) acc = acc + ld_rw(ylong + row);
            if (DOT || shift != 0.0) {
                const T xr = ld_ro(x + row);
                if (shift != 0.0) acc = acc + ST<T>::scale(shift, xr);
                if (DOT) dot += ST<T>::hdot(xr, acc);
            }
            if (ACC) acc = acc + ld_rw(y + row);
            st_rw(y + row, acc);
        }
    }
    if (DOT) {
        double v[1] = {dot};
        grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
            st->pq = s[0];
            st->alpha = st->rho / s[0];
            if (!(s[0] > 0.0)) st->done = 2;  // not positive definite / breakdown
        });
    }
}

// y (+)= A x (+ shift x); st != nullptr: skipped once st->done is set; dot: fused <x, y> reduction (needs st)
template <class T>
inline void spmv(const DeviceCsr<T>& A, const T* x, T* y, double shift, bool accumulate, bool dot, PcgState* st, ReduceWs ws, cudaStream_t s) {
    if (A.n_long) k_spmv_long<T><<<A.n_long, kThreads, 0, s>>>(A.long_rows.p, A.ptr.p, A.col.p, A.val.p, x, A.ylong.p, st);
    const int grid = stream_grid((int64_t)A.n_slices * 32, kThreads, 16);
    const unsigned int* lb = A.n_long ? A.long_bits.p : nullptr;
    if (dot && st)
        k_spmv_sell<T, true, false><<<grid, kThreads, 0, s>>>(A.rows, A.n_slices, A.sl_ptr.p, A.sl_col.p, A.sl_val.p, lb, A.ylong.p, x, y, shift, st, ws);
    else if (accumulate)
        k_spmv_sell<T, false, true><<<grid, kThreads, 0, s>>>(A.rows, A.n_slices, A.sl_ptr.p, A.sl_col.p, A.sl_val.p, lb, A.ylong.p, x, y, shift, st, ws);
    else
        k_spmv_sell<T, false, false><<<grid, kThreads, 0, s>>>(A.rows, A.n_slices, A.sl_ptr.p, A.sl_col.p, A.sl_val.p, lb, A.ylong.p, x, y, shift, st, ws);
    MK_CHECK_LAUNCH();
}

template <class T>
struct CsrOp : LinOp<T> {
    const DeviceCsr<T>* A = nullptr;
    double shift = 0.0;
    explicit CsrOp(const DeviceCsr<T>* A_, double shift_ = 0.0) : A(A_), shift(shift_) { this->n = A_->rows; }
    int launches_per_apply() const override { return A->n_long ? 2 : 1; }
    void apply(const T* x, T* y, PcgState* st, ReduceWs ws, cudaStream_t s, bool dot = true) override {
        spmv<T>(*A, x, y, shift, false, dot && st != nullptr, st, ws, s);
    }
};

}  // namespace mk
