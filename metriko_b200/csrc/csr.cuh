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
    DevBuf<int> ptr, col;
    DevBuf<T> val;
    DevBuf<int> long_rows;  // rows with more than kLongRow entries: one CTA each
    int n_long = 0;
    int group = 8;          // lanes per row in the vector kernel
};

constexpr int kLongRow = 512;

template <class T>
inline void upload_csr(DeviceCsr<T>& A, int rows, int cols, const std::vector<int>& ptr, const std::vector<int>& col,
                       const std::vector<T>& val, cudaStream_t s) {
    A.rows = rows;
    A.cols = cols;
    A.nnz = (int64_t)col.size();
    A.ptr.alloc(rows + 1);
    A.col.alloc(col.size());
    A.val.alloc(val.size());
    A.ptr.upload(ptr.data(), rows + 1, s);
    A.col.upload(col.data(), col.size(), s);
    A.val.upload(val.data(), val.size(), s);
    std::vector<int> lr;
    for (int i = 0; i < rows; i++)
        if (ptr[i + 1] - ptr[i] > kLongRow) lr.push_back(i);
    A.n_long = (int)lr.size();
    if (A.n_long) {
        A.long_rows.alloc(lr.size());
        A.long_rows.upload(lr.data(), lr.size(), s);
    }
    double avg = rows ? (double)A.nnz / rows : 0;
    A.group = avg <= 3 ? 2 : avg <= 6 ? 4 : avg <= 12 ? 8 : avg <= 24 ? 16 : 32;
    MK_CUDA(cudaStreamSynchronize(s));  // host vectors may die after return
}

// one CTA per long row
template <class T>
__global__ void __launch_bounds__(kThreads) k_spmv_long(const int* __restrict__ rows, const int* __restrict__ ptr,
                                                        const int* __restrict__ col, const T* __restrict__ val,
                                                        const T* __restrict__ x, T* __restrict__ y, const PcgState* st) {
    __shared__ double red[32];
    if (st && st->done) return;
    const int row = rows[blockIdx.x];
    const int b = ptr[row], e = ptr[row + 1];
    T acc = ST<T>::zero();
    for (int k = b + threadIdx.x; k < e; k += blockDim.x) acc = ST<T>::fma_(ld_ro(val + k), ld_ro(x + col[k]), acc);
    if constexpr (sizeof(T) == 16) {
        const cdouble* a = reinterpret_cast<const cdouble*>(&acc);
        double re = block_sum(a->x, red), im = block_sum(a->y, red);
        if (threadIdx.x == 0) { cdouble o = make_c(re, im); st_rw(reinterpret_cast<cdouble*>(y) + row, o); }
    } else {
        const double* a = reinterpret_cast<const double*>(&acc);
        double v = block_sum(*a, red);
        if (threadIdx.x == 0) reinterpret_cast<double*>(y)[row] = v;
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

// y = A x with G lanes per row; DOT: pq = <x, y>, alpha = rho / pq finished by the last CTA.
// SHIFT: y = A x + shift * x (used for A0 + 1e-4 I without touching the stored values).
template <class T, int G, bool DOT>
__global__ void __launch_bounds__(kThreads) k_spmv_csr(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                                                       const T* __restrict__ val, const T* __restrict__ x, T* __restrict__ y,
                                                       double shift, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st && st->done) return;
    const int gl = threadIdx.x & (G - 1);
    const int rows_per_pass = (gridDim.x * blockDim.x) / G;
    double dot = 0.0;
    for (int base = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / G; base < n; base += rows_per_pass) {
        const int row = base + ((threadIdx.x & 31) / G);
        T acc = ST<T>::zero();
        bool is_long = false;
        if (row < n) {
            const int b = ptr[row], e = ptr[row + 1];
            is_long = (e - b) > kLongRow;
            if (!is_long)
                for (int k = b + gl; k < e; k += G) acc = ST<T>::fma_(ld_ro(val + k), ld_ro(x + col[k]), acc);
        }
        acc = group_sum<T>(acc, G);
        if (gl == 0 && row < n) {
            T xr = ld_ro(x + row);
            if (is_long) acc = ld_rw(y + row);  // written by k_spmv_long, launched before this kernel
            else if (shift != 0.0) acc = acc + ST<T>::scale(shift, xr);
            if (!is_long) st_rw(y + row, acc);
            if (DOT) dot += ST<T>::hdot(xr, acc);
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

template <class T>
struct CsrOp : LinOp<T> {
    const DeviceCsr<T>* A = nullptr;
    double shift = 0.0;
    explicit CsrOp(const DeviceCsr<T>* A_, double shift_ = 0.0) : A(A_), shift(shift_) { this->n = A_->rows; }
    int launches_per_apply() const override { return A->n_long ? (shift != 0.0 ? 3 : 2) : 1; }

    template <int G>
    void launch(const T* x, T* y, PcgState* st, ReduceWs ws, cudaStream_t s, bool dot) {
        int64_t threads = (int64_t)A->rows * G;
        int grid = stream_grid(threads, kThreads, 16);
        if (st && dot) k_spmv_csr<T, G, true><<<grid, kThreads, 0, s>>>(A->rows, A->ptr.p, A->col.p, A->val.p, x, y, shift, st, ws);
        else k_spmv_csr<T, G, false><<<grid, kThreads, 0, s>>>(A->rows, A->ptr.p, A->col.p, A->val.p, x, y, shift, st, ws);
    }
    void apply(const T* x, T* y, PcgState* st, ReduceWs ws, cudaStream_t s, bool dot = true) override {
        if (A->n_long) {
            // long rows first; the shift is added here so the main kernel can just pick the value up
            k_spmv_long<T><<<A->n_long, kThreads, 0, s>>>(A->long_rows.p, A->ptr.p, A->col.p, A->val.p, x, y, st);
            if (shift != 0.0) add_shift_long(x, y, st, s);
        }
        switch (A->group) {
            case 2: launch<2>(x, y, st, ws, s, dot); break;
            case 4: launch<4>(x, y, st, ws, s, dot); break;
            case 8: launch<8>(x, y, st, ws, s, dot); break;
            case 16: launch<16>(x, y, st, ws, s, dot); break;
            default: launch<32>(x, y, st, ws, s, dot); break;
        }
        MK_CHECK_LAUNCH();
    }
    void add_shift_long(const T* x, T* y, PcgState* st, cudaStream_t s);
};

template <class T>
__global__ void k_shift_rows(int m, const int* __restrict__ rows, const T* __restrict__ x, T* __restrict__ y, double shift,
                             const PcgState* st) {
    if (st && st->done) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        int r = rows[i];
        st_rw(y + r, ld_rw(y + r) + ST<T>::scale(shift, ld_ro(x + r)));
    }
}
template <class T>
void CsrOp<T>::add_shift_long(const T* x, T* y, PcgState* st, cudaStream_t s) {
    k_shift_rows<T><<<(A->n_long + 127) / 128, 128, 0, s>>>(A->n_long, A->long_rows.p, x, y, shift, st);
}

}  // namespace mk
