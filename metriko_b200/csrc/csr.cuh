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
    // rows with more than kLongRow entries (dense dofs): empty in the SELL part, served by one extra CTA each in the same launch
    DevBuf<int> long_rows;           // "medium" rows (long_thresh < length <= kHugeRow): one WARP each, 8 per extra CTA
    DevBuf<int> huge_rows;           // longer rows (dense dofs): one CTA each
    DevBuf<unsigned int> long_bits;  // bit (row & 31) of word (row >> 5): row is served by the extra CTAs, not by the bulk part
    int n_long = 0, n_huge = 0;
    // operators whose rows are long throughout (mean > kRowModeMean entries: the coarse IGM levels) skip the SELL copy and run the
    // warp-per-row kernel on the CSR arrays
    bool row_mode = false;
    int row_lanes = 32;  // lanes per row in row mode (4, 8, 16 or 32, from the mean row length)
};

constexpr int kLongRow = 64;               // SELL: rows longer than this leave the slices (one wide row would serialise its whole slice)
constexpr int kHugeRow = 2048;            // rows longer than this get a whole CTA, the ones in between a warp
constexpr int kRowModeMean = 40;          // mean row length above which SELL is not attempted
constexpr double kMaxSellPadding = 1.5;   // stored / true entries above which SELL is dropped for the lanes-per-row kernel ...
constexpr double kRowModeMinMean = 8.0;   // ... unless the rows are this short (a thread per row then beats lanes per row even with padding)
constexpr int kRowModeLong = 512;         // row mode: rows longer than this get a CTA of their own
constexpr int kRowModeSmallRows = 4096;   // operators this small are latency-bound: lanes per row (one load round trip per ~32 entries) instead
                                          // of a thread per row (one per entry). profiles/launches_r01f / r01g: the fused kernel on the
                                          // ~2,100-row, 29-entries-per-row IGM level takes 8.0-8.5 us as k_spmv_rows<.., 32> and 13.3 us as k_spmv_sell

// ---- CSR -> SELL-32 conversion on the device (once per operator) ----
// one warp per slice: width = longest non-long row, long rows flagged in a bit word and appended to the long-row list
static __global__ void k_sell_width(int rows, int n_slices, const int* __restrict__ ptr, int long_thresh, int* __restrict__ width32,
                                    unsigned int* __restrict__ long_bits, int* __restrict__ long_rows, int* __restrict__ huge_rows,
                                    int* __restrict__ n_long) {
    const int lane = threadIdx.x & 31;
    const int sl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sl >= n_slices) return;
    const int row = sl * 32 + lane;
    int len = 0;
    if (row < rows) len = ptr[row + 1] - ptr[row];
    const bool is_long = len > long_thresh;
    const unsigned int bits = __ballot_sync(0xffffffffu, is_long);
    if (is_long) {
        if (len > kHugeRow) huge_rows[atomicAdd(n_long + 1, 1)] = row;
        else long_rows[atomicAdd(n_long, 1)] = row;
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

// Chooses and builds the SpMV layout of an operator whose plain CSR arrays are on the device (rows, cols, nnz, ptr, col, val set):
//   SELL-32 when the rows of a slice have similar lengths (stored / true entries <= kMaxSellPadding: the fine operators and the face
//   hierarchy), otherwise "row mode": G lanes per row on the CSR arrays (the coarse IGM levels: mean 35-120 entries per row with a
//   heavy tail, where SELL would store 4-5x the entries). Either way the very long rows (dense dofs) get one CTA each.
template <class T>
inline void build_sell(DeviceCsr<T>& A, cudaStream_t s) {
    const int rows = A.rows;
    const int ns = (rows + 31) / 32;
    A.n_slices = ns;
    A.row_mode = false;
    A.n_long = 0;
    A.n_huge = 0;
    if (rows <= 0) return;
    DevBuf<int> width(std::max(ns, 1)), counter(2);
    A.sl_ptr.alloc((size_t)ns + 1);
    A.long_bits.alloc(std::max(ns, 1));
    A.long_rows.alloc(std::max(1, rows));
    A.huge_rows.alloc(std::max(1, std::min(rows, 1 << 16)));  // more than 65k dense rows would not be a sparse operator
    const int wg = std::max(1, (ns * 32 + kThreads - 1) / kThreads);
    int nh = 0;
    auto scan_lengths = [&](int thresh, int& tot, int& nl) {
        counter.zero(s);
        k_sell_width<<<wg, kThreads, 0, s>>>(rows, ns, A.ptr.p, thresh, width.p, A.long_bits.p, A.long_rows.p, A.huge_rows.p, counter.p);
        k_scan_exclusive<<<1, 1024, 0, s>>>(ns, width.p, A.sl_ptr.p);
        MK_CHECK_LAUNCH();
        int c[2] = {0, 0};
        MK_CUDA(cudaMemcpyAsync(&tot, A.sl_ptr.p + ns, sizeof(int), cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaMemcpyAsync(c, counter.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
        nl = c[0];
        nh = c[1];
        if (tot < 0 || nh > (1 << 16)) throw std::invalid_argument("build_sell: matrix too large / too many dense rows for this layout");
    };
    int tot = 0, nl = 0;
    scan_lengths(kLongRow, tot, nl);
    A.padding = A.nnz ? (double)tot / (double)A.nnz : 1.0;  // long rows count as stored 0 / true len: only makes SELL look better
    const double mean = (double)A.nnz / rows;
    // (lanes-per-row on the 387k-row level 1 of the field hierarchy, 10.8 entries per row, was measured and loses: the cycle from level 1
    // downwards takes 197 instead of 129 us, profiles/cycle_probe_r02_z_*.json; so does batching the last 1-3 columns of a SELL row into one
    // load group, 255 against 248 us per field cycle -- SELL with the 4-wide unrolled loop stays the layout of every regular operator)
    if (mean > kRowModeMean || ((A.padding > kMaxSellPadding || rows <= kRowModeSmallRows) && mean >= kRowModeMinMean)) {
        A.row_mode = true;
        A.row_lanes = mean <= 6 ? 4 : (mean <= 12 ? 8 : (mean <= 24 ? 16 : 32));
        const double pad = A.padding;
        scan_lengths(kRowModeLong, tot, nl);  // only the really long rows leave the lanes-per-row kernel
        A.n_long = nl;
        A.n_huge = nh;
        if (getenv("MK_VERBOSE") && atoi(getenv("MK_VERBOSE")) >= 2)
            fprintf(stderr, "[mk] layout: %d x %d nnz %lld mean %.1f SELL stored/true %.2f -> %d lanes per row, long rows %d\n", A.rows, A.cols,
                    (long long)A.nnz, mean, pad, A.row_lanes, nl + nh);
        A.sl_ptr.release();
        return;
    }
    A.n_long = nl;
    A.n_huge = nh;
    if (getenv("MK_VERBOSE") && atoi(getenv("MK_VERBOSE")) >= 2)
        fprintf(stderr, "[mk] layout: %d x %d nnz %lld mean %.1f SELL stored/true %.2f long rows %d + %d\n", A.rows, A.cols, (long long)A.nnz, mean,
                A.padding, nl, nh);
    A.sl_col.alloc((size_t)std::max(tot, 1));
    A.sl_val.alloc((size_t)std::max(tot, 1));
    k_sell_fill<T><<<wg, kThreads, 0, s>>>(rows, ns, A.ptr.p, A.col.p, A.val.p, A.sl_ptr.p, A.sl_col.p, A.sl_val.p);
    MK_CHECK_LAUNCH();
}

// fp32-valued copy of an operator (same pattern) with its own SpMV layout: the V-cycle's storage format
template <class T, class Tlo>
static __global__ void k_narrow_values(int64_t n, const T* __restrict__ in, Tlo* __restrict__ out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) st_rw(out + k, to_low<Tlo, T>(ld_ro(in + k), 1.0));
}
template <class T, class Tlo>
inline void lower_csr(const DeviceCsr<T>& A, DeviceCsr<Tlo>& B, cudaStream_t s) {
    B.rows = A.rows;
    B.cols = A.cols;
    B.nnz = A.nnz;
    if (!A.ptr.p) return;  // operator not materialised (level 0 behind an external fine operator)
    B.ptr.alloc((size_t)A.rows + 1);
    B.col.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    B.val.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    MK_CUDA(cudaMemcpyAsync(B.ptr.p, A.ptr.p, sizeof(int) * ((size_t)A.rows + 1), cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemcpyAsync(B.col.p, A.col.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, s));
    if (A.nnz) k_narrow_values<T, Tlo><<<stream_grid(A.nnz), kThreads, 0, s>>>(A.nnz, A.val.p, B.val.p);
    MK_CHECK_LAUNCH();
    build_sell(B, s);
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

template <>
__device__ __forceinline__ float group_sum<float>(float v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <>
__device__ __forceinline__ cfloat group_sum<cfloat>(cfloat v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// sums a T over the CTA (result valid in thread 0)
__device__ __forceinline__ double block_sum_t(double v, double* red) { return block_sum(v, red); }
__device__ __forceinline__ cdouble block_sum_t(cdouble v, double* red) {
    const double re = block_sum(v.x, red), im = block_sum(v.y, red);
    return make_c(re, im);
}
__device__ __forceinline__ float block_sum_t(float v, double* red) { return (float)block_sum((double)v, red); }
__device__ __forceinline__ cfloat block_sum_t(cfloat v, double* red) {
    const double re = block_sum((double)v.x, red), im = block_sum((double)v.y, red);
    return make_cf((float)re, (float)im);
}

// What happens to a finished row sum q = (A v)_row. One code path serves the three row shapes (SELL slice lane, long row, warp-per-row):
//   kPlain : y = q                     kAcc  : y += q (prolongation)            kDot : y = q and <v, q> (the PCG step)
//   kStep  : x (+)= v ; r -= q ; out = c1 v + c2 dinv .* r      (Chebyshev step fused into the SpMV; v = direction d, out = the next one).
//            `first`: zero initial guess, i.e. x = v and r = b - q (b is read, r need not be initialised); when x aliases v (degree-1
//            smoothing keeps the start direction in the iterate buffer) the x store is skipped. out == nullptr: last pre-smoothing step,
//            no further direction is needed.
//   kResid : r = b - q ; out = c1 dinv .* r                     (residual + first direction; v = the iterate, c1 = 1 / theta)
//   kRestrictInit : y = q ; out = c1 dinv .* q                  (restriction b_c = R r fused with the coarse level's start direction
//            d0 = dinv_c b_c / theta_c: the coarse pre-smoother needs no kernel of its own)
//   kResidLast : out = v + c1 dinv .* (b - q)                   (degree-1 post-smoothing in ONE SpMV: the step that would only have
//            updated r and the next direction -- both dead after the last step -- is not run)
//   kStepLast  : r' = r - q ; out = x + v + (c1 v + c2 dinv .* r')   (last post-smoothing step of degree >= 2, same saving)
//   k*LastDot  : the same on level 0 of the mixed-precision cycle, plus the PCG reduction <r_hi, out> against the fp64 residual:
//            rho_new = zscale * sum, beta = rho_new / rho  (what k_dot_rz / k_precond_out_dot did in a kernel of their own)
// `out` must not alias `v` (other rows still gather it). Fusing the smoother saves a kernel and ~25% of the traffic per smoothing step
// against SpMV + update.
enum SpmvMode { kPlain = 0, kAcc = 1, kDot = 2, kStep = 3, kResid = 4, kRestrictInit = 5, kResidLast = 6, kStepLast = 7, kResidLastDot = 8, kStepLastDot = 9 };

template <class T> struct HiOf { using type = T; };
template <> struct HiOf<float> { using type = double; };
template <> struct HiOf<cfloat> { using type = cdouble; };
// Re(conj(a) b) of a high-precision a and a working-precision b
__device__ __forceinline__ double hdot_hi(double a, float b) { return a * (double)b; }
__device__ __forceinline__ double hdot_hi(cdouble a, cfloat b) { return a.x * (double)b.x + a.y * (double)b.y; }
__device__ __forceinline__ double hdot_hi(double a, double b) { return a * b; }
__device__ __forceinline__ double hdot_hi(cdouble a, cdouble b) { return a.x * b.x + a.y * b.y; }

template <class T>
struct SpmvEpi {
    const T* v;        // the vector A is applied to
    double shift;      // (A + shift I)
    T* y;              // kPlain / kAcc / kDot / kRestrictInit
    const T* b;        // kResid / kResidLast / kStep with `first`
    const T* dinv;     // kStep / kResid / kRestrictInit / k*Last
    T* r;              // kStep / kResid / kStepLast
    T* x;              // kStep / kStepLast
    T* out;            // kStep / kResid / kRestrictInit / k*Last
    double c1, c2;
    int first;         // kStep: x = v, r = b - q
    const typename HiOf<T>::type* r_hi;  // k*LastDot: the fp64 PCG residual
};

template <class T, int MODE>
__device__ __forceinline__ double row_epilogue(const SpmvEpi<T>& e, int row, T acc) {
    double dot = 0.0;
    if (MODE == kPlain || MODE == kAcc || MODE == kDot) {
        if (MODE == kDot || e.shift != 0.0) {
            const T vr = ld_ro(e.v + row);
            if (e.shift != 0.0) acc = acc + ST<T>::scale(e.shift, vr);
            if (MODE == kDot) dot = ST<T>::hdot(vr, acc);
        }
        if (MODE == kAcc) acc = acc + ld_rw(e.y + row);
        st_rw(e.y + row, acc);
    } else if (MODE == kStep) {
        const T di = ld_ro(e.v + row);
        if (e.shift != 0.0) acc = acc + ST<T>::scale(e.shift, di);
        const T ri = (e.first ? ld_ro(e.b + row) : ld_rw(e.r + row)) - acc;
        if (!e.first) st_rw(e.x + row, ld_rw(e.x + row) + di);
        else if (e.x != e.v) st_rw(e.x + row, di);
        st_rw(e.r + row, ri);
        if (e.out) st_rw(e.out + row, ST<T>::scale(e.c1, di) + ST<T>::scale(e.c2, ld_ro(e.dinv + row) * ri));
    } else if (MODE == kResid) {
        if (e.shift != 0.0) acc = acc + ST<T>::scale(e.shift, ld_ro(e.v + row));
        const T ri = ld_ro(e.b + row) - acc;
        st_rw(e.r + row, ri);
        st_rw(e.out + row, ST<T>::scale(e.c1, ld_ro(e.dinv + row) * ri));
    } else if (MODE == kRestrictInit) {
        st_rw(e.y + row, acc);
        st_rw(e.out + row, ST<T>::scale(e.c1, ld_ro(e.dinv + row) * acc));
    } else if (MODE == kResidLast || MODE == kResidLastDot) {
        const T xi = ld_ro(e.v + row);
        if (e.shift != 0.0) acc = acc + ST<T>::scale(e.shift, xi);
        const T ri = ld_ro(e.b + row) - acc;
        const T xo = xi + ST<T>::scale(e.c1, ld_ro(e.dinv + row) * ri);
        st_rw(e.out + row, xo);
        if (MODE == kResidLastDot) dot = hdot_hi(ld_ro(e.r_hi + row), xo);
    } else {  // kStepLast / kStepLastDot
        const T di = ld_ro(e.v + row);
        if (e.shift != 0.0) acc = acc + ST<T>::scale(e.shift, di);
        const T ri = ld_rw(e.r + row) - acc;
        const T dn = ST<T>::scale(e.c1, di) + ST<T>::scale(e.c2, ld_ro(e.dinv + row) * ri);
        const T xo = ld_rw(e.x + row) + di + dn;
        st_rw(e.out + row, xo);
        if (MODE == kStepLastDot) dot = hdot_hi(ld_ro(e.r_hi + row), xo);
    }
    return dot;
}

template <int MODE>
__device__ __forceinline__ void finish_dot(double dot, PcgState* st, ReduceWs ws, double* red) {
    if (MODE == kDot) {
        double v[1] = {dot};
        grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
            st->pq = s[0];
            st->alpha = st->rho / s[0];
            if (!(s[0] > 0.0)) st->done = 2;  // not positive definite / breakdown
        });
    }
    if (MODE == kResidLastDot || MODE == kStepLastDot) {
        double v[1] = {dot};
        grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
            const double rho_new = st->zscale * s[0];   // z = zscale * (fp32 cycle output)
            st->beta = rho_new / st->rho;
            st->rho = rho_new;
        });
    }
}
constexpr bool spmv_mode_reduces(int mode) { return mode == kDot || mode == kResidLastDot || mode == kStepLastDot; }

// one CTA sums a long row: four independent loads in flight per thread (a 15k-entry dense-dof row would otherwise serialise ~60 load
// round trips and set the duration of the whole kernel)
template <class T>
__device__ __forceinline__ T long_row_sum(int b, int en, const int* __restrict__ col, const T* __restrict__ val, const T* __restrict__ x, double* red) {
    T a0 = ST<T>::zero(), a1 = ST<T>::zero(), a2 = ST<T>::zero(), a3 = ST<T>::zero();
    const int bs = blockDim.x;
    int k = b + threadIdx.x;
    if constexpr (sizeof(T) > 8) {  // complex fp64: two streams (register budget of the kernels this is inlined into)
        for (; k + bs < en; k += 2 * bs) {
            const int c0 = __ldg(col + k), c1 = __ldg(col + k + bs);
            const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + bs);
            a0 = ST<T>::fma_(v0, ld_ro(x + c0), a0);
            a1 = ST<T>::fma_(v1, ld_ro(x + c1), a1);
        }
    }
    for (; sizeof(T) <= 8 && k + 3 * bs < en; k += 4 * bs) {
        const int c0 = __ldg(col + k), c1 = __ldg(col + k + bs), c2 = __ldg(col + k + 2 * bs), c3 = __ldg(col + k + 3 * bs);
        const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + bs), v2 = ld_ro(val + k + 2 * bs), v3 = ld_ro(val + k + 3 * bs);
        a0 = ST<T>::fma_(v0, ld_ro(x + c0), a0);
        a1 = ST<T>::fma_(v1, ld_ro(x + c1), a1);
        a2 = ST<T>::fma_(v2, ld_ro(x + c2), a2);
        a3 = ST<T>::fma_(v3, ld_ro(x + c3), a3);
    }
    for (; k < en; k += bs) a0 = ST<T>::fma_(ld_ro(val + k), ld_ro(x + __ldg(col + k)), a0);
    return block_sum_t((a0 + a1) + (a2 + a3), red);
}

// one warp sums a medium row (lanes stride along it, four loads in flight)
template <class T>
__device__ __forceinline__ T warp_row_sum(int b, int en, const int* __restrict__ col, const T* __restrict__ val, const T* __restrict__ x, int lane) {
    T a0 = ST<T>::zero(), a1 = ST<T>::zero(), a2 = ST<T>::zero(), a3 = ST<T>::zero();
    int k = b + lane;
    if constexpr (sizeof(T) > 8) {
        for (; k + 32 < en; k += 64) {
            const int c0 = __ldg(col + k), c1 = __ldg(col + k + 32);
            const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + 32);
            a0 = ST<T>::fma_(v0, ld_ro(x + c0), a0);
            a1 = ST<T>::fma_(v1, ld_ro(x + c1), a1);
        }
    }
    for (; sizeof(T) <= 8 && k + 96 < en; k += 128) {
        const int c0 = __ldg(col + k), c1 = __ldg(col + k + 32), c2 = __ldg(col + k + 64), c3 = __ldg(col + k + 96);
        const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + 32), v2 = ld_ro(val + k + 64), v3 = ld_ro(val + k + 96);
        a0 = ST<T>::fma_(v0, ld_ro(x + c0), a0);
        a1 = ST<T>::fma_(v1, ld_ro(x + c1), a1);
        a2 = ST<T>::fma_(v2, ld_ro(x + c2), a2);
        a3 = ST<T>::fma_(v3, ld_ro(x + c3), a3);
    }
    for (; k < en; k += 32) a0 = ST<T>::fma_(ld_ro(val + k), ld_ro(x + __ldg(col + k)), a0);
    return group_sum<T>((a0 + a1) + (a2 + a3), 32);
}

// The rows that are not served by the bulk part of an SpMV launch, in the FIRST CTAs of the grid (they are the longest-running ones):
// CTAs [0, n_huge): one huge row each; CTAs [n_huge, n_huge + ceil(n_long / 8)): one medium row per warp. Returns the dot contribution
// of the calling thread; `extra` = number of such CTAs.
template <class T, int MODE>
__device__ __forceinline__ double extra_rows(int n_huge, const int* __restrict__ huge_rows, int n_long, const int* __restrict__ long_rows,
                                             const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val,
                                             const SpmvEpi<T>& e, double* red) {
    const T* __restrict__ x = e.v;
    double dot = 0.0;
    if ((int)blockIdx.x < n_huge) {
        const int row = huge_rows[blockIdx.x];
        const T acc = long_row_sum<T>(ptr[row], ptr[row + 1], col, val, x, red);
        if (threadIdx.x == 0) dot = row_epilogue<T, MODE>(e, row, acc);
    } else {
        const int idx = ((int)blockIdx.x - n_huge) * (kThreads / 32) + (threadIdx.x >> 5);
        if (idx < n_long) {
            const int row = long_rows[idx];
            const T acc = warp_row_sum<T>(ptr[row], ptr[row + 1], col, val, x, threadIdx.x & 31);
            if ((threadIdx.x & 31) == 0) dot = row_epilogue<T, MODE>(e, row, acc);
        }
    }
    return dot;
}
inline int extra_blocks(int n_huge, int n_long) { return n_huge + (n_long + kThreads / 32 - 1) / (kThreads / 32); }

// SELL-32 SpMV: the CTAs after the `extra` ones run one thread per row, one warp per slice; rows longer than kLongRow are empty in the
// slices and served by the extra CTAs of the SAME launch (extra_rows) -- no second kernel, no dependency, and no slice is serialised by
// one wide row.
// Occupancy: the real fp32 kernels (every SpMV of the IGM V-cycle) fit 32 registers without spilling and run 8 CTAs per SM -- the cycle on the
// 7-entries-per-row operator is gather-latency bound and gains 12 % (285 -> 253 us per application at 1M vertices); cfloat / double / cdouble
// spill at 32 registers and measure slower (field cycle 247 -> 251 us), so they stay at 6 CTAs per SM (40 registers).
template <class T, int MODE>
__global__ void __launch_bounds__(kThreads, sizeof(T) <= 4 ? 8 : 6) k_spmv_sell(int n, int n_slices, const int* __restrict__ sl_ptr, const int* __restrict__ sl_col,
                                                        const T* __restrict__ sl_val, const unsigned int* __restrict__ long_bits, int extra,
                                                        int n_huge, const int* __restrict__ huge_rows, int n_long, const int* __restrict__ long_rows,
                                                        const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val,
                                                        SpmvEpi<T> e, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st && st->done) return;
    const T* __restrict__ x = e.v;
    double dot = 0.0;
    if ((int)blockIdx.x < extra) {
        dot = extra_rows<T, MODE>(n_huge, huge_rows, n_long, long_rows, ptr, col, val, e, red);
    } else {
        const int lane = threadIdx.x & 31;
        const int sell_blocks = gridDim.x - extra;
        const int nwarps = (sell_blocks * blockDim.x) >> 5;
        for (int sl = (((int)blockIdx.x - extra) * blockDim.x + threadIdx.x) >> 5; sl < n_slices; sl += nwarps) {
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
            if (row < n && !(long_bits != nullptr && ((__ldg(long_bits + sl) >> lane) & 1u))) dot += row_epilogue<T, MODE>(e, row, acc);
        }
    }
    finish_dot<MODE>(dot, st, ws, red);
}

// Lanes-per-row SpMV on the plain CSR arrays for operators with irregular row lengths (the coarse IGM levels): G lanes stride along
// one row (no padding, coalesced within the row); rows longer than kRowModeLong are served by the extra CTAs (extra_rows).
template <class T, int MODE, int G>
__global__ void __launch_bounds__(kThreads) k_spmv_rows(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val,
                                                        const unsigned int* __restrict__ long_bits, int extra, int n_huge,
                                                        const int* __restrict__ huge_rows, int n_long, const int* __restrict__ long_rows,
                                                        SpmvEpi<T> e, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st && st->done) return;
    const T* __restrict__ x = e.v;
    double dot = 0.0;
    if ((int)blockIdx.x < extra) {
        dot = extra_rows<T, MODE>(n_huge, huge_rows, n_long, long_rows, ptr, col, val, e, red);
    } else {
        const int sub = threadIdx.x & (G - 1);
        const int row_blocks = gridDim.x - extra;
        const int ngroups = (row_blocks * blockDim.x) / G;
        const int nloop = (n + ngroups - 1) / ngroups;  // same trip count for every thread: the shuffles below need whole warps
        int row = (((int)blockIdx.x - extra) * blockDim.x + threadIdx.x) / G;
        for (int it = 0; it < nloop; it++, row += ngroups) {
            const bool live = row < n && !(long_bits != nullptr && ((__ldg(long_bits + (row >> 5)) >> (row & 31)) & 1u));
            T acc = ST<T>::zero();
            if (live) {
                const int b = __ldg(ptr + row), en = __ldg(ptr + row + 1);
                int k = b + sub;
                T acc1 = ST<T>::zero();
                for (; k + 3 * G < en; k += 4 * G) {
                    const int c0 = __ldg(col + k), c1 = __ldg(col + k + G), c2 = __ldg(col + k + 2 * G), c3 = __ldg(col + k + 3 * G);
                    const T v0 = ld_ro(val + k), v1 = ld_ro(val + k + G), v2 = ld_ro(val + k + 2 * G), v3 = ld_ro(val + k + 3 * G);
                    acc = ST<T>::fma_(v0, ld_ro(x + c0), acc);
                    acc1 = ST<T>::fma_(v1, ld_ro(x + c1), acc1);
                    acc = ST<T>::fma_(v2, ld_ro(x + c2), acc);
                    acc1 = ST<T>::fma_(v3, ld_ro(x + c3), acc1);
                }
                for (; k < en; k += G) acc = ST<T>::fma_(ld_ro(val + k), ld_ro(x + __ldg(col + k)), acc);
                acc = acc + acc1;
            }
            acc = group_sum<T>(acc, G);
            if (live && sub == 0) dot += row_epilogue<T, MODE>(e, row, acc);
        }
    }
    finish_dot<MODE>(dot, st, ws, red);
}

template <class T, int MODE, int G>
inline void spmv_rows_launch(const DeviceCsr<T>& A, const SpmvEpi<T>& e, PcgState* st, ReduceWs ws, cudaStream_t s) {
    const int extra = extra_blocks(A.n_huge, A.n_long);
    int row_blocks = stream_grid((int64_t)A.rows * G, kThreads, 16);
    if (spmv_mode_reduces(MODE) && row_blocks + extra > kMaxPartials) row_blocks = std::max(1, kMaxPartials - extra);
    k_spmv_rows<T, MODE, G><<<row_blocks + extra, kThreads, 0, s>>>(A.rows, A.ptr.p, A.col.p, A.val.p, extra ? A.long_bits.p : nullptr, extra, A.n_huge,
                                                                    A.huge_rows.p, A.n_long, A.long_rows.p, e, st, ws);
}

template <class T, int MODE>
inline void spmv_launch(const DeviceCsr<T>& A, const SpmvEpi<T>& e, PcgState* st, ReduceWs ws, cudaStream_t s) {
    if (A.row_mode) {
        switch (A.row_lanes) {
            case 4: spmv_rows_launch<T, MODE, 4>(A, e, st, ws, s); break;
            case 8: spmv_rows_launch<T, MODE, 8>(A, e, st, ws, s); break;
            case 16: spmv_rows_launch<T, MODE, 16>(A, e, st, ws, s); break;
            default: spmv_rows_launch<T, MODE, 32>(A, e, st, ws, s); break;
        }
    } else {
        const int extra = extra_blocks(A.n_huge, A.n_long);
        int sell_blocks = stream_grid((int64_t)A.n_slices * 32, kThreads, 16);
        if (spmv_mode_reduces(MODE) && sell_blocks + extra > kMaxPartials) sell_blocks = std::max(1, kMaxPartials - extra);
        k_spmv_sell<T, MODE><<<sell_blocks + extra, kThreads, 0, s>>>(A.rows, A.n_slices, A.sl_ptr.p, A.sl_col.p, A.sl_val.p,
                                                                     extra ? A.long_bits.p : nullptr, extra, A.n_huge, A.huge_rows.p, A.n_long,
                                                                     A.long_rows.p, A.ptr.p, A.col.p, A.val.p, e, st, ws);
    }
    MK_CHECK_LAUNCH();
}

// pq = <x, y>, alpha = rho / pq as a kernel of its own: only for operators with so many long rows that the fused-dot launch would
// need more CTAs than the reduction workspace holds partials for
template <class T>
__global__ void __launch_bounds__(kThreads) k_dot_alpha(int n, const T* __restrict__ x, const T* __restrict__ y, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += ST<T>::hdot(ld_ro(x + i), ld_ro(y + i));
    finish_dot<kDot>(acc, st, ws, red);
}

// y (+)= A x (+ shift x); st != nullptr: skipped once st->done is set; dot: fused <x, y> reduction (needs st)
template <class T>
inline void spmv(const DeviceCsr<T>& A, const T* x, T* y, double shift, bool accumulate, bool dot, PcgState* st, ReduceWs ws, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = x; e.shift = shift; e.y = y;
    if (dot && st && extra_blocks(A.n_huge, A.n_long) > kMaxPartials / 2) {
        spmv_launch<T, kPlain>(A, e, st, ws, s);
        k_dot_alpha<T><<<stream_grid(A.rows), kThreads, 0, s>>>(A.rows, x, y, st, ws);
        MK_CHECK_LAUNCH();
        return;
    }
    if (dot && st) spmv_launch<T, kDot>(A, e, st, ws, s);
    else if (accumulate) spmv_launch<T, kAcc>(A, e, st, ws, s);
    else spmv_launch<T, kPlain>(A, e, st, ws, s);
}

template <class T>
inline void spmv_cheb_step(const DeviceCsr<T>& A, double shift, const T* d_in, T* d_out, T* r, T* x, const T* b, const T* dinv, double c1, double c2,
                           int first, const PcgState* st, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = d_in; e.shift = shift; e.dinv = dinv; e.r = r; e.x = x; e.b = b; e.out = d_out; e.c1 = c1; e.c2 = c2; e.first = first;
    spmv_launch<T, kStep>(A, e, const_cast<PcgState*>(st), ReduceWs{nullptr, nullptr}, s);
}

template <class T>
inline void spmv_cheb_resid(const DeviceCsr<T>& A, double shift, const T* x, const T* b, T* r, T* d_out, const T* dinv, double inv_theta,
                            const PcgState* st, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = x; e.shift = shift; e.b = b; e.dinv = dinv; e.r = r; e.out = d_out; e.c1 = inv_theta;
    spmv_launch<T, kResid>(A, e, const_cast<PcgState*>(st), ReduceWs{nullptr, nullptr}, s);
}

// y = A x ; out = c dinv .* y   (restriction fused with the coarse level's Chebyshev start direction)
template <class T>
inline void spmv_restrict_init(const DeviceCsr<T>& A, const T* x, T* y, T* out, const T* dinv, double c, const PcgState* st, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = x; e.y = y; e.out = out; e.dinv = dinv; e.c1 = c;
    spmv_launch<T, kRestrictInit>(A, e, const_cast<PcgState*>(st), ReduceWs{nullptr, nullptr}, s);
}

// xo = x + inv_theta dinv .* (b - A x); r_hi != nullptr: also the PCG reduction <r_hi, xo> (needs st and ws)
template <class T>
inline void spmv_cheb_resid_last(const DeviceCsr<T>& A, double shift, const T* x, const T* b, T* xo, const T* dinv, double inv_theta,
                                 const typename HiOf<T>::type* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = x; e.shift = shift; e.b = b; e.dinv = dinv; e.out = xo; e.c1 = inv_theta; e.r_hi = r_hi;
    if (r_hi) spmv_launch<T, kResidLastDot>(A, e, st, ws, s);
    else spmv_launch<T, kResidLast>(A, e, st, ReduceWs{nullptr, nullptr}, s);
}

// xo = x + d + (c1 d + c2 dinv .* (r - A d))
template <class T>
inline void spmv_cheb_step_last(const DeviceCsr<T>& A, double shift, const T* d_in, T* r, T* x, T* xo, const T* dinv, double c1, double c2,
                                const typename HiOf<T>::type* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
    SpmvEpi<T> e{};
    e.v = d_in; e.shift = shift; e.dinv = dinv; e.r = r; e.x = x; e.out = xo; e.c1 = c1; e.c2 = c2; e.r_hi = r_hi;
    if (r_hi) spmv_launch<T, kStepLastDot>(A, e, st, ws, s);
    else spmv_launch<T, kStepLast>(A, e, st, ReduceWs{nullptr, nullptr}, s);
}

template <class T>
struct CsrOp : LinOp<T> {
    using Hi = typename HiOf<T>::type;
    const DeviceCsr<T>* A = nullptr;
    double shift = 0.0;
    explicit CsrOp(const DeviceCsr<T>* A_, double shift_ = 0.0) : A(A_), shift(shift_) { this->n = A_->rows; }
    int launches_per_apply() const override { return 1; }
    void apply(const T* x, T* y, PcgState* st, ReduceWs ws, cudaStream_t s, bool dot = true) override {
        spmv<T>(*A, x, y, shift, false, dot && st != nullptr, st, ws, s);
    }
    void cheb_step(const T* d_in, T* d_out, T* r, T* x, const T* b, const T* dinv, double c1, double c2, int first, const PcgState* st,
                   cudaStream_t s) override {
        spmv_cheb_step<T>(*A, shift, d_in, d_out, r, x, b, dinv, c1, c2, first, st, s);
    }
    void cheb_resid(const T* x, const T* b, T* r, T* d_out, const T* dinv, double inv_theta, const PcgState* st, cudaStream_t s) override {
        spmv_cheb_resid<T>(*A, shift, x, b, r, d_out, dinv, inv_theta, st, s);
    }
    void cheb_resid_last(const T* x, const T* b, T* xo, const T* dinv, double inv_theta, const void* r_hi, PcgState* st, ReduceWs ws,
                         cudaStream_t s) override {
        spmv_cheb_resid_last<T>(*A, shift, x, b, xo, dinv, inv_theta, (const Hi*)r_hi, st, ws, s);
    }
    void cheb_step_last(const T* d_in, T* r, T* x, T* xo, const T* dinv, double c1, double c2, const void* r_hi, PcgState* st, ReduceWs ws,
                        cudaStream_t s) override {
        spmv_cheb_step_last<T>(*A, shift, d_in, r, x, xo, dinv, c1, c2, (const Hi*)r_hi, st, ws, s);
    }
};

}  // namespace mk
