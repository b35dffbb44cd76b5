// Device-side sparse set-up algebra: exclusive scans, CSR transposition and SpGEMM on DeviceCsr operands.
// This is what builds, on the GPU, the operators the reference builds with Eigen on the host before it factorises:
//   E = X2F^T X2F (ir_init.h:174), dampJ^T dampJ (lm.h:184), and -- standing where the factorisation itself does -- the
//   Galerkin products P^H A P of the multigrid hierarchy (amg_dev.cuh).
//
// Design: every product is "symbolic, then numeric", and both passes are atomic-free and deterministic.
//   symbolic : a warp (or a CTA for long rows) gathers the candidate columns of one output row into shared memory,
//              bitonic-sorts them and keeps the distinct ones -> row lengths -> scan -> the CSR slots of C. This is the
//              precomputed scatter map of the assembly.
//   numeric  : one thread per output slot GATHERS its value, C(i,c) = sum_k A(i,k) B(k,c), walking row i of A in order and
//              binary-searching c in the sorted row k of B. No slot is written by two threads, the summation order is
//              fixed by the column order of A, so results are bit-reproducible run to run.
//   dense rows (the seam-translation / constraint-kernel dofs of the IGM operator: thousands of entries) use one CTA with
//              a dense flag + accumulator array over the columns of B instead.
// Row and column indices are int32; CSR rows have sorted, distinct columns on input and output.
#pragma once
#include <climits>

#include "csr.cuh"

namespace mk {

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan of n ints (out has n + 1 entries, out[n] = total; out may alias in)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanTile = 2048;  // 256 threads x 8

static __global__ void __launch_bounds__(256) k_tile_sums(int n, const int* __restrict__ in, int* __restrict__ sums) {
    __shared__ int wsum[8];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (base + j < n) s += in[base + j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += wsum[w];
        sums[blockIdx.x] = t;
    }
}

static __global__ void __launch_bounds__(256) k_tile_scan(int n, int n_tiles, const int* in, const int* __restrict__ tile_off, int* out) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
    int v[8];
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        v[j] = (base + j < n) ? in[base + j] : 0;
        s += v[j];
    }
    int incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int run = tile_off[blockIdx.x] + incl - s;
    for (int w = 0; w < wid; w++) run += wsum[w];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_off[n_tiles];
}

inline void dev_exclusive_scan(const int* in, int* out, int n, cudaStream_t s) {
    if (n <= 0) {
        MK_CUDA(cudaMemsetAsync(out, 0, sizeof(int), s));
        return;
    }
    const int nt = (n + kScanTile - 1) / kScanTile;
    DevBuf<int> sums(nt), offs(nt + 1);
    k_tile_sums<<<nt, 256, 0, s>>>(n, in, sums.p);
    k_scan_exclusive<<<1, 1024, 0, s>>>(nt, sums.p, offs.p);
    k_tile_scan<<<nt, 256, 0, s>>>(n, nt, in, offs.p, out);
    MK_CHECK_LAUNCH();
}

inline int dev_read_int(const int* p, cudaStream_t s) {
    int v = 0;
    MK_CUDA(cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// cooperative sort of an int buffer (shared or global memory) by a group of G threads: G = 32 -> one warp, otherwise the CTA
// ---------------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 32) __syncwarp();
    else __syncthreads();
}

// ascending bitonic sort of buf[0, cnt); buf must have room for the next power of two >= max(cnt, 32)
template <int G>
__device__ inline void group_sort(int* buf, int cnt, int tid) {
    int cap = 32;
    while (cap < cnt) cap <<= 1;
    for (int i = cnt + tid; i < cap; i += G) buf[i] = INT_MAX;
    group_sync<G>();
    for (int k = 2; k <= cap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (cap >> 1); t += G) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const int a = buf[i], b = buf[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { buf[i] = b; buf[l] = a; }
            }
            group_sync<G>();
        }
}

// keeps the distinct values of the sorted buf[0, cnt) at the front (in place); returns their number to every thread.
// `bcast` = one shared int (CTA version only).
template <int G>
__device__ inline int group_unique(int* buf, int cnt, int tid, int* bcast) {
    int u = 0;
    if (tid == 0) {
        for (int i = 0; i < cnt; i++) {
            const int v = buf[i];
            if (i == 0 || v != buf[u - 1]) buf[u++] = v;
        }
        if (G != 32) *bcast = u;
    }
    if (G == 32) {
        u = __shfl_sync(0xffffffffu, u, 0);
        __syncwarp();
    } else {
        __syncthreads();
        u = *bcast;
        __syncthreads();
    }
    return u;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    *total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

// position of c in the sorted range col[b, e), or -1
__device__ __forceinline__ int find_col(const int* __restrict__ col, int b, int e, int c) {
    while (b < e) {
        const int m = (b + e) >> 1;
        const int v = __ldg(col + m);
        if (v == c) return m;
        if (v < c) b = m + 1;
        else e = m;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------------------------
// row binning
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBinThread = 24;       // <= : one THREAD per row, candidates insertion-sorted in local memory (P = S T and A P of the multigrid set-up:
                                     // 7 and ~18 candidates per row on 2 M rows -- a warp per row with a shared-memory bitonic sort was 4x the time)
constexpr int kBinWarpSmall = 128;   // <= : one warp, 128-int shared buffer
constexpr int kBinWarp = 1024;       // <= : one warp, 1024-int shared buffer
constexpr int kBinCta = 8192;        // <= : one CTA, 8192-int shared buffer; above: "huge" rows
constexpr int kLongARow = 512;       // rows of A longer than this also go to the huge path (the gather costs row length x outputs)

struct RowBins {
    DevBuf<int> list;   // 5 x n row ids (bin 4 = the thread-per-row bin)
    DevBuf<int> count;  // 5
    int h_count[5] = {0, 0, 0, 0, 0};
    int n = 0;
    const int* rows(int b) const { return list.p + (size_t)b * n; }
};

static __global__ void k_bin_rows(int n, const int* __restrict__ size, int* __restrict__ list, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int u = size[i];
    const int b = u <= kBinThread ? 4 : (u <= kBinWarpSmall ? 0 : (u <= kBinWarp ? 1 : (u <= kBinCta ? 2 : 3)));
    list[(size_t)b * n + atomicAdd(count + b, 1)] = i;  // list order does not influence any result
}

inline void bin_rows(RowBins& B, int n, const int* size, cudaStream_t s) {
    B.n = n;
    B.list.alloc((size_t)5 * std::max(n, 1));
    B.count.alloc(5);
    B.count.zero(s);
    if (n > 0) k_bin_rows<<<(n + 255) / 256, 256, 0, s>>>(n, size, B.list.p, B.count.p);
    MK_CHECK_LAUNCH();
    MK_CUDA(cudaMemcpyAsync(B.h_count, B.count.p, 5 * sizeof(int), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------------------------
// SpGEMM C = A B
// ---------------------------------------------------------------------------------------------------------------

// ub[i] = sum over row i of A of the lengths of the referenced rows of B (upper bound of the row length of C)
static __global__ void k_spgemm_ub(int n, const int* __restrict__ aptr, const int* __restrict__ acol, const int* __restrict__ bptr,
                                   int* __restrict__ ub) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        int s = 0;
        for (int k = aptr[i] + lane; k < aptr[i + 1]; k += 32) {
            const int j = __ldg(acol + k);
            s += __ldg(bptr + j + 1) - __ldg(bptr + j);
        }
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (aptr[i + 1] - aptr[i] > kLongARow) s = max(s, kBinCta + 1);
        if (lane == 0) ub[i] = s;
    }
}

// gathers the candidate columns of output row i into buf (any order); returns their number. One warp.
__device__ inline int warp_gather_candidates(int i, const int* __restrict__ aptr, const int* __restrict__ acol, const int* __restrict__ bptr,
                                             const int* __restrict__ bcol, int* buf, int lane) {
    int pos = 0;
    const int a0 = aptr[i], a1 = aptr[i + 1];
    for (int base = a0; base < a1; base += 32) {
        const int k = base + lane;
        int b0 = 0, len = 0;
        if (k < a1) {
            const int j = __ldg(acol + k);
            b0 = __ldg(bptr + j);
            len = __ldg(bptr + j + 1) - b0;
        }
        int total;
        const int off = warp_excl_scan(len, lane, &total);
        for (int l = 0; l < len; l++) buf[pos + off + l] = __ldg(bcol + b0 + l);
        pos += total;
    }
    __syncwarp();
    return pos;
}

// C(i, c) for one output slot: walks row i of A in order, binary search in the rows of B
template <class T>
__device__ __forceinline__ T gather_product(int a0, int a1, const int* __restrict__ acol, const T* __restrict__ aval, const int* __restrict__ bptr,
                                            const int* __restrict__ bcol, const T* __restrict__ bval, int c) {
    T acc = ST<T>::zero();
    for (int k = a0; k < a1; k++) {
        const int j = __ldg(acol + k);
        const int p = find_col(bcol, __ldg(bptr + j), __ldg(bptr + j + 1), c);
        if (p >= 0) acc = ST<T>::fma_(ld_ro(aval + k), ld_ro(bval + p), acc);
    }
    return acc;
}

// One warp per listed row. NUMERIC = false: rownnz[i] = distinct candidate columns. NUMERIC = true: writes the sorted
// columns and the gathered values into the CSR slots cptr[i]...
template <class T, int CAP, bool NUMERIC>
static __global__ void __launch_bounds__(128) k_spgemm_warp(int n_list, const int* __restrict__ list, const int* __restrict__ aptr,
                                                            const int* __restrict__ acol, const T* __restrict__ aval, const int* __restrict__ bptr,
                                                            const int* __restrict__ bcol, const T* __restrict__ bval, int* __restrict__ rownnz,
                                                            const int* __restrict__ cptr, int* __restrict__ ccol, T* __restrict__ cval) {
    __shared__ int sbuf[4 * CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int* buf = sbuf + wid * CAP;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_list; w += nw) {
        const int i = list[w];
        const int cnt = warp_gather_candidates(i, aptr, acol, bptr, bcol, buf, lane);
        group_sort<32>(buf, cnt, lane);
        const int u = group_unique<32>(buf, cnt, lane, nullptr);
        if (!NUMERIC) {
            if (lane == 0) rownnz[i] = u;
        } else {
            const int a0 = aptr[i], a1 = aptr[i + 1], c0 = cptr[i];
            for (int e = lane; e < u; e += 32) {
                const int c = buf[e];
                ccol[c0 + e] = c;
                st_rw(cval + c0 + e, gather_product<T>(a0, a1, acol, aval, bptr, bcol, bval, c));
            }
        }
        __syncwarp();
    }
}

// One thread per listed row (<= kBinThread candidates): the candidates (column, position in A, position in B) are insertion-sorted by
// column as they are generated -- stable, so equal columns stay in the order of row i of A, which is the summation order of gather_product:
// the values are bit-identical to the other bins'.
template <class T, bool NUMERIC>
static __global__ void __launch_bounds__(128) k_spgemm_thread(int n_list, const int* __restrict__ list, const int* __restrict__ aptr,
                                                              const int* __restrict__ acol, const T* __restrict__ aval, const int* __restrict__ bptr,
                                                              const int* __restrict__ bcol, const T* __restrict__ bval, int* __restrict__ rownnz,
                                                              const int* __restrict__ cptr, int* __restrict__ ccol, T* __restrict__ cval) {
    int cc[kBinThread], ka_[kBinThread], kb_[kBinThread];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_list; w += gridDim.x * blockDim.x) {
        const int i = list[w];
        int cnt = 0;
        for (int ka = aptr[i]; ka < aptr[i + 1]; ka++) {
            const int j = __ldg(acol + ka);
            for (int kb = __ldg(bptr + j); kb < __ldg(bptr + j + 1); kb++) {
                const int c = __ldg(bcol + kb);
                int p = cnt;
                while (p > 0 && cc[p - 1] > c) { cc[p] = cc[p - 1]; ka_[p] = ka_[p - 1]; kb_[p] = kb_[p - 1]; p--; }
                cc[p] = c; ka_[p] = ka; kb_[p] = kb;
                cnt++;
            }
        }
        if (!NUMERIC) {
            int u = 0;
            for (int k = 0; k < cnt; k++) u += (k == 0 || cc[k] != cc[k - 1]) ? 1 : 0;
            rownnz[i] = u;
        } else {
            int o = cptr[i] - 1;
            T acc = ST<T>::zero();
            for (int k = 0; k < cnt; k++) {
                if (k == 0 || cc[k] != cc[k - 1]) {
                    if (k > 0) st_rw(cval + o, acc);
                    o++;
                    ccol[o] = cc[k];
                    acc = ST<T>::zero();
                }
                acc = ST<T>::fma_(ld_ro(aval + ka_[k]), ld_ro(bval + kb_[k]), acc);
            }
            if (cnt > 0) st_rw(cval + o, acc);
        }
    }
}

// One CTA (256 threads) per listed row, candidates in shared memory (<= kBinCta)
template <class T, bool NUMERIC>
static __global__ void __launch_bounds__(256) k_spgemm_cta(int n_list, const int* __restrict__ list, const int* __restrict__ aptr,
                                                           const int* __restrict__ acol, const T* __restrict__ aval, const int* __restrict__ bptr,
                                                           const int* __restrict__ bcol, const T* __restrict__ bval, int* __restrict__ rownnz,
                                                           const int* __restrict__ cptr, int* __restrict__ ccol, T* __restrict__ cval) {
    __shared__ int buf[kBinCta];
    __shared__ int s_pos, s_u;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int w = blockIdx.x; w < n_list; w += gridDim.x) {
        const int i = list[w];
        const int a0 = aptr[i], a1 = aptr[i + 1];
        if (threadIdx.x == 0) s_pos = 0;
        __syncthreads();
        for (int base = a0 + wid * 32; base < a1; base += nwarp * 32) {
            const int k = base + lane;
            int b0 = 0, len = 0;
            if (k < a1) {
                const int j = __ldg(acol + k);
                b0 = __ldg(bptr + j);
                len = __ldg(bptr + j + 1) - b0;
            }
            int total;
            const int off = warp_excl_scan(len, lane, &total);
            int start = 0;
            if (lane == 0) start = atomicAdd(&s_pos, total);  // placement order is irrelevant: the buffer is sorted next
            start = __shfl_sync(0xffffffffu, start, 0);
            for (int l = 0; l < len; l++) buf[start + off + l] = __ldg(bcol + b0 + l);
        }
        __syncthreads();
        const int cnt = s_pos;
        group_sort<256>(buf, cnt, threadIdx.x);
        const int u = group_unique<256>(buf, cnt, threadIdx.x, &s_u);
        if (!NUMERIC) {
            if (threadIdx.x == 0) rownnz[i] = u;
        } else {
            const int c0 = cptr[i];
            for (int e = threadIdx.x; e < u; e += blockDim.x) {
                const int c = buf[e];
                ccol[c0 + e] = c;
                st_rw(cval + c0 + e, gather_product<T>(a0, a1, acol, aval, bptr, bcol, bval, c));
            }
        }
        __syncthreads();
    }
}

// Huge rows: one CTA (1024 threads) per row with a dense flag array and a dense accumulator over the columns of B
// (flag / acc: n_list x bcols, zero on entry of the symbolic pass; the numeric pass reuses the flags).
// The accumulation walks row i of A in order (chunks staged through shared memory); within one k the columns of row k of B
// are distinct, so threads never collide and the summation order is fixed.
// marks the columns reached from row i of A in the dense flag array
__device__ __forceinline__ void huge_mark(int a0, int a1, const int* __restrict__ acol, const int* __restrict__ bptr, const int* __restrict__ bcol, int* fl) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int k = a0 + wid; k < a1; k += nwarp) {
        const int j = __ldg(acol + k);
        for (int l = __ldg(bptr + j) + lane; l < __ldg(bptr + j + 1); l += 32) fl[__ldg(bcol + l)] = 1;
    }
    __syncthreads();
}

static __global__ void __launch_bounds__(1024) k_spgemm_huge_sym(int n_list, const int* __restrict__ list, const int* __restrict__ aptr,
                                                                 const int* __restrict__ acol, const int* __restrict__ bptr, const int* __restrict__ bcol,
                                                                 int bcols, int* __restrict__ flag, int* __restrict__ rownnz) {
    __shared__ int red[32];
    const int w = blockIdx.x;
    if (w >= n_list) return;
    const int i = list[w];
    int* fl = flag + (size_t)w * bcols;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    huge_mark(aptr[i], aptr[i + 1], acol, bptr, bcol, fl);
    int cnt = 0;
    for (int c = threadIdx.x; c < bcols; c += blockDim.x)
        if (fl[c]) { cnt++; fl[c] = 0; }  // leaves the buffer clean for the next batch
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) red[wid] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int q = 0; q < nwarp; q++) t += red[q];
        rownnz[i] = t;
    }
}

template <class T>
static __global__ void __launch_bounds__(1024) k_spgemm_huge_num(int n_list, const int* __restrict__ list, const int* __restrict__ aptr,
                                                                 const int* __restrict__ acol, const T* __restrict__ aval, const int* __restrict__ bptr,
                                                                 const int* __restrict__ bcol, const T* __restrict__ bval, int bcols,
                                                                 int* __restrict__ flag, T* __restrict__ acc, const int* __restrict__ cptr,
                                                                 int* __restrict__ ccol, T* __restrict__ cval) {
    __shared__ int s_b0[1024], s_len[1024];
    __shared__ T s_a[1024];
    __shared__ int wtot[32];
    __shared__ int s_carry;
    const int w = blockIdx.x;
    if (w >= n_list) return;
    const int i = list[w];
    int* fl = flag + (size_t)w * bcols;
    T* ac = acc + (size_t)w * bcols;
    const int a0 = aptr[i], a1 = aptr[i + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    huge_mark(a0, a1, acol, bptr, bcol, fl);
    for (int base = a0; base < a1; base += 1024) {
        const int k = base + threadIdx.x;
        if (k < a1) {
            const int j = __ldg(acol + k);
            const int b0 = __ldg(bptr + j);
            s_b0[threadIdx.x] = b0;
            s_len[threadIdx.x] = __ldg(bptr + j + 1) - b0;
            s_a[threadIdx.x] = ld_ro(aval + k);
        }
        __syncthreads();
        const int m = min(1024, a1 - base);
        for (int q = 0; q < m; q++) {
            const int b0 = s_b0[q], len = s_len[q];
            const T a = s_a[q];
            for (int l = threadIdx.x; l < len; l += 1024) {
                const int c = __ldg(bcol + b0 + l);
                st_rw(ac + c, ST<T>::fma_(a, ld_ro(bval + b0 + l), ld_rw(ac + c)));
            }
            if (len > 0) __syncthreads();  // uniform: every thread sees the same len
        }
        __syncthreads();
    }
    // ordered compaction of the flagged columns into the CSR slots; flags and accumulators are reset on the way
    if (threadIdx.x == 0) s_carry = cptr[i];
    __syncthreads();
    for (int base = 0; base < bcols; base += 1024) {
        const int c = base + threadIdx.x;
        const int f = (c < bcols) ? fl[c] : 0;
        int total;
        const int off = warp_excl_scan(f, lane, &total);
        if (lane == 0) wtot[wid] = total;
        __syncthreads();
        int before = s_carry;
        for (int q = 0; q < wid; q++) before += wtot[q];
        if (f) {
            ccol[before + off] = c;
            st_rw(cval + before + off, ld_rw(ac + c));
            fl[c] = 0;
            st_rw(ac + c, ST<T>::zero());
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < 32; q++) t += wtot[q];
            s_carry += t;
        }
        __syncthreads();
    }
}

template <class T>
inline void spgemm(const DeviceCsr<T>& A, const DeviceCsr<T>& B, DeviceCsr<T>& C, cudaStream_t s) {
    if (A.cols != B.rows) throw std::invalid_argument("spgemm: shape mismatch");
    const int n = A.rows;
    C.rows = n;
    C.cols = B.cols;
    DevBuf<int> ub(std::max(n, 1));
    k_spgemm_ub<<<stream_grid((int64_t)n * 32), kThreads, 0, s>>>(n, A.ptr.p, A.col.p, B.ptr.p, ub.p);
    MK_CHECK_LAUNCH();
    RowBins bins;
    bin_rows(bins, n, ub.p, s);
    static const bool dbg = getenv("MK_VERBOSE") && atoi(getenv("MK_VERBOSE")) >= 2;
    StageTimer tm("spgemm");
    tm.on = dbg;
    if (dbg) fprintf(stderr, "[mk] spgemm: %d x %d x %d, bins %d | %d %d %d %d\n", n, A.cols, B.cols, bins.h_count[4], bins.h_count[0], bins.h_count[1], bins.h_count[2], bins.h_count[3]);
    DevBuf<int> rownnz((size_t)n + 1);
    rownnz.zero(s);
    const int nh = bins.h_count[3];
    // huge rows run in batches that share one self-cleaning flag / accumulator buffer (<= ~128 MB: large transient blocks make the
    // stream-ordered pool grow, which costs driver time)
    const int hb = nh ? (int)std::max<int64_t>(1, std::min<int64_t>(nh, (int64_t)(1 << 27) / ((int64_t)std::max(B.cols, 1) * (4 + (int64_t)sizeof(T))))) : 0;
    DevBuf<int> flag;
    DevBuf<T> acc;
    if (nh) {
        flag.alloc((size_t)hb * B.cols);
        acc.alloc((size_t)hb * B.cols);
        flag.zero(s);
        acc.zero(s);
    }
    auto warp_grid = [](int cnt) { return std::max(1, std::min((cnt + 3) / 4, kSMs * 16)); };
    auto cta_grid = [](int cnt) { return std::max(1, std::min(cnt, kSMs * 4)); };
    auto thread_grid = [](int cnt) { return std::max(1, std::min((cnt + 127) / 128, kSMs * 32)); };
    // symbolic
    if (bins.h_count[4])
        k_spgemm_thread<T, false><<<thread_grid(bins.h_count[4]), 128, 0, s>>>(bins.h_count[4], bins.rows(4), A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p,
                                                                                B.val.p, rownnz.p, nullptr, nullptr, nullptr);
    if (bins.h_count[0])
        k_spgemm_warp<T, kBinWarpSmall, false><<<warp_grid(bins.h_count[0]), 128, 0, s>>>(bins.h_count[0], bins.rows(0), A.ptr.p, A.col.p, A.val.p, B.ptr.p,
                                                                                           B.col.p, B.val.p, rownnz.p, nullptr, nullptr, nullptr);
    if (bins.h_count[1])
        k_spgemm_warp<T, kBinWarp, false><<<warp_grid(bins.h_count[1]), 128, 0, s>>>(bins.h_count[1], bins.rows(1), A.ptr.p, A.col.p, A.val.p, B.ptr.p,
                                                                                      B.col.p, B.val.p, rownnz.p, nullptr, nullptr, nullptr);
    if (bins.h_count[2])
        k_spgemm_cta<T, false><<<cta_grid(bins.h_count[2]), 256, 0, s>>>(bins.h_count[2], bins.rows(2), A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p,
                                                                          B.val.p, rownnz.p, nullptr, nullptr, nullptr);
    for (int h0 = 0; h0 < nh; h0 += hb)
        k_spgemm_huge_sym<<<std::min(hb, nh - h0), 1024, 0, s>>>(std::min(hb, nh - h0), bins.rows(3) + h0, A.ptr.p, A.col.p, B.ptr.p, B.col.p, B.cols, flag.p,
                                                                  rownnz.p);
    MK_CHECK_LAUNCH();
    C.ptr.alloc((size_t)n + 1);
    dev_exclusive_scan(rownnz.p, C.ptr.p, n, s);
    const int nnz = dev_read_int(C.ptr.p + n, s);
    tm.lap("symbolic");
    if (nnz < 0) throw std::invalid_argument("spgemm: product too large for int32 CSR");
    C.nnz = nnz;
    C.col.alloc((size_t)std::max(nnz, 1));
    C.val.alloc((size_t)std::max(nnz, 1));
    // numeric
    if (bins.h_count[4])
        k_spgemm_thread<T, true><<<thread_grid(bins.h_count[4]), 128, 0, s>>>(bins.h_count[4], bins.rows(4), A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p,
                                                                               B.val.p, nullptr, C.ptr.p, C.col.p, C.val.p);
    if (bins.h_count[0])
        k_spgemm_warp<T, kBinWarpSmall, true><<<warp_grid(bins.h_count[0]), 128, 0, s>>>(bins.h_count[0], bins.rows(0), A.ptr.p, A.col.p, A.val.p, B.ptr.p,
                                                                                          B.col.p, B.val.p, nullptr, C.ptr.p, C.col.p, C.val.p);
    if (bins.h_count[1])
        k_spgemm_warp<T, kBinWarp, true><<<warp_grid(bins.h_count[1]), 128, 0, s>>>(bins.h_count[1], bins.rows(1), A.ptr.p, A.col.p, A.val.p, B.ptr.p,
                                                                                     B.col.p, B.val.p, nullptr, C.ptr.p, C.col.p, C.val.p);
    if (bins.h_count[2])
        k_spgemm_cta<T, true><<<cta_grid(bins.h_count[2]), 256, 0, s>>>(bins.h_count[2], bins.rows(2), A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p,
                                                                         B.val.p, nullptr, C.ptr.p, C.col.p, C.val.p);
    for (int h0 = 0; h0 < nh; h0 += hb)
        k_spgemm_huge_num<T><<<std::min(hb, nh - h0), 1024, 0, s>>>(std::min(hb, nh - h0), bins.rows(3) + h0, A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p,
                                                                     B.val.p, B.cols, flag.p, acc.p, C.ptr.p, C.col.p, C.val.p);
    MK_CHECK_LAUNCH();
    MK_CUDA(cudaStreamSynchronize(s));  // temporaries (bins, flags) are released on return
    tm.lap("numeric");
}

// ---------------------------------------------------------------------------------------------------------------
// transposition B = A^T (CONJ: A^H). Structure through integer atomics + a per-row sort of the row indices (so the
// result does not depend on the atomic order); values gathered by binary search in the rows of A.
// ---------------------------------------------------------------------------------------------------------------
static __global__ void k_count_cols(int64_t nnz, const int* __restrict__ col, int* __restrict__ cnt) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) atomicAdd(cnt + col[k], 1);
}

static __global__ void k_scatter_rows(int n, const int* __restrict__ aptr, const int* __restrict__ acol, const int* __restrict__ tptr,
                                      int* __restrict__ cursor, int* __restrict__ tcol) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw)
        for (int k = aptr[i] + lane; k < aptr[i + 1]; k += 32) {
            const int c = acol[k];
            tcol[tptr[c] + atomicAdd(cursor + c, 1)] = i;
        }
}

static __global__ void k_row_lengths(int n, const int* __restrict__ ptr, int* __restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = ptr[i + 1] - ptr[i];
}

template <int CAP>
static __global__ void __launch_bounds__(128) k_sort_rows_warp(int n_list, const int* __restrict__ list, const int* __restrict__ ptr, int* __restrict__ col) {
    __shared__ int sbuf[4 * CAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int* buf = sbuf + wid * CAP;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_list; w += nw) {
        const int i = list[w];
        const int b = ptr[i], cnt = ptr[i + 1] - b;
        for (int k = lane; k < cnt; k += 32) buf[k] = col[b + k];
        __syncwarp();
        group_sort<32>(buf, cnt, lane);
        for (int k = lane; k < cnt; k += 32) col[b + k] = buf[k];
        __syncwarp();
    }
}

static __global__ void __launch_bounds__(256) k_sort_rows_cta(int n_list, const int* __restrict__ list, const int* __restrict__ ptr, int* __restrict__ col) {
    __shared__ int buf[kBinCta];
    for (int w = blockIdx.x; w < n_list; w += gridDim.x) {
        const int i = list[w];
        const int b = ptr[i], cnt = ptr[i + 1] - b;
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) buf[k] = col[b + k];
        __syncthreads();
        group_sort<256>(buf, cnt, threadIdx.x);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) col[b + k] = buf[k];
        __syncthreads();
    }
}

// short rows (<= kBinThread): one thread per row, insertion sort in local memory
static __global__ void __launch_bounds__(128) k_sort_rows_thread(int n_list, const int* __restrict__ list, const int* __restrict__ ptr, int* __restrict__ col) {
    int buf[kBinThread];
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < n_list; w += gridDim.x * blockDim.x) {
        const int i = list[w];
        const int b = ptr[i], cnt = ptr[i + 1] - b;
        for (int k = 0; k < cnt; k++) {
            const int c = col[b + k];
            int p = k;
            while (p > 0 && buf[p - 1] > c) { buf[p] = buf[p - 1]; p--; }
            buf[p] = c;
        }
        for (int k = 0; k < cnt; k++) col[b + k] = buf[k];
    }
}

// huge rows: the same sort in a global scratch buffer (scratch_ptr[w] = offset of row w's power-of-two sized segment)
static __global__ void __launch_bounds__(1024) k_sort_rows_global(int n_list, const int* __restrict__ list, const int* __restrict__ ptr, int* __restrict__ col,
                                                                  const int64_t* __restrict__ scratch_ptr, int* __restrict__ scratch) {
    const int w = blockIdx.x;
    if (w >= n_list) return;
    const int i = list[w];
    const int b = ptr[i], cnt = ptr[i + 1] - b;
    int* buf = scratch + scratch_ptr[w];
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) buf[k] = col[b + k];
    __syncthreads();
    group_sort<1024>(buf, cnt, threadIdx.x);
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) col[b + k] = buf[k];
}

template <class T, bool CONJ>
static __global__ void k_transpose_values(int n_t, const int* __restrict__ tptr, const int* __restrict__ tcol, const int* __restrict__ aptr,
                                          const int* __restrict__ acol, const T* __restrict__ aval, T* __restrict__ tval) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_t; c += nw)
        for (int k = tptr[c] + lane; k < tptr[c + 1]; k += 32) {
            const int r = tcol[k];
            const int p = find_col(acol, aptr[r], aptr[r + 1], c);
            T v = p >= 0 ? ld_ro(aval + p) : ST<T>::zero();
            st_rw(tval + k, CONJ ? ST<T>::cj(v) : v);
        }
}

template <class T>
inline void transpose_dev(const DeviceCsr<T>& A, DeviceCsr<T>& B, bool conjugate, cudaStream_t s) {
    const int n = A.rows, m = A.cols;
    B.rows = m;
    B.cols = n;
    B.nnz = A.nnz;
    DevBuf<int> cnt((size_t)m + 1);
    cnt.zero(s);
    if (A.nnz) k_count_cols<<<stream_grid(A.nnz), kThreads, 0, s>>>(A.nnz, A.col.p, cnt.p);
    B.ptr.alloc((size_t)m + 1);
    dev_exclusive_scan(cnt.p, B.ptr.p, m, s);
    B.col.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    B.val.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    DevBuf<int> cursor((size_t)std::max(m, 1));
    cursor.zero(s);
    k_scatter_rows<<<stream_grid((int64_t)n * 32), kThreads, 0, s>>>(n, A.ptr.p, A.col.p, B.ptr.p, cursor.p, B.col.p);
    MK_CHECK_LAUNCH();
    // per-row sort of the row indices; cnt doubles as the row-length array
    RowBins bins;
    bin_rows(bins, m, cnt.p, s);
    auto warp_grid = [](int c) { return std::max(1, std::min((c + 3) / 4, kSMs * 16)); };
    if (bins.h_count[4])
        k_sort_rows_thread<<<std::max(1, std::min((bins.h_count[4] + 127) / 128, kSMs * 32)), 128, 0, s>>>(bins.h_count[4], bins.rows(4), B.ptr.p, B.col.p);
    if (bins.h_count[0]) k_sort_rows_warp<kBinWarpSmall><<<warp_grid(bins.h_count[0]), 128, 0, s>>>(bins.h_count[0], bins.rows(0), B.ptr.p, B.col.p);
    if (bins.h_count[1]) k_sort_rows_warp<kBinWarp><<<warp_grid(bins.h_count[1]), 128, 0, s>>>(bins.h_count[1], bins.rows(1), B.ptr.p, B.col.p);
    if (bins.h_count[2]) k_sort_rows_cta<<<std::min(bins.h_count[2], kSMs * 4), 256, 0, s>>>(bins.h_count[2], bins.rows(2), B.ptr.p, B.col.p);
    DevBuf<int> scratch;
    DevBuf<int64_t> scratch_ptr;
    if (bins.h_count[3]) {
        const int nh = bins.h_count[3];
        std::vector<int> rows(nh), hptr(2);
        MK_CUDA(cudaMemcpyAsync(rows.data(), bins.rows(3), sizeof(int) * nh, cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
        std::vector<int64_t> off(nh + 1, 0);
        for (int w = 0; w < nh; w++) {
            MK_CUDA(cudaMemcpyAsync(hptr.data(), B.ptr.p + rows[w], 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
            MK_CUDA(cudaStreamSynchronize(s));
            int64_t cap = 32;
            while (cap < hptr[1] - hptr[0]) cap <<= 1;
            off[w + 1] = off[w] + cap;
        }
        scratch.alloc((size_t)off[nh]);
        scratch_ptr.alloc(nh + 1);
        scratch_ptr.upload(off.data(), nh + 1, s);
        k_sort_rows_global<<<nh, 1024, 0, s>>>(nh, bins.rows(3), B.ptr.p, B.col.p, scratch_ptr.p, scratch.p);
        MK_CUDA(cudaStreamSynchronize(s));  // `off` is a host temporary
    }
    MK_CHECK_LAUNCH();
    if (conjugate)
        k_transpose_values<T, true><<<stream_grid((int64_t)m * 32), kThreads, 0, s>>>(m, B.ptr.p, B.col.p, A.ptr.p, A.col.p, A.val.p, B.val.p);
    else
        k_transpose_values<T, false><<<stream_grid((int64_t)m * 32), kThreads, 0, s>>>(m, B.ptr.p, B.col.p, A.ptr.p, A.col.p, A.val.p, B.val.p);
    MK_CHECK_LAUNCH();
    MK_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------------------------
// small helpers on CSR operands
// ---------------------------------------------------------------------------------------------------------------

// y = alpha * A x on the plain CSR arrays, one warp per row, lanes summed in a fixed order (set-up use: rho estimate, X^T F2)
template <class T>
static __global__ void __launch_bounds__(kThreads) k_csr_matvec(int n, const int* __restrict__ ptr, const int* __restrict__ col, const T* __restrict__ val,
                                                                const T* __restrict__ x, T* __restrict__ y, double alpha) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw) {
        T acc = ST<T>::zero();
        for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) acc = ST<T>::fma_(ld_ro(val + k), ld_ro(x + __ldg(col + k)), acc);
        acc = group_sum<T>(acc, 32);
        if (lane == 0) st_rw(y + i, ST<T>::scale(alpha, acc));
    }
}

template <class T>
inline void csr_matvec(const DeviceCsr<T>& A, const T* x, T* y, double alpha, cudaStream_t s) {
    k_csr_matvec<T><<<stream_grid((int64_t)A.rows * 32), kThreads, 0, s>>>(A.rows, A.ptr.p, A.col.p, A.val.p, x, y, alpha);
    MK_CHECK_LAUNCH();
}

template <class T>
static __global__ void k_scale_values(int64_t nnz, T* __restrict__ val, double alpha) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        st_rw(val + k, ST<T>::scale(alpha, ld_rw(val + k)));
}

// val(i, i) += shift (every row is expected to hold its diagonal)
template <class T>
static __global__ void k_shift_diagonal(int n, const int* __restrict__ ptr, const int* __restrict__ col, T* __restrict__ val, double shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = find_col(col, ptr[i], ptr[i + 1], i);
    if (p >= 0) reinterpret_cast<double*>(val + p)[0] += shift;  // real part (interleaved complex)
}

// the set-up algebra needs strictly increasing columns inside every row (binary searches); callers fall back to the host mirror otherwise
static __global__ void k_check_rows_sorted(int n, const int* __restrict__ ptr, const int* __restrict__ col, int ncols, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nw)
        for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) {
            const int c = col[k];
            if (c < 0 || c >= ncols || (k > ptr[i] && col[k - 1] >= c)) *bad = 1;
        }
}

template <class T>
inline bool rows_sorted_strict(const DeviceCsr<T>& A, cudaStream_t s) {
    DevBuf<int> bad(1);
    bad.zero(s);
    k_check_rows_sorted<<<stream_grid((int64_t)A.rows * 32), kThreads, 0, s>>>(A.rows, A.ptr.p, A.col.p, A.cols, bad.p);
    MK_CHECK_LAUNCH();
    return dev_read_int(bad.p, s) == 0;
}

static __global__ void k_narrow_ptr(int n1, const int64_t* __restrict__ in, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1) out[i] = (int)in[i];
}

template <class T>
inline void copy_csr(const DeviceCsr<T>& A, DeviceCsr<T>& B, cudaStream_t s) {
    B.rows = A.rows;
    B.cols = A.cols;
    B.nnz = A.nnz;
    B.ptr.alloc((size_t)A.rows + 1);
    B.col.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    B.val.alloc((size_t)std::max<int64_t>(A.nnz, 1));
    MK_CUDA(cudaMemcpyAsync(B.ptr.p, A.ptr.p, sizeof(int) * ((size_t)A.rows + 1), cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemcpyAsync(B.col.p, A.col.p, sizeof(int) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemcpyAsync(B.val.p, A.val.p, sizeof(T) * (size_t)A.nnz, cudaMemcpyDeviceToDevice, s));
}

// B = the listed rows of A, in list order (a rank's row block of a replicated operator, in the rank's local row order)
static __global__ void k_gather_row_lengths(int m, const int* __restrict__ rows, const int* __restrict__ aptr, int* __restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) len[i] = aptr[rows[i] + 1] - aptr[rows[i]];
}
template <class T>
static __global__ void k_gather_rows(int m, const int* __restrict__ rows, const int* __restrict__ aptr, const int* __restrict__ acol, const T* __restrict__ aval,
                                     const int* __restrict__ bptr, int* __restrict__ bcol, T* __restrict__ bval) {
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += nw) {
        const int a0 = aptr[rows[i]], len = aptr[rows[i] + 1] - a0, b0 = bptr[i];
        for (int k = lane; k < len; k += 32) { bcol[b0 + k] = acol[a0 + k]; st_rw(bval + b0 + k, ld_ro(aval + a0 + k)); }
    }
}
template <class T>
inline void csr_gather_rows(const DeviceCsr<T>& A, const int* rows_dev, int m, DeviceCsr<T>& B, cudaStream_t s) {
    DevBuf<int> len((size_t)m + 1);
    len.zero(s);
    k_gather_row_lengths<<<(m + 255) / 256, 256, 0, s>>>(m, rows_dev, A.ptr.p, len.p);
    B.rows = m;
    B.cols = A.cols;
    B.ptr.alloc((size_t)m + 1);
    dev_exclusive_scan(len.p, B.ptr.p, m, s);
    B.nnz = dev_read_int(B.ptr.p + m, s);
    B.col.alloc((size_t)std::max<int64_t>(B.nnz, 1));
    B.val.alloc((size_t)std::max<int64_t>(B.nnz, 1));
    k_gather_rows<T><<<stream_grid((int64_t)m * 32), kThreads, 0, s>>>(m, rows_dev, A.ptr.p, A.col.p, A.val.p, B.ptr.p, B.col.p, B.val.p);
    MK_CHECK_LAUNCH();
    MK_CUDA(cudaStreamSynchronize(s));
}
template <class T>
static __global__ void k_gather_values(int m, const int* __restrict__ idx, const T* __restrict__ in, T* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) st_rw(out + i, ld_ro(in + idx[i]));
}

// device CSR -> host CSR (tests, coarsest-level inverse)
template <class T, class S>
inline HostCsr<S> download_csr(const DeviceCsr<T>& A, cudaStream_t s) {
    static_assert(sizeof(T) == sizeof(S), "layout");
    HostCsr<S> H;
    H.rows = A.rows;
    H.cols = A.cols;
    H.ptr.resize((size_t)A.rows + 1);
    H.col.resize((size_t)A.nnz);
    H.val.resize((size_t)A.nnz);
    A.ptr.download(H.ptr.data(), (size_t)A.rows + 1, s);
    A.col.download(H.col.data(), (size_t)A.nnz, s);
    MK_CUDA(cudaMemcpyAsync(H.val.data(), A.val.p, sizeof(T) * (size_t)A.nnz, cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    return H;
}

}  // namespace mk
