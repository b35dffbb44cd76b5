// Shared device/host helpers for libmetriko_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace mk {

#ifndef MK_NVTX_RANGE_DEFINED
#define MK_NVTX_RANGE_DEFINED
// NVTX range (header-only nvtx3: a no-op unless a profiler injected its library); see hostcsr.hpp
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#endif

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define MK_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw mk::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                std::to_string(__LINE__) + ")");                                        \
    } while (0)

#define MK_CHECK_LAUNCH() MK_CUDA(cudaGetLastError())

constexpr int kSMs = 148;         // B200: 2 dies x 74 SMs
constexpr int kThreads = 256;     // default CTA size for the streaming kernels
constexpr int kMaxPartials = 4096;

// grid for a grid-stride streaming kernel: a multiple of the SM count, capped by the work
inline int stream_grid(int64_t n, int per_block = kThreads, int waves = 8) {
    int64_t need = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)kSMs * waves;
    int64_t g = need < cap ? need : cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- complex fp64 as double2 (interleaved re, im: the layout of std::complex / Eigen SprsC) ----
struct cdouble {
    double x, y;
};
__host__ __device__ inline cdouble make_c(double a, double b) { cdouble r; r.x = a; r.y = b; return r; }
__host__ __device__ inline cdouble operator+(cdouble a, cdouble b) { return make_c(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cdouble operator-(cdouble a, cdouble b) { return make_c(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cdouble operator*(cdouble a, cdouble b) { return make_c(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ inline cdouble operator*(double s, cdouble a) { return make_c(s * a.x, s * a.y); }
__host__ __device__ inline cdouble operator*(cdouble a, double s) { return make_c(s * a.x, s * a.y); }
__host__ __device__ inline cdouble conj(cdouble a) { return make_c(a.x, -a.y); }
__host__ __device__ inline cdouble& operator+=(cdouble& a, cdouble b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ inline cdouble& operator-=(cdouble& a, cdouble b) { a.x -= b.x; a.y -= b.y; return a; }

// scalar traits so that the vector / SpMV kernels are written once for real and complex fp64
template <class T> struct ST;
template <> struct ST<double> {
    __host__ __device__ static double zero() { return 0.0; }
    __host__ __device__ static double hdot(double a, double b) { return a * b; }         // Re(conj(a) b)
    __host__ __device__ static double abs2(double a) { return a * a; }
    __host__ __device__ static double fma_(double a, double b, double c) { return a * b + c; }
    __host__ __device__ static double scale(double s, double a) { return s * a; }
    __host__ __device__ static double cj(double a) { return a; }
};
template <> struct ST<cdouble> {
    __host__ __device__ static cdouble zero() { return make_c(0.0, 0.0); }
    __host__ __device__ static double hdot(cdouble a, cdouble b) { return a.x * b.x + a.y * b.y; }
    __host__ __device__ static double abs2(cdouble a) { return a.x * a.x + a.y * a.y; }
    __host__ __device__ static cdouble fma_(cdouble a, cdouble b, cdouble c) {
        return make_c(c.x + a.x * b.x - a.y * b.y, c.y + a.x * b.y + a.y * b.x);
    }
    __host__ __device__ static cdouble scale(double s, cdouble a) { return make_c(s * a.x, s * a.y); }
    __host__ __device__ static cdouble cj(cdouble a) { return make_c(a.x, -a.y); }
};

// ---- fp32 storage types of the multigrid V-cycle (the preconditioner only: PCG, residuals and every result stay fp64) ----
struct cfloat {
    float x, y;
};
__host__ __device__ inline cfloat make_cf(float a, float b) { cfloat r; r.x = a; r.y = b; return r; }
__host__ __device__ inline cfloat operator+(cfloat a, cfloat b) { return make_cf(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cfloat operator-(cfloat a, cfloat b) { return make_cf(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cfloat operator*(cfloat a, cfloat b) { return make_cf(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <> struct ST<float> {
    __host__ __device__ static float zero() { return 0.0f; }
    __host__ __device__ static double hdot(float a, float b) { return (double)a * (double)b; }
    __host__ __device__ static double abs2(float a) { return (double)a * (double)a; }
    __host__ __device__ static float fma_(float a, float b, float c) { return a * b + c; }
    __host__ __device__ static float scale(double s, float a) { return (float)s * a; }
    __host__ __device__ static float cj(float a) { return a; }
};
template <> struct ST<cfloat> {
    __host__ __device__ static cfloat zero() { return make_cf(0.0f, 0.0f); }
    __host__ __device__ static double hdot(cfloat a, cfloat b) { return (double)a.x * b.x + (double)a.y * b.y; }
    __host__ __device__ static double abs2(cfloat a) { return (double)a.x * a.x + (double)a.y * a.y; }
    __host__ __device__ static cfloat fma_(cfloat a, cfloat b, cfloat c) { return make_cf(c.x + a.x * b.x - a.y * b.y, c.y + a.x * b.y + a.y * b.x); }
    __host__ __device__ static cfloat scale(double s, cfloat a) { return make_cf((float)s * a.x, (float)s * a.y); }
    __host__ __device__ static cfloat cj(cfloat a) { return make_cf(a.x, -a.y); }
};
// narrowing / widening between the fp64 solver types and their fp32 storage types, with a real scale factor
template <class Tlo, class T> __device__ __forceinline__ Tlo to_low(T v, double s);
template <> __device__ __forceinline__ float to_low<float, double>(double v, double s) { return (float)(v * s); }
template <> __device__ __forceinline__ cfloat to_low<cfloat, cdouble>(cdouble v, double s) { return make_cf((float)(v.x * s), (float)(v.y * s)); }
template <class T, class Tlo> __device__ __forceinline__ T to_high(Tlo v, double s);
template <> __device__ __forceinline__ double to_high<double, float>(float v, double s) { return (double)v * s; }
template <> __device__ __forceinline__ cdouble to_high<cdouble, cfloat>(cfloat v, double s) { return make_c((double)v.x * s, (double)v.y * s); }
template <class T> struct LowOf;
template <> struct LowOf<double> { using type = float; };
template <> struct LowOf<cdouble> { using type = cfloat; };

// 16-byte vector load/store of a cdouble (LDG.E.128 / STG.E.128)
__device__ __forceinline__ cdouble ldc(const cdouble* p) {
    double2 v = *reinterpret_cast<const double2*>(p);
    return make_c(v.x, v.y);
}
__device__ __forceinline__ cdouble ldc_nc(const cdouble* p) {
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return make_c(v.x, v.y);
}
__device__ __forceinline__ void stc(cdouble* p, cdouble v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }

template <class T> __device__ __forceinline__ T ld_ro(const T* p);
template <> __device__ __forceinline__ double ld_ro<double>(const double* p) { return __ldg(p); }
template <> __device__ __forceinline__ cdouble ld_ro<cdouble>(const cdouble* p) { return ldc_nc(p); }
template <class T> __device__ __forceinline__ T ld_rw(const T* p);
template <> __device__ __forceinline__ double ld_rw<double>(const double* p) { return *p; }
template <> __device__ __forceinline__ cdouble ld_rw<cdouble>(const cdouble* p) { return ldc(p); }
template <class T> __device__ __forceinline__ void st_rw(T* p, T v);
template <> __device__ __forceinline__ void st_rw<double>(double* p, double v) { *p = v; }
template <> __device__ __forceinline__ void st_rw<cdouble>(cdouble* p, cdouble v) { stc(p, v); }
template <> __device__ __forceinline__ float ld_ro<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ cfloat ld_ro<cfloat>(const cfloat* p) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return make_cf(v.x, v.y); }
template <> __device__ __forceinline__ float ld_rw<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ cfloat ld_rw<cfloat>(const cfloat* p) { const float2 v = *reinterpret_cast<const float2*>(p); return make_cf(v.x, v.y); }
template <> __device__ __forceinline__ void st_rw<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_rw<cfloat>(cfloat* p, cfloat v) { *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y); }

// ---- deterministic block reduction: warp shuffles, then one smem pass ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums `v` over the CTA; result valid in thread 0. `red` = smem double[32].
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Grid-wide deterministic reduction of K values per CTA: every CTA stores its partials, the last CTA to
// arrive (atomic ticket) sums them in index order and calls `fin(sums)`. No host sync, no float atomics.
template <int K, class Fin>
__device__ __forceinline__ void grid_reduce(const double (&v)[K], double* partials, unsigned int* ticket, double* red, Fin fin) {
    __shared__ bool is_last;
    double mine[K];
#pragma unroll
    for (int k = 0; k < K; k++) mine[k] = block_sum(v[k], red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) partials[(size_t)k * gridDim.x + blockIdx.x] = mine[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double a = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a += ((volatile double*)partials)[(size_t)k * gridDim.x + i];
            s[k] = block_sum(a, red);
        }
        if (threadIdx.x == 0) {
            *ticket = 0u;
            fin(s);
        }
    }
}

// The same for many values (the multi-vector dot products of the recycle space): ONE barrier pair per CTA instead of one per value --
// every warp reduces all K values by shuffles and parks them in shared memory, the first K threads add the warps' partials.
// `smem` = double[K * 32] (K values x up to 32 warps).
template <int K, class Fin>
__device__ __forceinline__ void grid_reduce_many(double (&v)[K], double* partials, unsigned int* ticket, double* smem, Fin fin) {
    __shared__ bool is_last_many;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double w = warp_sum(v[k]);
        if (lane == 0) smem[k * 32 + wid] = w;
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
        double a = 0.0;
        for (int w = 0; w < nw; w++) a += smem[threadIdx.x * 32 + w];
        partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = a;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last_many = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last_many) {
        __threadfence();
        // warp w sums value w, w + nw, ... over the CTAs in index order (deterministic)
        for (int k = wid; k < K; k += nw) {
            double a = 0.0;
            for (int i = lane; i < (int)gridDim.x; i += 32) a += ((volatile double*)partials)[(size_t)k * gridDim.x + i];
            a = warp_sum(a);
            if (lane == 0) smem[k] = a;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *ticket = 0u;
            fin(smem);
        }
    }
}

}  // namespace mk
