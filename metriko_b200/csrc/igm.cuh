// IGM hot path on the device: Poisson solve (ir_init.h:173-212), NI closed form and the
// SeamlessIntegration rounding loop (ir_loop.h:29-124, ir.h:90-112) with the reference's 1e10 penalty rows replaced
// by exact constraint elimination: each rounded dof adds one orthonormalised constraint row q (rank-1 update of the
// projector P = I - Q^T Q) and the re-solve is a warm-started projected PCG on (A0 + 1e-4 I).
#pragma once
#include <cmath>
#include <unordered_map>
#include <vector>

#include "csr.cuh"
#include "hostcsr.hpp"
#include "solver.cuh"

namespace mk {

// v_S -= Q^T (Q v_S): one CTA; S = support columns of all constraint rows (small), Q is k x m row-major orthonormal.
static __global__ void __launch_bounds__(kThreads) k_project(int k, int m, const int* __restrict__ S, const double* __restrict__ Q,
                                                     double* __restrict__ v, const PcgState* st) {
    extern __shared__ double sm[];
    if (st && st->done) return;
    double* vS = sm;      // m
    double* c = sm + m;   // k
    for (int s = threadIdx.x; s < m; s += blockDim.x) vS[s] = v[S[s]];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = wid; j < k; j += nw) {
        double a = 0.0;
        for (int s = lane; s < m; s += 32) a += Q[(size_t)j * m + s] * vS[s];
        a = warp_sum(a);
        if (lane == 0) c[j] = a;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < m; s += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j < k; j++) a += c[j] * Q[(size_t)j * m + s];
        v[S[s]] = vS[s] - a;
    }
}

// Rank-1 constraint elimination for rounding step j: move x onto the hyperplane q_j . x_S = beta_j along q_j
// (q_j is orthogonal to all earlier rows, so earlier constraints stay satisfied). One CTA.
static __global__ void __launch_bounds__(kThreads) k_constraint_eliminate(int m, const int* __restrict__ S, const double* __restrict__ q, double beta,
                                                                   double* __restrict__ x) {
    __shared__ double red[32];
    __shared__ double dsh;
    double a = 0.0;
    for (int s = threadIdx.x; s < m; s += blockDim.x) a += q[s] * x[S[s]];
    a = block_sum(a, red);
    if (threadIdx.x == 0) dsh = beta - a;
    __syncthreads();
    const double d = dsh;
    for (int s = threadIdx.x; s < m; s += blockDim.x) x[S[s]] += d * q[s];
}

// out[i] = v[S[i]]
static __global__ void k_gather(int m, const int* __restrict__ S, const double* __restrict__ v, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = v[S[i]];
}

// y = a * x + b * z   (elementwise, used to form the SI right-hand sides)
static __global__ void __launch_bounds__(kThreads) k_axpby(int n, double a, const double* __restrict__ x, double b, const double* __restrict__ z,
                                                    double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + b * z[i];
}

struct ConstraintProjector : Projector<double> {
    int k = 0, m = 0;
    DevBuf<int> S;
    DevBuf<double> Q;  // capacity kmax x m
    int kmax = 0;
    void apply(double* v, const PcgState* st, cudaStream_t s) override {
        if (k == 0) return;
        k_project<<<1, kThreads, (size_t)(m + k) * sizeof(double), s>>>(k, m, S.p, Q.p, v, st);
        MK_CHECK_LAUNCH();
    }
};

// Host-side bookkeeping of the constraint rows restricted to the support S (tiny dense algebra, exact to round-off).
struct ConstraintSet {
    std::vector<int> S;
    std::unordered_map<int, int> pos;
    std::vector<std::vector<double>> Q;  // orthonormal rows
    std::vector<double> beta;
    // returns 1 if the row was independent (appended), 0 if dependent; `resid` = |c.x - v| consistency for dependent rows
    int add(const std::vector<double>& c, double v, const std::vector<double>& xS, double* resid) {
        std::vector<double> r = c;
        double b = v;
        double n0 = 0;
        for (double t : c) n0 += t * t;
        n0 = std::sqrt(n0);
        for (int pass = 0; pass < 2; pass++)
            for (size_t i = 0; i < Q.size(); i++) {
                double d = 0;
                for (size_t s = 0; s < r.size(); s++) d += r[s] * Q[i][s];
                for (size_t s = 0; s < r.size(); s++) r[s] -= d * Q[i][s];
                b -= d * beta[i];
            }
        double n1 = 0;
        for (double t : r) n1 += t * t;
        n1 = std::sqrt(n1);
        if (n1 <= 1e-9 * n0) {
            (void)xS;
            *resid = std::fabs(b);  // what is left of the target once the implied part is removed
            return 0;
        }
        for (double& t : r) t /= n1;
        Q.push_back(r);
        beta.push_back(b / n1);
        *resid = 0;
        return 1;
    }
};

}  // namespace mk
