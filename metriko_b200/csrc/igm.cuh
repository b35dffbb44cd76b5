// IGM hot path on the device: Poisson solve (ir_init.h:173-212), NI closed form and the
// SeamlessIntegration rounding loop (ir_loop.h:29-124, ir.h:90-112) with the reference's 1e10 penalty rows replaced
// by exact constraint elimination: each rounded dof adds one orthonormalised constraint row q (rank-1 update of the
// projector P = I - Q^T Q) and the re-solve is a warm-started projected PCG on (A0 + 1e-4 I).
#pragma once
#include <cmath>
#include <cstring>
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

// The same projection for large constraint sets (hundreds of singularities / seam dofs: k m beyond what one CTA should do, or m + k beyond
// its shared memory): c = Q v_S with one warp per constraint row, then v_S -= Q^T c with one thread per support column. Two launches, any size.
static __global__ void __launch_bounds__(kThreads) k_project_dots(int k, int m, const int* __restrict__ S, const double* __restrict__ Q,
                                                           const double* __restrict__ v, double* __restrict__ c, const PcgState* st) {
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= k) return;
    double a = 0.0;
    for (int s = lane; s < m; s += 32) a += Q[(size_t)j * m + s] * v[S[s]];
    a = warp_sum(a);
    if (lane == 0) c[j] = a;
}
static __global__ void __launch_bounds__(kThreads) k_project_apply(int k, int m, const int* __restrict__ S, const double* __restrict__ Q,
                                                            const double* __restrict__ c, double* __restrict__ v, const PcgState* st) {
    if (st && st->done) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= m) return;
    double a = 0.0;
    for (int j = 0; j < k; j++) a += c[j] * Q[(size_t)j * m + s];
    v[S[s]] -= a;
}

constexpr int kProjectOneCtaWork = 1 << 16;   // k * m up to which the single-CTA kernel is used (one launch, everything in shared memory)

struct ConstraintProjector : Projector<double> {
    int k = 0, m = 0;
    DevBuf<int> S;
    DevBuf<double> Q;  // capacity kmax x m
    DevBuf<double> coef;  // kmax
    int kmax = 0;
    bool one_cta() const { return (long)k * m <= kProjectOneCtaWork && (size_t)(m + k) * sizeof(double) <= 48 * 1024; }
    void apply(double* v, const PcgState* st, cudaStream_t s) override {
        if (k == 0) return;
        if (one_cta()) {
            k_project<<<1, kThreads, (size_t)(m + k) * sizeof(double), s>>>(k, m, S.p, Q.p, v, st);
        } else {
            k_project_dots<<<(k * 32 + kThreads - 1) / kThreads, kThreads, 0, s>>>(k, m, S.p, Q.p, v, coef.p, st);
            k_project_apply<<<(m + kThreads - 1) / kThreads, kThreads, 0, s>>>(k, m, S.p, Q.p, coef.p, v, st);
        }
        MK_CHECK_LAUNCH();
    }
};

// ---- constraint preconditioner -------------------------------------------------------------------------------------------------
// The projected preconditioner P M^{-1} P differs from the inverse of the constrained operator by a term of rank 2k, and PCG pays
// ~10-20 iterations per constraint for it (more, the better M^{-1} resolves the low modes: profiles/igm_per_solve_iterations_r01_*).
// The constraint preconditioner removes that term exactly:
//     z = M^{-1} r - W (Q W)^{-1} Q M^{-1} r,      W = M^{-1} Q^T  (n x k, one V-cycle per added constraint row),
// so that Q z = 0 and the preconditioned operator on the feasible subspace is as well conditioned as M^{-1} A is on the whole space.
// CPU mirror (tests/tools/mirror_rounding.py, 65k vertices, Poisson + 16 re-solves): 985 -> 693 iterations, with typed aggregation 1179 -> 580.
constexpr int kCprecMax = 64;  // constraints beyond this many are left to the projector alone (each column costs 4 n bytes per application)

// y = (Q W)^{-1} Q z_S: one CTA (Q is k x m over the support S, Sinv is k x k with leading dimension ld)
static __global__ void __launch_bounds__(kThreads) k_cprec_coef(int k, int m, const int* __restrict__ S, const double* __restrict__ Q,
                                                        const double* __restrict__ Sinv, int ld, const double* __restrict__ z,
                                                        double* __restrict__ y, const PcgState* st) {
    extern __shared__ double sm[];
    if (st && st->done) return;
    double* vS = sm;     // m
    double* t = sm + m;  // k
    for (int s = threadIdx.x; s < m; s += blockDim.x) vS[s] = z[S[s]];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = wid; j < k; j += nw) {
        double a = 0.0;
        for (int s = lane; s < m; s += 32) a += Q[(size_t)j * m + s] * vS[s];
        a = warp_sum(a);
        if (lane == 0) t[j] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j < k; j++) a += Sinv[(size_t)i * ld + j] * t[j];
        y[i] = a;
    }
}

// z -= sum_j y_j W_j (W column-major n x k, stored in fp32: it is M^{-1} Q^T with the fp32 cycle's accuracy anyway, and (Q W)^{-1} is formed
// from the STORED values, so Q z = 0 holds for what is subtracted; half the bytes of the one streaming kernel whose cost grows with k)
static __global__ void __launch_bounds__(kThreads) k_cprec_axpy(int n, int k, const float* __restrict__ W, const double* __restrict__ y,
                                                        double* __restrict__ z, const PcgState* st) {
    __shared__ double c[kCprecMax];
    if (st && st->done) return;
    if ((int)threadIdx.x < k) c[threadIdx.x] = y[threadIdx.x];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a = z[i];
#pragma unroll 4
        for (int j = 0; j < k; j++) a -= c[j] * (double)__ldg(W + (size_t)j * n + i);
        z[i] = a;
    }
}

// The same coefficients from the fp32 cycle output (z = zscale * z_lo is not formed yet): y = (Q W)^{-1} Q (zscale z_lo)_S
static __global__ void __launch_bounds__(kThreads) k_cprec_coef_lo(int k, int m, const int* __restrict__ S, const double* __restrict__ Q,
                                                           const double* __restrict__ Sinv, int ld, const float* __restrict__ zl,
                                                           double* __restrict__ y, const PcgState* st) {
    extern __shared__ double sm[];
    if (st->done) return;
    double* vS = sm;     // m
    double* t = sm + m;  // k
    const double zs = st->zscale;
    for (int s = threadIdx.x; s < m; s += blockDim.x) vS[s] = zs * (double)zl[S[s]];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = wid; j < k; j += nw) {
        double a = 0.0;
        for (int s = lane; s < m; s += 32) a += Q[(size_t)j * m + s] * vS[s];
        a = warp_sum(a);
        if (lane == 0) t[j] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double a = 0.0;
        for (int j = 0; j < k; j++) a += Sinv[(size_t)i * ld + j] * t[j];
        y[i] = a;
    }
}

// The end of the preconditioner fused with the PCG reduction: z = zscale z_lo - W y (fp64, written once), rho_new = <r, z>, beta = rho_new / rho.
// Replaces the hand-back kernel, the streaming correction and k_dot_rz (6 + 32 + 10 us at 1M vertices) by one pass.
static __global__ void __launch_bounds__(kThreads) k_cprec_finish(int n, int k, const float* __restrict__ W, const double* __restrict__ y,
                                                          const float* __restrict__ zl, const double* __restrict__ r, double* __restrict__ z,
                                                          PcgState* st, ReduceWs ws) {
    __shared__ double c[kCprecMax];
    __shared__ double red[32];
    if (st->done) return;
    if ((int)threadIdx.x < k) c[threadIdx.x] = y[threadIdx.x];
    __syncthreads();
    const double zs = st->zscale;
    double acc = 0.0;
    // two rows per thread: 8-byte loads from every column of W (n is even or the tail row is handled alone; the columns start n * 4 bytes apart,
    // so pairs stay 8-byte aligned exactly when n is even), 8 columns in flight
    const int n2 = (n & 1) ? 0 : n / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const float2 zz = __ldg(reinterpret_cast<const float2*>(zl) + i);
        double a0 = zs * (double)zz.x, a1 = zs * (double)zz.y;
        int j = 0;
        for (; j + 8 <= k; j += 8) {
            float2 w[8];
#pragma unroll
            for (int u = 0; u < 8; u++) w[u] = __ldg(reinterpret_cast<const float2*>(W + (size_t)(j + u) * n) + i);
#pragma unroll
            for (int u = 0; u < 8; u++) { a0 -= c[j + u] * (double)w[u].x; a1 -= c[j + u] * (double)w[u].y; }
        }
        for (; j < k; j++) {
            const float2 w = __ldg(reinterpret_cast<const float2*>(W + (size_t)j * n) + i);
            a0 -= c[j] * (double)w.x;
            a1 -= c[j] * (double)w.y;
        }
        const double2 rr = __ldg(reinterpret_cast<const double2*>(r) + i);
        reinterpret_cast<double2*>(z)[i] = make_double2(a0, a1);
        acc += rr.x * a0 + rr.y * a1;
    }
    for (int i = 2 * n2 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a = zs * (double)__ldg(zl + i);
        for (int j = 0; j < k; j++) a -= c[j] * (double)__ldg(W + (size_t)j * n + i);
        z[i] = a;
        acc += ld_ro(r + i) * a;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
        st->beta = s[0] / st->rho;
        st->rho = s[0];
    });
}

static __global__ void __launch_bounds__(kThreads) k_narrow_column(int n, const double* __restrict__ v, float* __restrict__ w) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) w[i] = (float)v[i];
}
static __global__ void k_gather_f32(int m, const int* __restrict__ S, const float* __restrict__ v, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = (double)v[S[i]];
}

// v[S[i]] = q[i] (v zeroed before)
static __global__ void k_scatter(int m, const int* __restrict__ S, const double* __restrict__ q, double* __restrict__ v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) v[S[i]] = q[i];
}

struct ConstraintPrecond : Precond<double> {
    Precond<double>* inner = nullptr;
    ConstraintProjector* proj = nullptr;
    int n = 0, k = 0;
    DevBuf<float> W;
    DevBuf<double> Sinv, y, qfull, wcol, wS;
    DevBuf<PcgState> unit;                // rr = 1, done = 0: solver state for V-cycle applications outside a solve
    std::vector<std::vector<double>> WS;  // host copies of W_j restricted to the support
    long setup_launches = 0;

    void init(Precond<double>* inner_, ConstraintProjector* proj_, int n_, cudaStream_t s) {
        inner = inner_; proj = proj_; n = n_; k = 0;
        W.alloc((size_t)n * kCprecMax);
        Sinv.alloc((size_t)kCprecMax * kCprecMax);
        y.alloc(kCprecMax);
        qfull.alloc(n);
        wcol.alloc(n);
        wS.alloc(std::max(proj->m, 1));
        unit.alloc(1);
        PcgState h;
        std::memset(&h, 0, sizeof(h));
        h.rr = 1.0; h.rho = 1.0; h.bnorm2 = 1.0;
        unit.upload(&h, 1, s);
        if ((size_t)(proj->m + kCprecMax) * sizeof(double) > 200 * 1024) throw std::invalid_argument("constraint preconditioner: support too large (disable with MK_IGM_CPREC=0)");
        if ((size_t)(proj->m + kCprecMax) * sizeof(double) > 48 * 1024) {
            MK_CUDA(cudaFuncSetAttribute(k_cprec_coef, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((proj->m + kCprecMax) * sizeof(double))));
            MK_CUDA(cudaFuncSetAttribute(k_cprec_coef_lo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((proj->m + kCprecMax) * sizeof(double))));
        }
        MK_CUDA(cudaStreamSynchronize(s));
    }
    void reset() { k = 0; WS.clear(); }
    // column j = M^{-1} q_j for the constraint row just appended to `Qh` (rows 0..j are on the device in proj->Q); rebuilds (Q W)^{-1}
    void add_column(const std::vector<std::vector<double>>& Qh, cudaStream_t s) {
        const int j = (int)Qh.size() - 1, m = proj->m;
        if (j != k || j >= kCprecMax) return;  // beyond the capacity: the projector alone handles the row
        qfull.zero(s);
        k_scatter<<<(m + 127) / 128, 128, 0, s>>>(m, proj->S.p, proj->Q.p + (size_t)j * m, qfull.p);
        inner->apply(qfull.p, wcol.p, unit.p, s);
        k_narrow_column<<<stream_grid(n), kThreads, 0, s>>>(n, wcol.p, W.p + (size_t)j * n);
        k_gather_f32<<<(m + 127) / 128, 128, 0, s>>>(m, proj->S.p, W.p + (size_t)j * n, wS.p);
        MK_CHECK_LAUNCH();
        WS.emplace_back(m);
        wS.download(WS.back().data(), m, s);
        MK_CUDA(cudaStreamSynchronize(s));
        setup_launches += inner->launches_per_apply() + 3;
        const int kk = j + 1;
        std::vector<double> A((size_t)kk * kk), I((size_t)kk * kk, 0.0);
        for (int a = 0; a < kk; a++)
            for (int b = 0; b < kk; b++) {
                double d = 0;
                for (int t = 0; t < m; t++) d += Qh[a][t] * WS[b][t];
                A[(size_t)a * kk + b] = d;
            }
        for (int a = 0; a < kk; a++) I[(size_t)a * kk + a] = 1.0;
        // Gauss-Jordan with partial pivoting (kk <= 32)
        for (int c = 0; c < kk; c++) {
            int piv = c;
            for (int r = c + 1; r < kk; r++)
                if (std::fabs(A[(size_t)r * kk + c]) > std::fabs(A[(size_t)piv * kk + c])) piv = r;
            if (!(std::fabs(A[(size_t)piv * kk + c]) > 0.0)) return;  // singular: keep the previous k (projector still exact)
            if (piv != c)
                for (int b = 0; b < kk; b++) { std::swap(A[(size_t)piv * kk + b], A[(size_t)c * kk + b]); std::swap(I[(size_t)piv * kk + b], I[(size_t)c * kk + b]); }
            const double inv = 1.0 / A[(size_t)c * kk + c];
            for (int b = 0; b < kk; b++) { A[(size_t)c * kk + b] *= inv; I[(size_t)c * kk + b] *= inv; }
            for (int r = 0; r < kk; r++) {
                if (r == c) continue;
                const double f = A[(size_t)r * kk + c];
                if (f == 0.0) continue;
                for (int b = 0; b < kk; b++) { A[(size_t)r * kk + b] -= f * A[(size_t)c * kk + b]; I[(size_t)r * kk + b] -= f * I[(size_t)c * kk + b]; }
            }
        }
        std::vector<double> padded((size_t)kCprecMax * kCprecMax, 0.0);
        for (int a = 0; a < kk; a++)
            for (int b = 0; b < kk; b++) padded[(size_t)a * kCprecMax + b] = I[(size_t)a * kk + b];
        Sinv.upload(padded.data(), padded.size(), s);
        MK_CUDA(cudaStreamSynchronize(s));
        k = kk;
    }
    void apply(const double* r, double* z, const PcgState* st, cudaStream_t s) override {
        inner->apply(r, z, st, s);
        if (k == 0) return;
        k_cprec_coef<<<1, kThreads, (size_t)(proj->m + k) * sizeof(double), s>>>(k, proj->m, proj->S.p, proj->Q.p, Sinv.p, kCprecMax, z, y.p, st);
        k_cprec_axpy<<<stream_grid(n), kThreads, 0, s>>>(n, k, W.p, y.p, z, st);
        MK_CHECK_LAUNCH();
    }
    int launches_per_apply() const override { return inner->launches_per_apply() + (k ? 2 : 0); }
    // fused PCG body with a projector (solver.cuh): the inner mixed-precision cycle hands its fp32 output over to k_cprec_finish
    bool fused_projected() const override { return inner->fused_pcg() && inner->lo_result() != nullptr; }
    void update_xr_handover(int n_, double* x, double* r, const double* p, const double* q, PcgState* st, ReduceWs ws, cudaStream_t s) override {
        inner->update_xr_handover(n_, x, r, p, q, st, ws, s);
    }
    void cycle_to_z(const double* r, double* z, PcgState* st, ReduceWs ws, cudaStream_t s) override {
        inner->cycle_only(st, s);
        const float* zl = static_cast<const float*>(inner->lo_result());
        if (k > 0)
            k_cprec_coef_lo<<<1, kThreads, (size_t)(proj->m + k) * sizeof(double), s>>>(k, proj->m, proj->S.p, proj->Q.p, Sinv.p, kCprecMax, zl, y.p, st);
        k_cprec_finish<<<stream_grid(n), kThreads, 0, s>>>(n, k, W.p, y.p, zl, r, z, st, ws);
        MK_CHECK_LAUNCH();
    }
    int launches_fused_projected() const override { return inner->launches_fused() - 1 + (k ? 2 : 1); }
};

// On by default (1M vertices, B200: IGM solve 594 -> 412 ms, 1347 -> 841 inner iterations, same integers, fullx within 1.5e-9;
// profiles/cprec_probe_r01.txt); MK_IGM_CPREC=0 restores the projected preconditioner P M^{-1} P.
inline bool constraint_precond_enabled() {
    const char* e = getenv("MK_IGM_CPREC");
    return !(e && std::strcmp(e, "0") == 0);
}

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
