// The localInjectivity path of iterative_rounding on the device (example_qgp.cpp:26-29 and example_qex.cpp:59-63 run it):
// the per-face injectivity barrier (igm/iterative_rounding/injective_barrier.h:19-109), its Jacobian block in the two LMs
// (iter_rounding_common.h:70-85) and the per-face algebra that lets the NI-stage LM over [x; f] (iter_rounding_init.h:60-72, 211-223) be
// solved in the reduced space.
//
// Barrier (barrier.h:54-109), per face i and field pair j: r = (v_j x v_{j+1}) / V_ij with V_ij = |v_j| |v_{j+1}| of the INPUT field,
// bar(r) = (r/s)^3 - 3 (r/s)^2 + 3 r/s, s = 0.5; residual fB = 1 / bar - 1 (+inf for r <= 0, 0 for r >= s); derivative
// barDer = -bar'(r) / bar(r)^2 (0 where |fB| < 10e-9); d r / d (v_j, v_{j+1}) = (v1.y, -v1.x, -v0.y, v0.x) / V. One thread per face.
//
// NI stage: the LM unknown is [x; f] (reduced uv dofs; 2N field doubles per face). Its damped normal matrix has the per-face block
// K_f = (1e8 + 1)(1 + lambda) I + 1e8 (G^T G + lambda diag(G^T G)) in the f rows (G = barrier Jacobian w.r.t. f: 4 entries per row), so f is
// eliminated per face (an 8 x 8 inverse for the few faces whose barrier is active, a scalar otherwise) and the direction comes from the
// Schur complement  S = X^T W X + lambda D_x,  W_f = 1e8 (K_f - 1e8 I) K_f^{-1}  (formed without cancellation), X = length * X2F -- two
// SpMVs with X2F and one per-face kernel per application, preconditioned by the multigrid hierarchy of A0 = X^T X that exists anyway.
// tests/tools/barrier_mirror.py is the numpy mirror of this algebra; it reproduces the oracle's NI iterate to 1e-13.
#pragma once
#include "csr.cuh"
#include "devcsr.cuh"
#include "solver.cuh"

namespace mk {

constexpr int kBarMaxN = 4;  // the barrier path supports N = 2 and N = 4 (2N <= 8 field doubles per face)

static __global__ void __launch_bounds__(kThreads) k_bar_V(int nF, int N, const double* __restrict__ F2, double* __restrict__ V) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        const double* f = F2 + (size_t)2 * N * i;
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            V[(size_t)N * i + j] = sqrt(f[2 * j] * f[2 * j] + f[2 * j + 1] * f[2 * j + 1]) * sqrt(f[2 * j1] * f[2 * j1] + f[2 * j1 + 1] * f[2 * j1 + 1]);
        }
    }
}

// fB[N nF], gco[4 N nF] = barDer * S (the non-zeros of row (i, j) of G at columns 2j, 2j+1, 2(j+1), 2(j+1)+1 of face i), active[nF] = 1 if any
// barDer of the face is non-zero; aux0 = sum fB^2 (+inf propagates), aux1 = number of active faces. gco / active may be null (energy only).
static __global__ void __launch_bounds__(kThreads) k_bar_eval(int nF, int N, const double* __restrict__ fld, const double* __restrict__ V,
                                                       double* __restrict__ fB, double* __restrict__ gco, int* __restrict__ active, PcgState* st,
                                                       ReduceWs ws) {
    __shared__ double red[32];
    const double s = 0.5;
    double e = 0.0, na = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        const double* f = fld + (size_t)2 * N * i;
        int act = 0;
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            const double v0x = f[2 * j], v0y = f[2 * j + 1], v1x = f[2 * j1], v1y = f[2 * j1 + 1];
            const double Vij = V[(size_t)N * i + j];
            const double r = (v0x * v1y - v0y * v1x) / Vij;
            double b2, der;
            if (r <= 0.0) { b2 = INFINITY; der = INFINITY; }
            else if (r >= s) { b2 = 0.0; der = 0.0; }
            else {
                const double q = r / s;
                const double bar = q * q * q - 3.0 * q * q + 3.0 * q;
                b2 = 1.0 / bar - 1.0;
                const double d = 3.0 * (r * r / (s * s * s)) - 6.0 * (r / (s * s)) + 3.0 / s;
                der = -d / (bar * bar);
                if (fabs(b2) < 10e-9) der = 0.0;
            }
            if (fB) fB[(size_t)N * i + j] = b2;
            e += b2 * b2;
            if (gco) {
                double* g = gco + ((size_t)N * i + j) * 4;
                if (der == 0.0) { g[0] = g[1] = g[2] = g[3] = 0.0; }
                else { g[0] = der * v1y / Vij; g[1] = -der * v1x / Vij; g[2] = -der * v0y / Vij; g[3] = der * v0x / Vij; act = 1; }
            }
        }
        if (active) active[i] = act;
        na += act;
    }
    double v[2] = {e, na};
    grid_reduce<2>(v, ws.partials, ws.ticket, red, [&](const double* t) { st->aux0 = t[0]; st->aux1 = t[1]; });
}

// out (+)= scale * G^T w  (w: N per face, out: 2N per face)
static __global__ void __launch_bounds__(kThreads) k_bar_Gt(int nF, int N, const double* __restrict__ gco, const double* __restrict__ w, double scale,
                                                     double* __restrict__ out, int accumulate) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        double o[2 * kBarMaxN];
        for (int c = 0; c < 2 * N; c++) o[c] = 0.0;
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            const double* g = gco + ((size_t)N * i + j) * 4;
            double wj = w[(size_t)N * i + j];
            if (g[0] == 0.0 && g[1] == 0.0 && g[2] == 0.0 && g[3] == 0.0) continue;  // inactive row: its fB may be finite or 0, never counted
            o[2 * j] += g[0] * wj; o[2 * j + 1] += g[1] * wj; o[2 * j1] += g[2] * wj; o[2 * j1 + 1] += g[3] * wj;
        }
        for (int c = 0; c < 2 * N; c++) {
            double* p = out + (size_t)2 * N * i + c;
            *p = (accumulate ? *p : 0.0) + scale * o[c];
        }
    }
}

// t <- scale * G^T G t per face (in place, 2N per face); inactive faces give 0
static __global__ void __launch_bounds__(kThreads) k_bar_GtG(int nF, int N, const double* __restrict__ gco, const int* __restrict__ active, double scale,
                                                      double* __restrict__ t) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        double* f = t + (size_t)2 * N * i;
        if (!active[i]) { for (int c = 0; c < 2 * N; c++) f[c] = 0.0; continue; }
        double gw[kBarMaxN], o[2 * kBarMaxN];
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            const double* g = gco + ((size_t)N * i + j) * 4;
            gw[j] = g[0] * f[2 * j] + g[1] * f[2 * j + 1] + g[2] * f[2 * j1] + g[3] * f[2 * j1 + 1];
        }
        for (int c = 0; c < 2 * N; c++) o[c] = 0.0;
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            const double* g = gco + ((size_t)N * i + j) * 4;
            o[2 * j] += g[0] * gw[j]; o[2 * j + 1] += g[1] * gw[j]; o[2 * j1] += g[2] * gw[j]; o[2 * j1 + 1] += g[3] * gw[j];
        }
        for (int c = 0; c < 2 * N; c++) f[c] = scale * o[c];
    }
}

// diag += scale * diag(Jb^T Jb), Jb = G X: row (i, j) of Jb = sum_a gco[a] * X[row_a, :] (a 4-way merge of sorted rows), squared entries
// added per column. Only active rows do any work (a handful); double atomics.
static __global__ void __launch_bounds__(kThreads) k_bar_diag(int nF, int N, const double* __restrict__ gco, const int* __restrict__ Xp,
                                                       const int* __restrict__ Xc, const double* __restrict__ Xv, double xscale, double scale,
                                                       double* __restrict__ diag) {
    const int64_t rows = (int64_t)N * nF;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const double* g = gco + r * 4;
        if (g[0] == 0.0 && g[1] == 0.0 && g[2] == 0.0 && g[3] == 0.0) continue;
        const int i = (int)(r / N), j = (int)(r - (int64_t)i * N), j1 = (j + 1) % N;
        const int xr[4] = {2 * N * i + 2 * j, 2 * N * i + 2 * j + 1, 2 * N * i + 2 * j1, 2 * N * i + 2 * j1 + 1};
        int b[4], e[4];
        for (int a = 0; a < 4; a++) { b[a] = Xp[xr[a]]; e[a] = Xp[xr[a] + 1]; }
        while (true) {
            int col = 0x7fffffff;
            for (int a = 0; a < 4; a++)
                if (b[a] < e[a]) col = min(col, Xc[b[a]]);
            if (col == 0x7fffffff) break;
            double v = 0.0;
            for (int a = 0; a < 4; a++)
                if (b[a] < e[a] && Xc[b[a]] == col) { v += g[a] * xscale * Xv[b[a]]; b[a]++; }
            atomicAdd(diag + col, scale * v * v);
        }
    }
}

// ---- NI stage: per-face blocks of the eliminated field unknowns ----
// mode 0: t <- K_f^{-1} t ; mode 1: t <- W_f t, W = 1e8 (K - 1e8 I) K^{-1}.  K = k0 I + 1e8 (M + lambda diag M), M = G^T G. Inactive faces:
// scalars. Active faces: the 2N x 2N system is solved in registers (Cholesky, K is SPD) -- no per-face matrices are stored.
static __global__ void __launch_bounds__(kThreads) k_ni_face_solve(int nF, int N, const double* __restrict__ gco, const int* __restrict__ active, double lambda,
                                                            int mode, double* __restrict__ t) {
    const double k0 = (1e8 + 1.0) * (1.0 + lambda), r0 = 1.0 + lambda * (1e8 + 1.0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nF; i += gridDim.x * blockDim.x) {
        double* f = t + (size_t)2 * N * i;
        const int m = 2 * N;
        if (!active[i]) {
            const double sc = mode == 0 ? 1.0 / k0 : 1e8 * r0 / k0;
            for (int c = 0; c < m; c++) f[c] *= sc;
            continue;
        }
        double M[2 * kBarMaxN][2 * kBarMaxN];
        for (int a = 0; a < m; a++)
            for (int c = 0; c < m; c++) M[a][c] = 0.0;
        for (int j = 0; j < N; j++) {
            const int j1 = (j + 1) % N;
            const double* g = gco + ((size_t)N * i + j) * 4;
            const int cc[4] = {2 * j, 2 * j + 1, 2 * j1, 2 * j1 + 1};
            for (int a = 0; a < 4; a++)
                for (int c = 0; c < 4; c++) M[cc[a]][cc[c]] += g[a] * g[c];
        }
        double A[2 * kBarMaxN][2 * kBarMaxN], y[2 * kBarMaxN], rhs[2 * kBarMaxN];
        for (int a = 0; a < m; a++) {
            for (int c = 0; c < m; c++) A[a][c] = 1e8 * M[a][c];
            A[a][a] = 1e8 * (1.0 + lambda) * M[a][a];   // M + lambda diag(M) on the diagonal
            rhs[a] = f[a];
        }
        // Cholesky of K = k0 I + A
        double L[2 * kBarMaxN][2 * kBarMaxN];
        for (int a = 0; a < m; a++) {
            for (int c = 0; c <= a; c++) {
                double sum = A[a][c] + (a == c ? k0 : 0.0);
                for (int k = 0; k < c; k++) sum -= L[a][k] * L[c][k];
                L[a][c] = (a == c) ? sqrt(sum) : sum / L[c][c];
            }
        }
        for (int a = 0; a < m; a++) {
            double sum = rhs[a];
            for (int k = 0; k < a; k++) sum -= L[a][k] * y[k];
            y[a] = sum / L[a][a];
        }
        for (int a = m - 1; a >= 0; a--) {
            double sum = y[a];
            for (int k = a + 1; k < m; k++) sum -= L[k][a] * y[k];
            y[a] = sum / L[a][a];
        }
        if (mode == 0) {
            for (int c = 0; c < m; c++) f[c] = y[c];
        } else {
            // W t = 1e8 (r0 I + A) K^{-1} t
            for (int a = 0; a < m; a++) {
                double sum = r0 * y[a];
                for (int c = 0; c < m; c++) sum += A[a][c] * y[c];
                f[a] = 1e8 * sum;
            }
        }
    }
}

// ---- small vector kernels of the LM driver ----
static __global__ void __launch_bounds__(kThreads) k_vec_lincomb(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                                                          double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}
// aux0 = <a, b>
static __global__ void __launch_bounds__(kThreads) k_vec_dot(int64_t n, const double* __restrict__ a, const double* __restrict__ b, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += a[i] * b[i];
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* t) { st->aux0 = t[0]; });
}
// y += lambda * d .* x
static __global__ void __launch_bounds__(kThreads) k_vec_add_diag(int n, double lambda, const double* __restrict__ d, const double* __restrict__ x,
                                                           double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += lambda * d[i] * x[i];
}
// d += c
static __global__ void __launch_bounds__(kThreads) k_vec_add_const(int n, double c, double* __restrict__ d) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] += c;
}
// z = r ./ d   (Jacobi preconditioner of the heavily damped systems)
static __global__ void __launch_bounds__(kThreads) k_vec_div(int n, const double* __restrict__ r, const double* __restrict__ d, double* __restrict__ z,
                                                      const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] = r[i] / d[i];
}
// diag[i] = A(i, i) of a CSR matrix
static __global__ void __launch_bounds__(kThreads) k_csr_diag(int n, const int* __restrict__ ptr, const int* __restrict__ col, const double* __restrict__ val,
                                                       double* __restrict__ diag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double d = 0.0;
        for (int k = ptr[i]; k < ptr[i + 1]; k++)
            if (col[k] == i) d += val[k];
        diag[i] = d;
    }
}
// locally_injectivity_check (integ.ipp:46-63): aux0 = number of faces whose uv triangle (first two functions of nfn at the three corners)
// has negative orientation
static __global__ void __launch_bounds__(kThreads) k_orientation_count(int nF, int N, const int* __restrict__ cutF, const double* __restrict__ nfn,
                                                                PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    double bad = 0.0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        const double* a = nfn + (size_t)N * cutF[3 * f];
        const double* b = nfn + (size_t)N * cutF[3 * f + 1];
        const double* c = nfn + (size_t)N * cutF[3 * f + 2];
        const double d1x = b[0] - a[0], d1y = b[1] - a[1], d2x = c[0] - a[0], d2y = c[1] - a[1];
        if (d1x * d2y - d1y * d2x < 0.0) bad += 1.0;
    }
    double v[1] = {bad};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* t) { st->aux0 = t[0]; });
}

// Scaled / Jacobi preconditioners of the LM systems
struct ScaledPrecond : Precond<double> {
    Precond<double>* inner = nullptr;
    double scale = 1.0;  // z = scale * inner(r)
    int n = 0;
    void apply(const double* r, double* z, const PcgState* st, cudaStream_t s) override {
        inner->apply(r, z, st, s);
        if (scale != 1.0) { k_vec_lincomb<<<stream_grid(n), kThreads, 0, s>>>(n, scale, z, 0.0, nullptr, z); MK_CHECK_LAUNCH(); }
    }
    int launches_per_apply() const override { return inner->launches_per_apply() + 1; }
};
struct JacobiPrecond : Precond<double> {
    const double* d = nullptr;
    int n = 0;
    void apply(const double* r, double* z, const PcgState* st, cudaStream_t s) override {
        k_vec_div<<<stream_grid(n), kThreads, 0, s>>>(n, r, d, z, st);
        MK_CHECK_LAUNCH();
    }
    int launches_per_apply() const override { return 1; }
};

}  // namespace mk
