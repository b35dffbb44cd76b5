// Smoothed-aggregation algebraic multigrid used as the PCG preconditioner. It stands where the reference has a sparse
// direct factorisation (Eigen::SparseLU sls.h:42-43 / CHOLMOD LDL^T ss/suitesparse_pdefinite.h:26-41 /
// Eigen::SimplicialLLT ir.h:14-20): plain CG needs O(sqrt(kappa)) ~ thousands of iterations on these operators
// (kappa grows with the face count), the V-cycle keeps the count nearly mesh-independent while every
// piece of it stays an HBM-bound SpMV / fused vector kernel.
//
// Hierarchy set-up (aggregation, prolongator smoothing, Galerkin products) runs once per operator on the host with
// OpenMP; the V-cycle runs on the device (Chebyshev smoothing, explicit R = P^H, dense inverse on the coarsest level).
//
// Tentative prolongator:
//   * complex connection Laplacian: there is no global near-kernel vector, but locally parallel fields are low energy,
//     so each aggregate transports a unit value from its root along the matrix' own connection phases
//     (value_j = -a_ji/|a_ji| value_i), recursively on every level;
//   * real IGM operator: a caller-supplied near-kernel candidate (smoothed constants), standard SA.
#pragma once
#include <complex>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "csr.cuh"
#include "hostcsr.hpp"
#include "solver.cuh"
#include "vtail.cuh"

namespace mk {

template <class T> struct HostScalar;
template <> struct HostScalar<double> { using type = double; };
template <> struct HostScalar<cdouble> { using type = std::complex<double>; };
template <> struct HostScalar<float> { using type = float; };
template <> struct HostScalar<cfloat> { using type = std::complex<float>; };

struct AmgParams {
    double theta = 0.0;        // strength threshold (0: every off-diagonal is strong)
    int coarse_max = 300;      // stop coarsening below this many rows (dense inverse there)
    int max_levels = 14;
    int cheb_degree = 1;       // SpMVs per pre- and per post-smoothing (measured at 1M vertices: degree 1 = 2330 ms of solves, 2 = 2593, 3 = 2978)
    double cheb_ratio = 8.0;   // smoother targets [lmax / ratio, lmax] of D^{-1} A
    double omega_scale = 4.0 / 3.0;  // prolongator smoothing omega = omega_scale / rho(D^{-1}A)
    int cheb_degree_coarse = 2;  // degree on the levels below the finest (latency-bound: launches, not bytes); 0 = cheb_degree. Measured at
                                 // 1M vertices (profiles/amg_tuning_r01g_typed_degcoarse.txt): field solve 533 / 506 / 531 ms at 1 / 2 / 3
                                 // (1186 / 1017 / 988 iterations), IGM solve 619 / 595 ms at 1 / 2 (1537 / 1350 iterations)
    int dense_row = 64;        // fine-level rows longer than this are "dense dofs": never aggregated with others, carried
                               // as singletons to the coarsest level (solved exactly there), l1-scaled in the smoother
    int agg = 1;               // 0 = sequential greedy (Vanek), 1 = MIS(2) roots (parallel; what the device set-up runs)
    int inject_dense = 1;      // prolongator rows of dense dofs are not smoothed (injection from their singleton aggregate)
    int weighted = 1;          // transport mode: a coarse dof stands for an aggregate whose tentative column was normalised, so a field of
                               // constant modulus has coarse coefficients proportional to the column norms (sqrt of the aggregate size on
                               // level 1). The transported unit value is therefore scaled by that norm on every level below the finest --
                               // what smoothed aggregation's local QR does with a one-column candidate. 0 = unit modulus on every level
                               // (rounds 1-2a: the V-cycle lost against the two-grid method level by level: 66 against 19 iterations
                               // at 1M vertices in the CPU mirror, tests/tools/field_cycle_study.py)
};

// real_op: the real symmetric IGM / LM operators use a strength threshold (drops weak and wrong-signed cotan couplings
// from the aggregation graph); the complex face operator has unit-modulus off-diagonals and keeps theta = 0.
inline AmgParams amg_params_from_env(bool real_op = false) {
    AmgParams p;
    // theta of the complex operator only acts below the finest level (its own off-diagonals all have unit modulus): dropping the weakest
    // Galerkin couplings from the aggregation graph keeps the coarse aggregates compact (level 1 -> 2 coarsens 8:1 instead of 17:1).
    // 1M vertices, B200: 480 / 359 / 336 / 330 / 343 / 838 inner iterations per 50-step solve at theta = 0 / 0.02 / 0.03 / 0.05 / 0.08 / 0.1
    // (profiles/amg_tuning_r02_*): 0.03 sits in the flat part, far from the cliff at 0.1; torus and 16k ... 4M-vertex spheres agree.
    p.theta = real_op ? 0.1 : 0.03;
    if (const char* e = getenv("MK_AMG_DEG")) p.cheb_degree = atoi(e);
    if (const char* e = getenv("MK_AMG_DEG_COARSE")) p.cheb_degree_coarse = atoi(e);
    if (const char* e = getenv("MK_AMG_RATIO")) p.cheb_ratio = atof(e);
    if (const char* e = getenv("MK_AMG_THETA")) p.theta = atof(e);
    if (const char* e = getenv("MK_AMG_COARSE")) p.coarse_max = atoi(e);
    if (const char* e = getenv("MK_AMG_OMEGA")) p.omega_scale = atof(e);
    if (const char* e = getenv("MK_AMG_DENSE")) p.dense_row = atoi(e);
    if (const char* e = getenv("MK_AMG_AGG")) p.agg = atoi(e);
    if (const char* e = getenv("MK_AMG_INJECT")) p.inject_dense = atoi(e);
    if (const char* e = getenv("MK_AMG_WEIGHTED")) p.weighted = atoi(e);
    // per-operator overrides (the complex face operator and the real IGM / LM operators want different values)
    if (const char* e = getenv(real_op ? "MK_AMG_THETA_REAL" : "MK_AMG_THETA_FIELD")) p.theta = atof(e);
    if (const char* e = getenv(real_op ? "MK_AMG_COARSE_REAL" : "MK_AMG_COARSE_FIELD")) p.coarse_max = atoi(e);
    return p;
}

// MK_SETUP=host: build hierarchies (and the IGM normal matrix) with the host mirror instead of the device set-up
inline bool setup_on_host() {
    const char* e = getenv("MK_SETUP");
    return e && std::strcmp(e, "host") == 0;
}

template <class S>
struct AmgHostLevel {
    HostCsr<S> A, P, R;
    std::vector<S> dinv;
    double lmax = 2.0;
};

// ---- host set-up ---------------------------------------------------------------------------------

template <class S>
inline double estimate_rho_dinvA(const HostCsr<S>& A, const std::vector<S>& dinv, int iters = 12) {
    const int n = A.rows;
    std::vector<S> v(n), w;
    for (int i = 0; i < n; i++) v[i] = S(std::sin(1.0 + 0.7 * i) + 0.3);
    double rho = 1.0;
    for (int it = 0; it < iters; it++) {
        matvec(A, v, w);
        double nw = 0, nv = 0;
#pragma omp parallel for reduction(+ : nw, nv)
        for (int i = 0; i < n; i++) {
            w[i] = dinv[i] * w[i];
            nw += std::norm(w[i]);
            nv += std::norm(v[i]);
        }
        rho = std::sqrt(nw / nv);
        const double s = 1.0 / std::sqrt(nw);
#pragma omp parallel for
        for (int i = 0; i < n; i++) v[i] = w[i] * s;
    }
    return rho;
}

template <class S>
inline S unit_phase(S v) {
    double a = std::abs(v);
    return a > 0 ? v / a : S(1);
}

// Greedy root-plus-neighbours aggregation (Vanek). Returns the aggregate id per node and the tentative value per node.
// `dense` (may be empty): rows that must stay alone. They are skipped by all passes, ignored as neighbours, and appended
// as singleton aggregates at the end; their coarse ids are returned in `dense_coarse`.
template <class S>
inline int aggregate(const HostCsr<S>& A, double theta, const std::vector<S>* cand, const std::vector<char>& dense, std::vector<int>& agg,
                     std::vector<S>& tval, std::vector<int>& dense_coarse) {
    const int n = A.rows;
    const bool has_dense = !dense.empty();
    std::vector<double> dabs(n, 0.0);
    for (int i = 0; i < n; i++)
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (A.col[k] == i) dabs[i] = std::abs(A.val[k]);
    auto strong = [&](int i, int k) {
        const int j = A.col[k];
        if (j == i || (has_dense && dense[j])) return false;
        const double a = std::abs(A.val[k]);
        return a > 0 && a * a >= theta * theta * dabs[i] * dabs[j];
    };
    agg.assign(n, -1);
    tval.assign(n, S(0));
    int nagg = 0;
    // pass 1: roots whose whole strong neighbourhood is free
    auto skip = [&](int i) { return agg[i] != -1 || (has_dense && dense[i]); };
    for (int i = 0; i < n; i++) {
        if (skip(i)) continue;
        bool ok = true, any = false;
        for (int k = A.ptr[i]; k < A.ptr[i + 1] && ok; k++)
            if (strong(i, k)) { any = true; if (agg[A.col[k]] != -1) ok = false; }
        if (!ok || !any) continue;
        agg[i] = nagg;
        tval[i] = S(1);
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (strong(i, k)) {
                const int j = A.col[k];
                agg[j] = nagg;
                tval[j] = unit_phase(S(-1) * hconj(A.val[k]));  // -a_ji = -conj(a_ij) (Hermitian): transport root -> j
            }
        nagg++;
    }
    // pass 2: attach the rest to the strongest neighbouring pass-1 aggregate
    std::vector<int> agg1 = agg;
    for (int i = 0; i < n; i++) {
        if (skip(i)) continue;
        double best = -1;
        int bk = -1;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (strong(i, k) && agg1[A.col[k]] != -1 && std::abs(A.val[k]) > best) { best = std::abs(A.val[k]); bk = k; }
        if (bk >= 0) {
            const int j = A.col[bk];
            agg[i] = agg1[j];
            tval[i] = unit_phase(S(-1) * A.val[bk]) * tval[j];  // value_i = -a_ij/|a_ij| value_j
        }
    }
    // pass 3: leftovers (isolated or surrounded by pass-2 nodes) form their own aggregates
    for (int i = 0; i < n; i++) {
        if (skip(i)) continue;
        agg[i] = nagg;
        tval[i] = S(1);
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (strong(i, k) && agg[A.col[k]] == -1) {
                const int j = A.col[k];
                agg[j] = nagg;
                tval[j] = unit_phase(S(-1) * hconj(A.val[k]));
            }
        nagg++;
    }
    dense_coarse.clear();
    if (has_dense)
        for (int i = 0; i < n; i++)
            if (dense[i]) { agg[i] = nagg; tval[i] = S(1); dense_coarse.push_back(nagg); nagg++; }
    if (cand) tval = *cand;  // global-candidate mode: tentative values are the candidate itself
    return nagg;
}


// ---- MIS(2) aggregation (Bell, Dalton, Olson 2012): the parallel aggregation the device set-up runs; this host version
// executes the same synchronous rounds (every round reads the previous round's arrays only), so both produce the same
// aggregates bit for bit and the host build stays the CPU mirror of the device build.
__host__ __device__ inline uint32_t mis_hash(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// key = state (2 bits: 0 removed, 1 undecided, 2 root) | 30 priority bits | 32 bits of ~index (ties go to the smaller index).
// The first kMisIndexRounds rounds use priority 0, i.e. the smallest index wins: that is the lexicographically-first MIS(2) =
// the sequential greedy root choice, which tiles structured mesh numberings regularly (fewer PCG iterations) and needs only a
// handful of rounds on them; numberings with long dependency chains fall through to hashed priorities (O(log n) rounds).
constexpr int kMisIndexRounds = 24;
constexpr int kMisMaxRounds = 256;
// a coupling must be stronger by this factor to displace an earlier candidate: equal-modulus couplings (the face operator's are all 1
// up to an ulp of hypot) tie towards the smaller column on the host and on the device alike
constexpr double kStrongerBy = 1.0 + 1e-12;
__host__ __device__ inline uint64_t mis_key(int state, int i, bool hashed) {
    return ((uint64_t)state << 62) | ((uint64_t)(hashed ? (mis_hash((uint32_t)i) & 0x3fffffffU) : 0u) << 32) | (uint32_t)(~(uint32_t)i);
}
__host__ __device__ inline int mis_key_index(uint64_t k) { return (int)(~(uint32_t)k); }

template <class S>
inline int aggregate_mis2(const HostCsr<S>& A, double theta, const std::vector<char>& dense, std::vector<int>& agg, std::vector<S>& tval,
                          std::vector<int>& dense_coarse, const std::vector<int>* types = nullptr) {
    const int n = A.rows;
    const bool has_dense = !dense.empty();
    std::vector<double> dabs(n, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++)
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (A.col[k] == i) dabs[i] = std::abs(A.val[k]);
    const bool typed = types && !types->empty();
    auto strong = [&](int i, int k) {
        const int j = A.col[k];
        if (j == i || (has_dense && (dense[j] || dense[i]))) return false;
        if (typed && (*types)[i] != (*types)[j]) return false;  // dofs of different type never share an aggregate (see AmgParams / DESIGN.md)
        const double a = std::abs(A.val[k]);
        return a > 0 && a * a >= theta * theta * dabs[i] * dabs[j];
    };
    std::vector<int> state(n);
    std::vector<uint64_t> t1(n), t2(n);
    for (int i = 0; i < n; i++) state[i] = (has_dense && dense[i]) ? 0 : 1;
    for (int round = 0; round < kMisMaxRounds; round++) {
        const bool hashed = round >= kMisIndexRounds;
        long undecided = 0;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            uint64_t m = mis_key(state[i], i, hashed);
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
                if (strong(i, k)) m = std::max(m, mis_key(state[A.col[k]], A.col[k], hashed));
            t1[i] = m;
        }
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            uint64_t m = t1[i];
            for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
                if (strong(i, k)) m = std::max(m, t1[A.col[k]]);
            t2[i] = m;
        }
#pragma omp parallel for schedule(static) reduction(+ : undecided)
        for (int i = 0; i < n; i++) {
            if (state[i] != 1) continue;
            if (mis_key_index(t2[i]) == i) state[i] = 2;
            else if ((t2[i] >> 62) == 2) state[i] = 0;
            else undecided++;
        }
        if (!undecided || round == kMisMaxRounds - 1) {
            if (getenv("MK_VERBOSE")) fprintf(stderr, "[mk] mis2: n %d rounds %d undecided %ld\n", n, round + 1, undecided);
            break;
        }
    }
    // roots are numbered in index order
    agg.assign(n, -1);
    tval.assign(n, S(0));
    int nagg = 0;
    for (int i = 0; i < n; i++)
        if (state[i] == 2) { agg[i] = nagg++; tval[i] = S(1); }
    // round 1: neighbours of a root join it (strongest coupling, then smallest index)
    std::vector<int> agg1 = agg;
    std::vector<S> tv1 = tval;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        if (state[i] == 2 || (has_dense && dense[i])) continue;
        double best = -1;
        int bk = -1;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (strong(i, k) && state[A.col[k]] == 2 && std::abs(A.val[k]) > best * kStrongerBy) { best = std::abs(A.val[k]); bk = k; }
        if (bk >= 0) { agg1[i] = agg[A.col[bk]]; tv1[i] = unit_phase(S(-1) * A.val[bk]); }
    }
    // round 2: the rest (distance 2 from a root) join through their strongest assigned neighbour
    agg = agg1;
    tval = tv1;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        if (agg1[i] != -1 || (has_dense && dense[i])) continue;
        double best = -1;
        int bk = -1;
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++)
            if (strong(i, k) && agg1[A.col[k]] != -1 && std::abs(A.val[k]) > best * kStrongerBy) { best = std::abs(A.val[k]); bk = k; }
        if (bk >= 0) { agg[i] = agg1[A.col[bk]]; tval[i] = unit_phase(S(-1) * A.val[bk]) * tv1[A.col[bk]]; }
    }
    // leftovers (none when the strength graph is symmetric) and dense dofs become singletons, in index order
    dense_coarse.clear();
    for (int i = 0; i < n; i++)
        if (agg[i] == -1 && !(has_dense && dense[i])) { agg[i] = nagg++; tval[i] = S(1); }
    if (has_dense)
        for (int i = 0; i < n; i++)
            if (dense[i]) { agg[i] = nagg; tval[i] = S(1); dense_coarse.push_back(nagg); nagg++; }
    return nagg;
}

// `types` (optional, one int per row): dofs of different type are never put into the same aggregate, on any level (a coarse dof inherits
// the type of its members). The IGM operator needs it: its near-kernel is two-dimensional (u- and v-translation) and ONE transported
// value per aggregate cannot represent both where an aggregate would mix u- and v-type dofs across a seam.
template <class S>
inline std::vector<AmgHostLevel<S>> build_hierarchy(HostCsr<S> A, std::vector<S> cand, bool use_cand, const AmgParams& prm,
                                                    std::vector<int> types = {}) {
    std::vector<AmgHostLevel<S>> lv;
    StageTimer tm("build_hierarchy");
    // dense dofs (e.g. the seam-translation / constraint-kernel columns of the IGM operator: a few rows with thousands of
    // entries) are detected once on the finest level
    std::vector<char> dense;
    int n_dense = 0;
    if (prm.dense_row > 0 && A.rows > prm.coarse_max) {
        for (int i = 0; i < A.rows; i++)
            if (A.ptr[i + 1] - A.ptr[i] > prm.dense_row) {
                if (dense.empty()) dense.assign(A.rows, 0);
                dense[i] = 1;
                n_dense++;
            }
    }
    tm.lap("dense detection");
    while (true) {
        AmgHostLevel<S> L;
        L.A = std::move(A);
        const int n = L.A.rows;
        std::vector<S> d = diagonal(L.A);
        L.dinv.resize(n);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            if (n_dense && dense[i]) {
                // l1 scaling: keeps rho(D^{-1}A) (hence the Chebyshev interval and the prolongator damping) governed by the sparse rows
                double l1 = 0;
                for (int k = L.A.ptr[i]; k < L.A.ptr[i + 1]; k++) l1 += std::abs(L.A.val[k]);
                L.dinv[i] = l1 > 0 ? S(1.0 / l1) : S(1);
            } else {
                L.dinv[i] = std::abs(d[i]) > 0 ? S(1) / d[i] : S(1);
            }
        }
        tm.lap("dinv");
        L.lmax = 1.1 * estimate_rho_dinvA(L.A, L.dinv);
        tm.lap("rho estimate");
        if (n - n_dense <= prm.coarse_max || (int)lv.size() + 1 >= prm.max_levels) { lv.push_back(std::move(L)); break; }
        std::vector<int> agg, dense_coarse;
        std::vector<S> tv;
        const int nc = (prm.agg == 1 && !use_cand) ? aggregate_mis2(L.A, prm.theta, dense, agg, tv, dense_coarse, &types)
                                                   : aggregate(L.A, prm.theta, use_cand ? &cand : nullptr, dense, agg, tv, dense_coarse);
        tm.lap("aggregate");
        if (nc >= n * 0.8 || nc == 0) { lv.push_back(std::move(L)); break; }  // stagnation: stop here
        if (!use_cand && prm.weighted && !lv.empty())  // transport mode below the finest level: unit phase x the weight this dof stands for
            for (int i = 0; i < n; i++) tv[i] *= cand[i];
        // normalise the tentative columns; the column norms are the coarse candidate
        std::vector<double> nrm(nc, 0.0);
        for (int i = 0; i < n; i++) nrm[agg[i]] += std::norm(tv[i]);
        std::vector<S> ccand(nc);
        for (int a = 0; a < nc; a++) { nrm[a] = std::sqrt(nrm[a]); ccand[a] = S(nrm[a]); }
        HostCsr<S> T;
        T.rows = n;
        T.cols = nc;
        T.ptr.resize(n + 1);
        T.col.resize(n);
        T.val.resize(n);
        for (int i = 0; i < n; i++) {
            T.ptr[i] = i;
            T.col[i] = agg[i];
            T.val[i] = nrm[agg[i]] > 0 ? tv[i] / nrm[agg[i]] : S(0);
        }
        T.ptr[n] = n;
        // P = (I - omega D^{-1} A) T
        HostCsr<S> AT = matmul(L.A, T);
        tm.lap("A T");
        const double omega = prm.omega_scale / (L.lmax / 1.1);
        {
            // merge T (one entry per row) into -omega dinv AT
            HostCsr<S>& P = L.P;
            P.rows = n;
            P.cols = nc;
            P.ptr.assign(n + 1, 0);
            P.col.reserve(AT.col.size() + n);
            P.val.reserve(AT.col.size() + n);
            for (int i = 0; i < n; i++) {
                bool placed = false;
                if (n_dense && dense[i] && prm.inject_dense) {
                    // dense dofs are interpolated by injection from their own singleton aggregate: smoothing their row would spread
                    // them over every neighbouring aggregate and couple all of those pairwise in the Galerkin product (a clique per seam)
                    P.col.push_back(T.col[i]);
                    P.val.push_back(T.val[i]);
                    P.ptr[i + 1] = (int)P.col.size();
                    continue;
                }
                for (int k = AT.ptr[i]; k < AT.ptr[i + 1]; k++) {
                    int c = AT.col[k];
                    S v = S(-omega) * L.dinv[i] * AT.val[k];
                    if (!placed && T.col[i] < c) { P.col.push_back(T.col[i]); P.val.push_back(T.val[i]); placed = true; }
                    if (c == T.col[i]) { v += T.val[i]; placed = true; }
                    P.col.push_back(c);
                    P.val.push_back(v);
                }
                if (!placed) { P.col.push_back(T.col[i]); P.val.push_back(T.val[i]); }
                P.ptr[i + 1] = (int)P.col.size();
            }
        }
        tm.lap("P");
        L.R = transpose(L.P, true);
        tm.lap("R = P^H");
        HostCsr<S> AP = matmul(L.A, L.P);
        tm.lap("A P");
        A = matmul(L.R, AP);
        tm.lap("R (A P)");
        cand = std::move(ccand);
        if (!types.empty()) {
            std::vector<int> ctypes(nc, 0);
            for (int i = 0; i < n; i++) ctypes[agg[i]] = types[i];  // all members of an aggregate carry the same type
            types = std::move(ctypes);
        }
        if (n_dense) {
            dense.assign(nc, 0);
            for (int c : dense_coarse) dense[c] = 1;
        }
        lv.push_back(std::move(L));
    }
    return lv;
}

// dense inverse of the coarsest operator (Gauss-Jordan with partial pivoting), row-major n x n
template <class S>
inline std::vector<S> dense_inverse(const HostCsr<S>& A) {
    const int n = A.rows;
    std::vector<S> M((size_t)n * n, S(0)), I((size_t)n * n, S(0));
    for (int i = 0; i < n; i++) {
        for (int k = A.ptr[i]; k < A.ptr[i + 1]; k++) M[(size_t)i * n + A.col[k]] += A.val[k];
        I[(size_t)i * n + i] = S(1);
    }
    for (int c = 0; c < n; c++) {
        int p = c;
        for (int r = c + 1; r < n; r++)
            if (std::abs(M[(size_t)r * n + c]) > std::abs(M[(size_t)p * n + c])) p = r;
        if (std::abs(M[(size_t)p * n + c]) == 0) throw std::runtime_error("amg: singular coarsest operator");
        if (p != c)
            for (int j = 0; j < n; j++) { std::swap(M[(size_t)p * n + j], M[(size_t)c * n + j]); std::swap(I[(size_t)p * n + j], I[(size_t)c * n + j]); }
        const S inv = S(1) / M[(size_t)c * n + c];
        for (int j = 0; j < n; j++) { M[(size_t)c * n + j] *= inv; I[(size_t)c * n + j] *= inv; }
#pragma omp parallel for schedule(static)
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const S f = M[(size_t)r * n + c];
            if (f == S(0)) continue;
            for (int j = 0; j < n; j++) { M[(size_t)r * n + j] -= f * M[(size_t)c * n + j]; I[(size_t)r * n + j] -= f * I[(size_t)c * n + j]; }
        }
    }
    return I;
}

// ---- device V-cycle ------------------------------------------------------------------------------

// r = b (or b - q) ; d = inv_theta * dinv .* r ; x untouched (first Chebyshev step overwrites it)
template <class T>
__global__ void __launch_bounds__(kThreads) k_cheb_init(int n, const T* __restrict__ b, const T* __restrict__ q, const T* __restrict__ dinv,
                                                        T* __restrict__ r, T* __restrict__ d, double inv_theta, const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T ri = ld_ro(b + i);
        if (q) ri = ri - ld_ro(q + i);
        st_rw(r + i, ri);
        st_rw(d + i, ST<T>::scale(inv_theta, ld_ro(dinv + i) * ri));
    }
}

// d = inv_theta * dinv .* b : the start direction of a zero-guess smoothing (no residual copy: the first fused step reads b itself)
template <class T>
__global__ void __launch_bounds__(kThreads) k_start_dir(int n, const T* __restrict__ b, const T* __restrict__ dinv, T* __restrict__ d, double inv_theta,
                                                        const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        st_rw(d + i, ST<T>::scale(inv_theta, ld_ro(dinv + i) * ld_ro(b + i)));
}

// x (+)= d ; r -= q ; d = c1 d + c2 dinv .* r      (one Chebyshev step after q = A d; the partitioned operator, whose SpMV is split
// around the halo exchange, cannot fuse it)
template <class T>
__global__ void __launch_bounds__(kThreads) k_cheb_step(int n, const T* __restrict__ q, const T* __restrict__ dinv, T* __restrict__ r,
                                                        T* __restrict__ x, T* __restrict__ d, double c1, double c2, int first,
                                                        const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T di = ld_rw(d + i);
        T xi = first ? di : ld_rw(x + i) + di;
        T ri = ld_rw(r + i) - ld_ro(q + i);
        st_rw(x + i, xi);
        st_rw(r + i, ri);
        st_rw(d + i, ST<T>::scale(c1, di) + ST<T>::scale(c2, ld_ro(dinv + i) * ri));
    }
}

// x = Ainv b for the dense coarsest level: one warp per row
template <class T>
__global__ void __launch_bounds__(kThreads) k_dense_matvec(int n, const T* __restrict__ Ainv, const T* __restrict__ b, T* __restrict__ x,
                                                           const PcgState* st) {
    if (st && st->done) return;
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    T acc = ST<T>::zero();
    for (int j = lane; j < n; j += 32) acc = ST<T>::fma_(ld_ro(Ainv + (size_t)row * n + j), ld_ro(b + j), acc);
    acc = group_sum<T>(acc, 32);
    if (lane == 0) st_rw(x + row, acc);
}

// y (+)= A x inside the V-cycle (no reduction; honours the done flag for graph replay)
template <class T>
inline void spmv_plain(const DeviceCsr<T>& A, const T* x, T* y, bool accumulate, const PcgState* st, cudaStream_t s) {
    spmv<T>(A, x, y, 0.0, accumulate, false, const_cast<PcgState*>(st), ReduceWs{nullptr, nullptr}, s);
}

template <class T>
struct AmgLevelDev {
    int n = 0;
    DeviceCsr<T> A;  // unused on level 0 when an external fine operator is given
    DeviceCsr<T> P, R;
    DevBuf<T> dinv, x, xo, b, r, d, q;  // x: the iterate of the cycle on this level, xo: its result (the last smoothing kernel writes out of place)
    double lmax = 2.0;
};

// Hands the coarse part of a V-cycle to another solver (the fp64 tail of the mixed-precision cycle)
template <class T>
struct CoarseHandoff {
    virtual ~CoarseHandoff() = default;
    virtual void solve(const T* b, T* x, const PcgState* st, cudaStream_t s) = 0;  // x = V(b) on level `handoff_level`
    virtual int launches() const = 0;
};

// The V-cycle. Schedule per level l (degree k_l Chebyshev smoothing, every smoother step fused into the operator's SpMV, csr.cuh SpmvMode):
//     pre   : k_l fused steps from a zero guess (the start direction dinv b / theta was written by whoever produced b: the hand-over
//             kernel on level 0, the restriction above on the others); the last one stores no further direction
//     down  : b_{l+1} = R r fused with the child's start direction
//     child : recursion; dense inverse on the coarsest level (optionally the small levels as ONE cluster kernel, vtail.cuh: measured
//             slower than separate launches on the B200 and off by default)
//     up    : x += P x_{l+1}
//     post  : residual + k_l - 1 steps, the last of them writing the result out of place -- k_l SpMVs, not k_l + 1: the step that
//             would only refresh r and the next direction after the last update is dead code and is not run
// i.e. 2 k_l + 2 launches per level (4 + 2 k_l in round 1).
template <class T>
struct AmgPrecond : Precond<T> {
    using S = typename HostScalar<T>::type;
    AmgParams prm;
    std::vector<AmgLevelDev<T>> lv;
    LinOp<T>* fine_op = nullptr;  // level-0 operator (e.g. the face stencil); falls back to lv[0].A
    std::unique_ptr<CsrOp<T>> own_op0;
    DevBuf<T> coarse_inv;
    int coarse_n = 0;
    bool coarse_dense = false;
    double op_complexity = 0, grid_complexity = 0;
    ReduceWs dummy_ws{nullptr, nullptr};
    CoarseHandoff<T>* handoff = nullptr;  // levels >= handoff_level are solved by it (lv[handoff_level] only holds x and b)
    int handoff_level = -1;
    // coarse tail (vtail.cuh)
    bool prepared = false;
    bool allow_tail = true;
    int first_level = 0;              // levels below it hold no work buffers (the partitioned mode runs level 0 itself)
    int tail_start = -1;              // first level inside the tail kernel; -1: no tail
    TailDesc<T> tail{};
    std::vector<DevBuf<int>> tail_long;

    int degree(int l) const { return (l > 0 && prm.cheb_degree_coarse > 0) ? prm.cheb_degree_coarse : prm.cheb_degree; }
    double theta(int l) const { return 0.5 * (lv[l].lmax + lv[l].lmax / prm.cheb_ratio); }
    double delta(int l) const { return 0.5 * (lv[l].lmax - lv[l].lmax / prm.cheb_ratio); }
    double inv_theta0() const { return 1.0 / theta(0); }
    // where the zero-guess start direction of level l lives: in the iterate buffer itself for degree 1 (x = d0 needs no copy then)
    T* start_dir(int l, T* x) { return degree(l) == 1 ? x : lv[l].d.p; }
    LinOp<T>* op(int l) { return (l == 0 && fine_op) ? fine_op : nullptr; }

    // work buffers + the tail descriptor; called before the first cycle (and before any graph capture: it allocates and synchronises)
    void prepare(cudaStream_t s) {
        if (prepared) return;
        prepared = true;
        const int nl = (int)lv.size();
        for (int l = first_level; l < nl; l++) {
            AmgLevelDev<T>& L = lv[l];
            if (handoff && l == handoff_level) { if (!L.xo.p) L.xo.alloc(L.n); break; }
            if (!L.x.p) L.x.alloc(L.n);
            if (!L.xo.p) L.xo.alloc(L.n);
            if (!L.b.p && l > 0) L.b.alloc(L.n);
            if (!L.r.p) L.r.alloc(L.n);
            if (!L.d.p) L.d.alloc(L.n);
            if (!L.q.p) L.q.alloc(L.n);
        }
        build_tail(s);
    }

    static int lanes_for(double mean) { return mean <= 3 ? 2 : (mean <= 6 ? 4 : (mean <= 12 ? 8 : (mean <= 24 ? 16 : 32))); }
    // row classes of one operand of the tail kernel (vtail.cuh TailMat)
    TailMat<T> tail_mat(const DeviceCsr<T>& A, cudaStream_t s) {
        TailMat<T> M{};
        M.ptr = A.ptr.p; M.col = A.col.p; M.val = A.val.p;
        const double mean = A.rows > 0 ? (double)A.nnz / A.rows : 1.0;
        M.gmax = lanes_for(mean);
        M.cut = std::max(kTailMedRowMin, (int)(4.0 * mean));
        if (A.rows <= 0 || !A.ptr.p) return M;
        std::vector<int> ptr((size_t)A.rows + 1), med, huge;
        MK_CUDA(cudaMemcpyAsync(ptr.data(), A.ptr.p, sizeof(int) * ((size_t)A.rows + 1), cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < A.rows; i++) {
            const int len = ptr[i + 1] - ptr[i];
            if (len > kTailHugeRow) huge.push_back(i);
            else if (len > M.cut) med.push_back(i);
        }
        auto put = [&](const std::vector<int>& v, const int*& out, int& cnt) {
            cnt = (int)v.size();
            out = nullptr;
            if (v.empty()) return;
            tail_long.emplace_back(v.size());
            tail_long.back().upload(v.data(), v.size(), s);
            MK_CUDA(cudaStreamSynchronize(s));
            out = tail_long.back().p;
        };
        put(med, M.med, M.n_med);
        put(huge, M.huge, M.n_huge);
        return M;
    }
    void build_tail(cudaStream_t s) {
        tail_start = -1;
        const int nl = (int)lv.size();
        const int thresh = tail_rows_threshold();
        if (thresh <= 0 || handoff || !allow_tail || !coarse_dense || nl < 2) return;
        int t0 = nl - 1;
        while (t0 - 1 >= std::max(1, first_level) && lv[t0 - 1].n <= thresh) t0--;
        if (nl - t0 > kTailMaxLevels) t0 = nl - kTailMaxLevels;
        tail.nl = nl - t0;
        tail.coarse_n = coarse_n;
        tail.coarse_inv = coarse_inv.p;
        for (int l = t0; l < nl; l++) {
            AmgLevelDev<T>& L = lv[l];
            TailLevel<T>& D = tail.lv[l - t0];
            D = TailLevel<T>{};
            D.n = L.n;
            D.deg = degree(l);
            D.dinv = L.dinv.p;
            D.x = L.x.p; D.xo = L.xo.p; D.b = L.b.p; D.r = L.r.p; D.d = L.d.p; D.q = L.q.p;
            D.inv_theta = 1.0 / theta(l);
            D.delta = delta(l);
            D.sigma = theta(l) / delta(l);
            if (l + 1 < nl) {
                D.A = tail_mat(L.A, s);
                D.P = tail_mat(L.P, s);
                D.R = tail_mat(L.R, s);
            }
        }
        tail_cluster_size<T>();  // attribute + occupancy query outside any graph capture
        tail_start = t0;
        if (getenv("MK_VERBOSE")) {
            fprintf(stderr, "[mk] V-cycle tail: levels %d..%d (", t0, nl - 1);
            for (int l = t0; l < nl; l++) fprintf(stderr, "%d%s", lv[l].n, l + 1 < nl ? " " : "");
            fprintf(stderr, " rows) in one cluster of %d CTAs\n", tail_cluster_size<T>());
        }
    }

    // ---- fused smoother kernels of level l: the external fine operator on level 0, the level's own CSR operator otherwise
    void k_step(int l, const T* d_in, T* d_out, T* x, const T* b, double c1, double c2, int first, const PcgState* st, cudaStream_t s) {
        AmgLevelDev<T>& L = lv[l];
        if (op(l)) op(l)->cheb_step(d_in, d_out, L.r.p, x, b, L.dinv.p, c1, c2, first, st, s);
        else spmv_cheb_step<T>(L.A, 0.0, d_in, d_out, L.r.p, x, b, L.dinv.p, c1, c2, first, st, s);
    }
    void k_resid(int l, const T* x, const T* b, T* d_out, const PcgState* st, cudaStream_t s) {
        AmgLevelDev<T>& L = lv[l];
        if (op(l)) op(l)->cheb_resid(x, b, L.r.p, d_out, L.dinv.p, 1.0 / theta(l), st, s);
        else spmv_cheb_resid<T>(L.A, 0.0, x, b, L.r.p, d_out, L.dinv.p, 1.0 / theta(l), st, s);
    }
    void k_resid_last(int l, const T* x, const T* b, T* xo, const void* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
        AmgLevelDev<T>& L = lv[l];
        if (op(l)) op(l)->cheb_resid_last(x, b, xo, L.dinv.p, 1.0 / theta(l), r_hi, st, ws, s);
        else spmv_cheb_resid_last<T>(L.A, 0.0, x, b, xo, L.dinv.p, 1.0 / theta(l), (const typename HiOf<T>::type*)r_hi, st, ws, s);
    }
    void k_step_last(int l, const T* d_in, T* x, T* xo, double c1, double c2, const void* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
        AmgLevelDev<T>& L = lv[l];
        if (op(l)) op(l)->cheb_step_last(d_in, L.r.p, x, xo, L.dinv.p, c1, c2, r_hi, st, ws, s);
        else spmv_cheb_step_last<T>(L.A, 0.0, d_in, L.r.p, x, xo, L.dinv.p, c1, c2, (const typename HiOf<T>::type*)r_hi, st, ws, s);
    }

    // Generic Chebyshev smoothing x <- x + p(A) (b - A x) with `deg` steps (only the non-dense coarsest level uses it: 8 steps stand in
    // for the solve there). Leaves r = b - A x in L.r.
    void smooth_generic(int l, const T* b, T* x, int deg, const PcgState* st, cudaStream_t s) {
        AmgLevelDev<T>& L = lv[l];
        const double th = theta(l), de = delta(l), sigma = th / de;
        double rho = 1.0 / sigma;
        T* d0 = L.d.p;
        T* d1 = L.q.p;
        k_start_dir<T><<<stream_grid(L.n), kThreads, 0, s>>>(L.n, b, L.dinv.p, d0, 1.0 / th, st);
        for (int k = 0; k < deg; k++) {
            const double rho1 = 1.0 / (2.0 * sigma - rho);
            k_step(l, d0, d1, x, b, rho1 * rho, 2.0 * rho1 / de, k == 0 ? 1 : 0, st, s);
            std::swap(d0, d1);
            rho = rho1;
        }
        MK_CHECK_LAUNCH();
    }

    void coarsest(const T* b, T* xo, const PcgState* st, cudaStream_t s) {
        const int l = (int)lv.size() - 1;
        AmgLevelDev<T>& L = lv[l];
        if (coarse_dense) {
            k_dense_matvec<T><<<(L.n * 32 + kThreads - 1) / kThreads, kThreads, 0, s>>>(L.n, coarse_inv.p, b, xo, st);
            MK_CHECK_LAUNCH();
        } else {
            smooth_generic(l, b, xo, 8, st, s);
        }
    }

    // The cycle on level l. On entry b holds the right-hand side and start_dir(l, x) the direction dinv b / theta; on return xo = V(b).
    // r_hi (level 0 of the mixed-precision cycle only): the last kernel also finishes the PCG reduction against the fp64 residual.
    void descend(int l, const T* b, T* x, T* xo, const void* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
        const int nl = (int)lv.size();
        AmgLevelDev<T>& L = lv[l];
        AmgLevelDev<T>& C = lv[l + 1];
        const int deg = degree(l);
        const double de = delta(l), sigma = theta(l) / de;
        // pre-smoothing from a zero guess
        {
            T* d_cur = start_dir(l, x);
            double rho = 1.0 / sigma;
            for (int k = 0; k < deg; k++) {
                const double rho1 = 1.0 / (2.0 * sigma - rho);
                T* d_next = (k == deg - 1) ? nullptr : (d_cur == L.d.p ? L.q.p : L.d.p);
                k_step(l, d_cur, d_next, x, b, rho1 * rho, 2.0 * rho1 / de, k == 0 ? 1 : 0, st, s);
                d_cur = d_next;
                rho = rho1;
            }
        }
        // restriction (+ the child's start direction) and the child
        const bool child_handoff = handoff && l + 1 == handoff_level;
        const bool child_last = l + 2 == nl;
        if (child_handoff || (child_last && tail_start != l + 1)) {
            spmv_plain(L.R, L.r.p, C.b.p, false, st, s);
        } else {
            spmv_restrict_init<T>(L.R, L.r.p, C.b.p, start_dir(l + 1, C.x.p), C.dinv.p, 1.0 / theta(l + 1), st, s);
        }
        if (child_handoff) handoff->solve(C.b.p, C.xo.p, st, s);
        else if (tail_start == l + 1) launch_vcycle_tail<T>(tail, C.xo.p, st, s);
        else if (child_last) coarsest(C.b.p, C.xo.p, st, s);
        else descend(l + 1, C.b.p, C.x.p, C.xo.p, nullptr, st, ws, s);
        // prolongation and post-smoothing
        spmv_plain(L.P, C.xo.p, x, true, st, s);
        if (deg == 1) {
            k_resid_last(l, x, b, xo, r_hi, st, ws, s);
        } else {
            k_resid(l, x, b, L.d.p, st, s);
            T* d_cur = L.d.p;
            double rho = 1.0 / sigma;
            for (int k = 0; k + 1 < deg; k++) {
                const double rho1 = 1.0 / (2.0 * sigma - rho);
                if (k + 2 == deg) {
                    k_step_last(l, d_cur, x, xo, rho1 * rho, 2.0 * rho1 / de, r_hi, st, ws, s);
                } else {
                    T* d_next = d_cur == L.d.p ? L.q.p : L.d.p;
                    k_step(l, d_cur, d_next, x, nullptr, rho1 * rho, 2.0 * rho1 / de, 0, st, s);
                    d_cur = d_next;
                }
                rho = rho1;
            }
        }
        MK_CHECK_LAUNCH();
    }

    // x_out = V(b) on level l (b and x_out are the caller's; x_out must not be lv[l].x)
    void vcycle(int l, const T* b, T* x_out, const PcgState* st, cudaStream_t s) {
        prepare(s);
        AmgLevelDev<T>& L = lv[l];
        if (l == (int)lv.size() - 1) { coarsest(b, x_out, st, s); return; }
        k_start_dir<T><<<stream_grid(L.n), kThreads, 0, s>>>(L.n, b, L.dinv.p, start_dir(l, L.x.p), 1.0 / theta(l), st);
        MK_CHECK_LAUNCH();
        if (l == tail_start && b == L.b.p) launch_vcycle_tail<T>(tail, x_out, st, s);
        else descend(l, b, L.x.p, x_out, nullptr, const_cast<PcgState*>(st), dummy_ws, s);
    }

    void apply(const T* r, T* z, const PcgState* st, cudaStream_t s) override { vcycle(0, r, z, st, s); }

    // rows and stored entries of A_l and P_l per level (level 0 behind an external fine operator reports 0 entries for A)
    void describe(std::vector<int64_t>& rows, std::vector<int64_t>& nnz_a, std::vector<int64_t>& nnz_p) const {
        rows.clear(); nnz_a.clear(); nnz_p.clear();
        for (size_t l = 0; l < lv.size(); l++) {
            rows.push_back(lv[l].n);
            nnz_a.push_back(lv[l].A.ptr.p ? lv[l].A.nnz : 0);
            nnz_p.push_back(l + 1 < lv.size() ? lv[l].P.nnz : 0);
        }
    }

    int launches_per_apply() const override { return 1 + launches_from_level(0); }
    // kernels of descend(l): pre, restrict, child, prolong, post
    int launches_from_level(int l) const {
        const int nl = (int)lv.size();
        if (l >= nl - 1) return coarse_dense ? 1 : 9;
        if (l == tail_start) return 1;
        int child;
        if (handoff && l + 1 == handoff_level) child = handoff->launches();
        else child = launches_from_level(l + 1);
        return 2 * degree(l) + 2 + child;
    }
};

// ---- mixed precision: the V-cycle in fp32 storage and arithmetic, wrapped as a preconditioner of the fp64 PCG ----
// The V-cycle is the bandwidth of an iteration (its operators and work vectors are ~80 % of the bytes moved), while PCG only needs
// z ~ M^{-1} r to a few digits: halving its bytes costs nothing in the final accuracy (residuals, dots and updates stay fp64; the true
// residual is re-checked at exit as before). r is scaled to roughly unit 2-norm on the way down (its norm falls by 1e-10 over a solve);
// the cycle is linear, so the factor (PcgState::zscale) cancels exactly on the way back and need not be the current norm.

template <class T, class Tlo>
static __global__ void k_widen_vector(int n, const Tlo* __restrict__ in, T* __restrict__ out, const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_rw(out + i, to_high<T, Tlo>(ld_ro(in + i), 1.0));
}
template <class T, class Tlo>
static __global__ void k_narrow_vector_st(int n, const T* __restrict__ in, Tlo* __restrict__ out, const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_rw(out + i, to_low<Tlo, T>(ld_ro(in + i), 1.0));
}

// Optional fp64 tail: levels >= kMixedHandoffLevel of the mixed cycle run in fp64 (they are a few percent of the bytes and carry the low
// modes). Measured at 1M vertices it does NOT pay: the IGM operator needs 1,378 inner iterations with an fp64 cycle and 1,559 with an
// all-fp32 one, and still 1,555 with the tail from level 2 -- the loss comes from the fine levels (the 1e-4 shift against row sums that
// cancel to ~1e-6 of the diagonal in fp32), not from the coarse solve -- while the conversions cost 4 % of the step. Hence disabled
// (hand-off level beyond any hierarchy depth); the fp32 cycle is still the faster one overall (field solve -23 %, IGM solve -3 %).
constexpr int kMixedHandoffLevel = 99;
// MK_FP64_TAIL=k overrides it at run time: k > 0 = hand-off level k, k < 0 counts from the coarsest level (-1: only the coarsest dense
// solve runs in fp64). The coarsest IGM operator has a condition number of 5e6 (4.8e8 with typed aggregation) at 1M vertices, beyond
// what an fp32 inverse resolves.
inline int mixed_handoff_level(int nl) {
    int k = kMixedHandoffLevel;
    if (const char* e = getenv("MK_FP64_TAIL")) k = atoi(e);
    if (k < 0) k = nl + k;
    return (k >= 1 && k < nl) ? k : -1;  // -1: everything in fp32
}
template <class T>
struct Fp64Tail : CoarseHandoff<typename LowOf<T>::type> {
    using Tlo = typename LowOf<T>::type;
    AmgPrecond<T>* hi = nullptr;
    int level = 0, n = 0;
    DevBuf<T> b, x;
    void solve(const Tlo* b_lo, Tlo* x_lo, const PcgState* st, cudaStream_t s) override {
        const int grid = stream_grid(n);
        k_widen_vector<T, Tlo><<<grid, kThreads, 0, s>>>(n, b_lo, b.p, st);
        hi->vcycle(level, b.p, x.p, st, s);
        k_narrow_vector_st<T, Tlo><<<grid, kThreads, 0, s>>>(n, x.p, x_lo, st);
        MK_CHECK_LAUNCH();
    }
    int launches() const override { return 3 + hi->launches_from_level(level); }
};

// hand-over: b = r / zscale (fp32), start direction d0 = dinv b / theta; zscale = |r|
template <class T, class Tlo>
static __global__ void __launch_bounds__(kThreads) k_precond_in_init(int n, const T* __restrict__ r, const Tlo* __restrict__ dinv, Tlo* __restrict__ b,
                                                              Tlo* __restrict__ d0, double inv_theta, PcgState* st) {
    if (st->done) return;
    const double rr = st->rr;
    const double s = rr > 0.0 ? rsqrt(rr) : 1.0;
    if (blockIdx.x == 0 && threadIdx.x == 0) st->zscale = rr > 0.0 ? sqrt(rr) : 1.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Tlo bi = to_low<Tlo, T>(ld_ro(r + i), s);
        st_rw(b + i, bi);
        st_rw(d0 + i, ST<Tlo>::scale(inv_theta, ld_ro(dinv + i) * bi));
    }
}
// hand-back: z = zscale x (fp64)
template <class T, class Tlo>
static __global__ void __launch_bounds__(kThreads) k_precond_out(int n, const Tlo* __restrict__ x, T* __restrict__ z, const PcgState* st) {
    if (st->done) return;
    const double s = st->zscale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_rw(z + i, to_high<T, Tlo>(ld_ro(x + i), s));
}

// The PCG update fused with the hand-over (solver.cuh Precond::update_xr_handover):
//     x += alpha p ; r -= alpha q ; rr = <r, r> ; b_lo = r / zscale ; d0 = dinv b_lo / theta
// zscale is the residual norm of the PREVIOUS iteration (the current one is only known after this kernel's own reduction; any factor
// of the right magnitude serves, see above). Every CTA reads st->rr before it takes its reduction ticket and the last CTA overwrites
// it only after all tickets are taken, so all CTAs see the same value. 48 (real) / 96 (complex) bytes of fp64 traffic per row plus
// 12 / 24 of fp32, against the same fp64 traffic + a second pass over r in the two kernels this replaces.
template <class T, class Tlo>
static __global__ void __launch_bounds__(kThreads) k_update_xr_handover(int n, T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p,
                                                                 const T* __restrict__ q, const Tlo* __restrict__ dinv, Tlo* __restrict__ b,
                                                                 Tlo* __restrict__ d0, double inv_theta, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    const double alpha = st->alpha;
    const double rr_old = st->rr;
    const double s = rr_old > 0.0 ? rsqrt(rr_old) : 1.0;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const T pi = ld_ro(p + i), qi = ld_ro(q + i);
        T xi = ld_rw(x + i), ri = ld_rw(r + i);
        xi = xi + ST<T>::scale(alpha, pi);
        ri = ri - ST<T>::scale(alpha, qi);
        st_rw(x + i, xi);
        st_rw(r + i, ri);
        acc += ST<T>::abs2(ri);
        const Tlo bi = to_low<Tlo, T>(ri, s);
        st_rw(b + i, bi);
        st_rw(d0 + i, ST<Tlo>::scale(inv_theta, ld_ro(dinv + i) * bi));
    }
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* t) {
        st->zscale = rr_old > 0.0 ? sqrt(rr_old) : 1.0;
        st->rr = t[0];
        st->iter += 1;
        if (t[0] <= st->thresh) st->done = 1;
        else if (st->iter >= st->maxit) st->done = 3;
        else if (!(t[0] == t[0])) st->done = 2;
    });
}

// p = zscale z_lo + beta p   (the hand-back fused with the direction update: z never exists in fp64)
template <class T, class Tlo>
static __global__ void __launch_bounds__(kThreads) k_update_p_lo(int n, T* __restrict__ p, const Tlo* __restrict__ z, const PcgState* st) {
    if (st->done) return;
    const double beta = st->beta, s = st->zscale;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        st_rw(p + i, to_high<T, Tlo>(ld_ro(z + i), s) + ST<T>::scale(beta, ld_rw(p + i)));
}

template <class T>
struct MixedPrecond : Precond<T> {
    using Tlo = typename LowOf<T>::type;
    std::unique_ptr<AmgPrecond<T>> hi;    // fp64 hierarchy: only its levels >= the hand-off level are kept
    Fp64Tail<T> tail;
    std::unique_ptr<AmgPrecond<Tlo>> lo;
    std::unique_ptr<LinOp<Tlo>> fine_lo;  // level-0 operator in fp32 storage (owned here)
    DevBuf<Tlo> b_lo, z_lo, zo_lo;        // right-hand side, iterate and result of level 0
    int n = 0;
    bool single_level() const { return lo->lv.size() < 2; }
    void prepare(cudaStream_t s) {
        lo->prepare(s);
        if (hi) hi->prepare(s);
    }
    // the cycle on the residual already handed over (b_lo and the start direction are in place); result in zo_lo
    void cycle(const void* r_hi, PcgState* st, ReduceWs ws, cudaStream_t s) {
        lo->descend(0, b_lo.p, z_lo.p, zo_lo.p, r_hi, st, ws, s);
    }
    void handover(const T* r, PcgState* st, cudaStream_t s) {
        AmgLevelDev<Tlo>& L0 = lo->lv[0];
        k_precond_in_init<T, Tlo><<<stream_grid(n), kThreads, 0, s>>>(n, r, L0.dinv.p, b_lo.p, lo->start_dir(0, z_lo.p), lo->inv_theta0(), st);
    }
    void apply(const T* r, T* z, const PcgState* st_c, cudaStream_t s) override {
        PcgState* st = const_cast<PcgState*>(st_c);
        prepare(s);
        handover(r, st, s);
        if (single_level()) lo->coarsest(b_lo.p, zo_lo.p, st, s);
        else cycle(nullptr, st, ReduceWs{nullptr, nullptr}, s);
        k_precond_out<T, Tlo><<<stream_grid(n), kThreads, 0, s>>>(n, zo_lo.p, z, st);
        MK_CHECK_LAUNCH();
    }
    int launches_per_apply() const override { return 2 + lo->launches_from_level(0); }
    // fused PCG body
    bool fused_pcg() const override { return !single_level(); }
    void update_xr_handover(int n_, T* x, T* r, const T* p, const T* q, PcgState* st, ReduceWs ws, cudaStream_t s) override {
        AmgLevelDev<Tlo>& L0 = lo->lv[0];
        k_update_xr_handover<T, Tlo><<<stream_grid(n_), kThreads, 0, s>>>(n_, x, r, p, q, L0.dinv.p, b_lo.p, lo->start_dir(0, z_lo.p), lo->inv_theta0(), st, ws);
    }
    void cycle_dot(const T* r, PcgState* st, ReduceWs ws, cudaStream_t s) override { cycle(r, st, ws, s); }
    void update_p(int n_, T* p, const PcgState* st, cudaStream_t s) override {
        k_update_p_lo<T, Tlo><<<stream_grid(n_), kThreads, 0, s>>>(n_, p, zo_lo.p, st);
    }
    int launches_fused() const override { return 2 + lo->launches_from_level(0); }
    void cycle_only(PcgState* st, cudaStream_t s) override { cycle(nullptr, st, ReduceWs{nullptr, nullptr}, s); }
    const void* lo_result() const override { return zo_lo.p; }
};

// ---- live timing of the cycle, level by level (bench.py / profiles: where an iteration's time goes, measured warm and back to back as
// in a solve, not under a profiler): `body` is captured into a CUDA graph and replayed `reps` times between two events.
template <class F>
inline double time_graph_us(cudaStream_t s, int reps, F body) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    MK_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    body();
    MK_CUDA(cudaStreamEndCapture(s, &g));
    MK_CUDA(cudaGraphInstantiate(&ge, g, 0));
    cudaGraphDestroy(g);
    for (int i = 0; i < 3; i++) MK_CUDA(cudaGraphLaunch(ge, s));
    cudaEvent_t e0, e1;
    MK_CUDA(cudaEventCreate(&e0));
    MK_CUDA(cudaEventCreate(&e1));
    MK_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) MK_CUDA(cudaGraphLaunch(ge, s));
    MK_CUDA(cudaEventRecord(e1, s));
    MK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaGraphExecDestroy(ge);
    return (double)ms * 1e3 / reps;
}

// us[0] = one application of the preconditioner `M` to a vector of length n (hand-over + cycle + hand-back), us[l] = the cycle from
// level l downwards (start direction + descend) for every level l >= 1 of `A`
template <class T, class TA>
inline std::vector<double> profile_cycle(Precond<T>& M, AmgPrecond<TA>& A, int n, int reps, cudaStream_t s) {
    DevBuf<T> r(n), z(n);
    r.zero(s);
    DevBuf<PcgState> unit(1);
    PcgState h;
    std::memset(&h, 0, sizeof(h));
    h.rr = 1.0; h.rho = 1.0; h.bnorm2 = 1.0; h.zscale = 1.0;
    unit.upload(&h, 1, s);
    M.apply(r.p, z.p, unit.p, s);  // allocates the work buffers outside any capture
    MK_CUDA(cudaStreamSynchronize(s));
    std::vector<double> us;
    us.push_back(time_graph_us(s, reps, [&] { M.apply(r.p, z.p, unit.p, s); }));
    for (int l = 1; l < (int)A.lv.size(); l++)
        us.push_back(time_graph_us(s, reps, [&] { A.vcycle(l, A.lv[l].b.p, A.lv[l].xo.p, unit.p, s); }));
    // the two fused PCG vector kernels either side of the cycle (alpha = beta = 0: the vectors stay what they are)
    if (M.fused_pcg()) {
        DevBuf<T> x(n), p(n), q(n);
        x.zero(s); p.zero(s); q.zero(s);
        DevBuf<double> partials((size_t)kMaxPartials * 4);
        DevBuf<unsigned int> ticket(1);
        ticket.zero(s);
        h.thresh = -1.0;
        h.maxit = 0x7fffffff;
        unit.upload(&h, 1, s);
        MK_CUDA(cudaStreamSynchronize(s));
        ReduceWs ws{partials.p, ticket.p};
        us.push_back(time_graph_us(s, reps, [&] { M.update_xr_handover(n, x.p, r.p, p.p, q.p, unit.p, ws, s); }));
        us.push_back(time_graph_us(s, reps, [&] { M.cycle_dot(r.p, unit.p, ws, s); }));
        us.push_back(time_graph_us(s, reps, [&] { M.update_p(n, p.p, unit.p, s); }));
    }
    return us;
}

template <class T, class Tlo>
static __global__ void k_narrow_vector(int n, const T* __restrict__ in, Tlo* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st_rw(out + i, to_low<Tlo, T>(ld_ro(in + i), 1.0));
}

// fp32 copy of the fine levels of a device hierarchy (operators converted, work vectors allocated in fp32); the coarse levels stay in
// `hi`, which the result owns. fine_lo: the level-0 operator in fp32 (or null when level 0 of `hi` holds its own CSR operator).
template <class T>
inline std::unique_ptr<MixedPrecond<T>> lower_precision(cudaStream_t s, std::unique_ptr<AmgPrecond<T>> hi_in,
                                                        std::unique_ptr<LinOp<typename LowOf<T>::type>> fine_lo) {
    using Tlo = typename LowOf<T>::type;
    auto mp = std::make_unique<MixedPrecond<T>>();
    mp->hi = std::move(hi_in);
    AmgPrecond<T>& hi = *mp->hi;
    mp->lo = std::make_unique<AmgPrecond<Tlo>>();
    AmgPrecond<Tlo>& lo = *mp->lo;
    lo.prm = hi.prm;
    mp->fine_lo = std::move(fine_lo);
    lo.fine_op = mp->fine_lo.get();
    lo.op_complexity = hi.op_complexity;
    lo.grid_complexity = hi.grid_complexity;
    const int nl = (int)hi.lv.size();
    const int K = mixed_handoff_level(nl);  // -1: hierarchy too shallow / no fp64 tail, everything in fp32
    const int n_lo = K > 0 ? K + 1 : nl;
    lo.lv.resize(n_lo);
    for (int l = 0; l < n_lo; l++) {
        AmgLevelDev<T>& H = hi.lv[l];
        AmgLevelDev<Tlo>& L = lo.lv[l];
        const int n = H.n;
        L.n = n;
        L.lmax = H.lmax;
        L.x.alloc(n); L.b.alloc(n);
        if (K > 0 && l == K) break;  // the hand-off level only exchanges x and b
        if (!(l == 0 && lo.fine_op)) lower_csr<T, Tlo>(H.A, L.A, s);
        if (l + 1 < nl) {
            lower_csr<T, Tlo>(H.P, L.P, s);
            lower_csr<T, Tlo>(H.R, L.R, s);
        }
        L.dinv.alloc(n);
        k_narrow_vector<T, Tlo><<<stream_grid(n), kThreads, 0, s>>>(n, H.dinv.p, L.dinv.p);
        L.r.alloc(n); L.d.alloc(n); L.q.alloc(n);
        MK_CHECK_LAUNCH();
    }
    if (K > 0) {
        mp->tail.hi = mp->hi.get();
        mp->tail.level = K;
        mp->tail.n = hi.lv[K].n;
        mp->tail.b.alloc(mp->tail.n);
        mp->tail.x.alloc(mp->tail.n);
        lo.handoff = &mp->tail;
        lo.handoff_level = K;
        hi.allow_tail = false;  // its fine levels are dropped below
        MK_CUDA(cudaStreamSynchronize(s));
        for (int l = 0; l < K; l++) hi.lv[l] = AmgLevelDev<T>();  // the fine fp64 levels are no longer needed
        hi.lv[0].n = lo.lv[0].n;
    } else {
        lo.coarse_n = hi.coarse_n;
        lo.coarse_dense = hi.coarse_dense;
        if (hi.coarse_dense) {
            const size_t m = hi.coarse_inv.n;
            lo.coarse_inv.alloc(m);
            k_narrow_vector<T, Tlo><<<stream_grid((int64_t)m), kThreads, 0, s>>>((int)m, hi.coarse_inv.p, lo.coarse_inv.p);
            MK_CHECK_LAUNCH();
        }
        MK_CUDA(cudaStreamSynchronize(s));
        mp->hi.reset();
    }
    mp->n = lo.lv[0].n;
    mp->b_lo.alloc(mp->n);
    mp->z_lo.alloc(mp->n);
    mp->zo_lo.alloc(mp->n);
    MK_CUDA(cudaStreamSynchronize(s));
    return mp;
}

// fp32 copy of the levels >= `from` of an fp64 hierarchy as a hierarchy of its own (its level 0 = level `from`, a materialised CSR operator):
// the replicated coarse part of the partitioned preconditioner (dist.cuh) runs the same fp32 cycle as the single-GPU solver.
template <class T>
inline std::unique_ptr<AmgPrecond<typename LowOf<T>::type>> lower_tail(cudaStream_t s, AmgPrecond<T>& hi, int from) {
    using Tlo = typename LowOf<T>::type;
    auto lo = std::make_unique<AmgPrecond<Tlo>>();
    lo->prm = hi.prm;
    lo->prm.cheb_degree = hi.prm.cheb_degree_coarse > 0 ? hi.prm.cheb_degree_coarse : hi.prm.cheb_degree;  // its level 0 is a coarse level
    const int nl = (int)hi.lv.size();
    lo->lv.resize(nl - from);
    for (int l = from; l < nl; l++) {
        AmgLevelDev<T>& H = hi.lv[l];
        AmgLevelDev<Tlo>& L = lo->lv[l - from];
        const int n = H.n;
        L.n = n;
        L.lmax = H.lmax;
        L.x.alloc(n); L.b.alloc(n); L.xo.alloc(n);
        lower_csr<T, Tlo>(H.A, L.A, s);
        if (l + 1 < nl) {
            lower_csr<T, Tlo>(H.P, L.P, s);
            lower_csr<T, Tlo>(H.R, L.R, s);
        }
        L.dinv.alloc(n);
        k_narrow_vector<T, Tlo><<<stream_grid(n), kThreads, 0, s>>>(n, H.dinv.p, L.dinv.p);
        L.r.alloc(n); L.d.alloc(n); L.q.alloc(n);
        MK_CHECK_LAUNCH();
    }
    lo->coarse_n = hi.coarse_n;
    lo->coarse_dense = hi.coarse_dense;
    if (hi.coarse_dense) {
        const size_t m = hi.coarse_inv.n;
        lo->coarse_inv.alloc(m);
        k_narrow_vector<T, Tlo><<<stream_grid((int64_t)m), kThreads, 0, s>>>((int)m, hi.coarse_inv.p, lo->coarse_inv.p);
        MK_CHECK_LAUNCH();
    }
    MK_CUDA(cudaStreamSynchronize(s));
    return lo;
}

// Typed aggregation of the IGM hierarchy (u- and v-type dofs never share an aggregate). It resolves the two near-kernel vectors better
// (Poisson solve 116 -> 100 iterations at 1M vertices), which only pays together with the constraint preconditioner (igm.cuh): under the
// plain projected preconditioner every added rounding constraint then costs 10-20 extra iterations (2215 against 1537 in total), with
// it the 1M-vertex rounding loop needs 698 instead of 841 (profiles/cprec_probe_r01.txt). Default: on iff MK_IGM_CPREC is not 0;
// MK_AMG_TYPED=0 / 1 forces it.
inline bool typed_aggregation() {
    if (const char* e = getenv("MK_AMG_TYPED")) return std::strcmp(e, "0") != 0;
    const char* c = getenv("MK_IGM_CPREC");
    return !(c && std::strcmp(c, "0") == 0);
}

// MK_PRECOND_FP32=0 keeps the V-cycle in fp64
inline bool precond_fp32() {
    const char* e = getenv("MK_PRECOND_FP32");
    return !(e && std::strcmp(e, "0") == 0);
}

template <class T>
inline void upload_host_csr(DeviceCsr<T>& D, const HostCsr<typename HostScalar<T>::type>& H, cudaStream_t s) {
    using S = typename HostScalar<T>::type;
    static_assert(sizeof(S) == sizeof(T), "layout");
    std::vector<T> v(H.val.size());
    std::memcpy(v.data(), H.val.data(), H.val.size() * sizeof(T));
    upload_csr<T>(D, H.rows, H.cols, H.ptr, H.col, v, s);
}

// Builds the hierarchy for A (host CSR) and uploads it. `fine_op` non-null: level 0 applies that operator instead of a
// CSR copy of A (A is still needed on the host for the Galerkin products).
template <class T>
inline std::unique_ptr<AmgPrecond<T>> build_amg(cudaStream_t s, HostCsr<typename HostScalar<T>::type> A,
                                                std::vector<typename HostScalar<T>::type> cand, bool use_cand, LinOp<T>* fine_op,
                                                const AmgParams& prm, std::vector<int> types = {}) {
    using S = typename HostScalar<T>::type;
    auto amg = std::make_unique<AmgPrecond<T>>();
    amg->prm = prm;
    amg->fine_op = fine_op;
    const double nnz0 = (double)A.nnz(), n0 = (double)A.rows;
    StageTimer tm("build_amg");
    std::vector<AmgHostLevel<S>> H = build_hierarchy<S>(std::move(A), std::move(cand), use_cand, prm, std::move(types));
    tm.lap("hierarchy (host)");
    amg->lv.resize(H.size());
    double nnz_sum = 0, n_sum = 0;
    for (size_t l = 0; l < H.size(); l++) {
        AmgLevelDev<T>& D = amg->lv[l];
        D.n = H[l].A.rows;
        D.lmax = H[l].lmax;
        nnz_sum += (double)H[l].A.nnz();
        n_sum += D.n;
        if (!(l == 0 && fine_op)) upload_host_csr<T>(D.A, H[l].A, s);
        if (l + 1 < H.size()) {
            upload_host_csr<T>(D.P, H[l].P, s);
            upload_host_csr<T>(D.R, H[l].R, s);
        }
        D.dinv.alloc(D.n);
        D.dinv.upload((const T*)H[l].dinv.data(), D.n, s);
        D.r.alloc(D.n); D.d.alloc(D.n); D.q.alloc(D.n);
        if (l > 0) { D.x.alloc(D.n); D.b.alloc(D.n); }
        MK_CUDA(cudaStreamSynchronize(s));
    }
    tm.lap("upload levels");
    amg->op_complexity = nnz_sum / nnz0;
    amg->grid_complexity = n_sum / n0;
    const AmgHostLevel<S>& last = H.back();
    amg->coarse_n = last.A.rows;
    if (last.A.rows <= 1500) {
        std::vector<S> inv = dense_inverse(last.A);
        amg->coarse_inv.alloc(inv.size());
        amg->coarse_inv.upload((const T*)inv.data(), inv.size(), s);
        MK_CUDA(cudaStreamSynchronize(s));
        amg->coarse_dense = true;
        tm.lap("coarse dense inverse");
    }
    if (tm.on) {
        fprintf(stderr, "[mk] build_amg: levels");
        for (auto& l : H) fprintf(stderr, " %d(%.1f)", l.A.rows, (double)l.A.nnz() / l.A.rows);
        fprintf(stderr, " op_cx %.2f grid_cx %.2f\n", amg->op_complexity, amg->grid_complexity);
    }
    return amg;
}

}  // namespace mk
