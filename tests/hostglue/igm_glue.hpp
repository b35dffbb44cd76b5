// Host stage that produces the inputs of the IGM hot path: restates the reference's index bookkeeping in
// include/metriko/core/igm/parameterization.h (compute_cut_mesh :126-152), parameterization_setup.ipp
// (he2matching :3-9, he2transidx :11-50, setup :52-170), parameterization_integ.ipp (:4-43 operator/field
// assembly, :88-98 result scatter), solver/matrix_ops.h:12-28 (row reduction) and
// iterative_rounding/iter_rounding_init.h:98-154 (UFull from the node constraints).
// OUT OF SCOPE of the GPU path: built once on the host, consumed identically by the
// oracle and the CUDA path. Closed meshes only.
#pragma once
#include <set>

#include "bag.hpp"
#include "field_glue.hpp"
#include "mesh.hpp"
#include "sparse.hpp"

namespace hg {

struct CutMesh {
    int nV = 0;
    std::vector<V3> pos;
    std::vector<std::array<int, 3>> idx;
    std::vector<int> cut2raw;
};

struct IgmStage {
    int N = 4, n = 2, nT = 0, nR = 0, nS = 0;
    CutMesh cut;
    std::vector<int> he2matching, he2transidx;
    Csr fullx2Nfunc;  // N*cut.nV x n*nR
    Csr UFull;        // n*nR x cols
    double length = 0;
};

// small dense integer N x N matrices (row major)
struct IMat {
    int N;
    std::vector<int> a;
    explicit IMat(int N_ = 0) : N(N_), a((size_t)N_ * N_, 0) {}
    int& operator()(int i, int j) { return a[(size_t)i * N + j]; }
    int operator()(int i, int j) const { return a[(size_t)i * N + j]; }
    static IMat identity(int N, int s = 1) { IMat m(N); for (int i = 0; i < N; i++) m(i, i) = s; return m; }
    IMat operator*(const IMat& o) const {
        IMat r(N);
        for (int i = 0; i < N; i++) for (int k = 0; k < N; k++) if ((*this)(i, k))
            for (int j = 0; j < N; j++) r(i, j) += (*this)(i, k) * o(k, j);
        return r;
    }
    bool any() const { for (int v : a) if (v) return true; return false; }
};

inline void assign_block(std::vector<Trip>& T, const IMat& m, int r0, int c0) {
    for (int i = 0; i < m.N; i++) for (int j = 0; j < m.N; j++) if (m(i, j) != 0) T.push_back({r0 + i, c0 + j, (double)m(i, j)});
}

// param.h:126-152
inline CutMesh compute_cut_mesh(const Hmesh& m, const std::vector<char>& seam) {
    CutMesh c;
    c.idx.assign(m.nF, {-1, -1, -1});
    for (int v = 0; v < m.nV; v++) {
        int start = m.vert2half[v], bgn = start, h = start;
        do {
            if (seam[m.edge[h]]) { bgn = h; break; }
            h = m.vnext_cw(h);
        } while (h != start);
        h = bgn;
        do {
            if (seam[m.edge[h]] || h == bgn) { c.pos.push_back(m.pos[v]); c.cut2raw.push_back(v); }
            int f = m.face[h];
            for (int j = 0; j < 3; j++) if (m.idx[f][j] == v) c.idx[f][j] = (int)c.pos.size() - 1;
            h = m.vnext_ccw(h);
        } while (h != bgn);
    }
    c.nV = (int)c.pos.size();
    return c;
}

// Dense helper: pick a maximal set of linearly independent rows (stand-in for the SparseQR-based
// reduce_to_linearly_independent, matrix_ops.h:12-28; row choice is unpinned, row space identical).
inline std::vector<int> independent_rows(const std::vector<std::vector<double>>& rows) {
    std::vector<int> keep;
    std::vector<std::vector<double>> basis;
    for (size_t i = 0; i < rows.size(); i++) {
        std::vector<double> r = rows[i];
        double n0 = 0;
        for (double x : r) n0 += x * x;
        n0 = std::sqrt(n0);
        if (n0 == 0) continue;
        for (int pass = 0; pass < 2; pass++)
            for (const auto& b : basis) {
                double d = 0;
                for (size_t k = 0; k < r.size(); k++) d += r[k] * b[k];
                for (size_t k = 0; k < r.size(); k++) r[k] -= d * b[k];
            }
        double n1 = 0;
        for (double x : r) n1 += x * x;
        n1 = std::sqrt(n1);
        if (n1 > 1e-9 * n0) {
            for (double& x : r) x /= n1;
            basis.push_back(r);
            keep.push_back((int)i);
        }
    }
    return keep;
}

// Kernel basis of a dense r x c matrix via full-pivot LU, laid out as Eigen::FullPivLU::kernel()
// does (ir_init.h:127-131 calls it): K = Q * [-U1^{-1} U2 ; I].
inline std::vector<std::vector<double>> fullpiv_kernel(std::vector<std::vector<double>> A, int r, int c) {
    std::vector<int> colperm(c);
    std::iota(colperm.begin(), colperm.end(), 0);
    int rank = 0;
    double maxpiv = 0;
    const int sz = std::min(r, c);
    for (int k = 0; k < sz; k++) {
        int pr = -1, pc = -1;
        double best = 0;
        for (int j = k; j < c; j++)  // column-major scan, first maximum wins
            for (int i = k; i < r; i++)
                if (std::fabs(A[i][j]) > best) { best = std::fabs(A[i][j]); pr = i; pc = j; }
        if (k == 0) maxpiv = best;
        if (best <= maxpiv * 1e-13 * std::max(r, c) || pr < 0) break;
        std::swap(A[k], A[pr]);
        for (int i = 0; i < r; i++) std::swap(A[i][k], A[i][pc]);
        std::swap(colperm[k], colperm[pc]);
        for (int i = k + 1; i < r; i++) {
            double f = A[i][k] / A[k][k];
            A[i][k] = 0;
            if (f != 0) for (int j = k + 1; j < c; j++) A[i][j] -= f * A[k][j];
        }
        rank++;
    }
    const int dk = c - rank;
    std::vector<std::vector<double>> K(c, std::vector<double>(dk, 0.0));
    // solve U1 * X = U2 (back substitution), kernel rows = -X in permuted order
    std::vector<std::vector<double>> X(rank, std::vector<double>(dk, 0.0));
    for (int j = 0; j < dk; j++)
        for (int i = rank - 1; i >= 0; i--) {
            double s = A[i][rank + j];
            for (int l = i + 1; l < rank; l++) s -= A[i][l] * X[l][j];
            X[i][j] = s / A[i][i];
        }
    for (int i = 0; i < rank; i++) for (int j = 0; j < dk; j++) K[colperm[i]][j] = -X[i][j];
    for (int k = 0; k < dk; k++) K[colperm[rank + k]][k] = 1.0;
    return K;
}

inline void put_csr(Bag& bag, const std::string& name, const Csr& A) {
    std::vector<int> dims = {A.rows, A.cols};
    bag.put(name + "_shape", dims);
    bag.put_raw(name + "_ptr", I64, A.ptr.data(), (int64_t)A.ptr.size());
    bag.put(name + "_col", A.col);
    bag.put(name + "_val", A.val);
}

inline void run_igm_stage(const Hmesh& raw, const FieldStage& fs, Bag& bag, double gridscale, IgmStage& st) {
    const int N = fs.N;
    if (N % 2 != 0) throw std::runtime_error("igm: only even N (sign symmetry) is supported");
    const int n = N / 2;
    st.N = N;
    st.n = n;
    const auto& seam = fs.seam;
    const auto& singular = fs.singular;
    const auto& matching = fs.cmatching;
    st.cut = compute_cut_mesh(raw, seam);
    const CutMesh& cut = st.cut;
    auto is_singular = [&](int v) { return singular[v] != 0; };

    // ---- he2matching (setup.ipp:3-9)
    st.he2matching.resize(raw.nH);
    for (int h = 0; h < raw.nH; h++) {
        int mm = (raw.canonical(h) ? -1 : 1) * matching[raw.edge[h]];
        st.he2matching[h] = mm < 0 ? (N + mm % N) % N : mm % N;
    }
    // ---- he2transidx (setup.ipp:11-50)
    st.he2transidx.assign(raw.nH, 32767);
    {
        std::vector<int> valence(raw.nV, 0);
        std::vector<char> claimed(raw.nE, 0);
        for (int e = 0; e < raw.nE; e++)
            if (seam[e]) { valence[raw.edge2vert[e][0]]++; valence[raw.edge2vert[e][1]]++; }
        int tid = 1;
        for (int v = 0; v < raw.nV; v++) {
            if ((valence[v] == 2 && !is_singular(v)) || valence[v] == 0) continue;
            int h0 = raw.vert2half[v], cH = h0;
            do {
                int cE = raw.edge[cH];
                if (seam[cE] && !claimed[cE]) {
                    st.he2transidx[cH] = tid;
                    st.he2transidx[raw.twin[cH]] = -tid;
                    claimed[cE] = 1;
                    int jv = raw.head[cH];
                    while (valence[jv] == 2 && !is_singular(jv)) {
                        int nh = -1;
                        int g0 = raw.vert2half[jv], g = g0;
                        do {
                            if (seam[raw.edge[g]] && !claimed[raw.edge[g]]) { nh = g; break; }
                            g = raw.vnext_ccw(g);
                        } while (g != g0);
                        if (nh < 0) break;  // closed seam loop returned to its start
                        st.he2transidx[nh] = tid;
                        st.he2transidx[raw.twin[nh]] = -tid;
                        claimed[raw.edge[nh]] = 1;
                        jv = raw.head[nh];
                    }
                    tid++;
                }
                cH = raw.vnext_ccw(cH);
            } while (cH != h0);
        }
        st.nT = tid - 1;
        for (int e = 0; e < raw.nE; e++)
            if (seam[e] && !claimed[e]) throw std::runtime_error("igm: seam edge without transition id (isolated seam loop)");
    }
    st.nR = raw.nV + st.nT;
    st.nS = 0;
    for (int v = 0; v < raw.nV; v++) st.nS += is_singular(v) ? 1 : 0;
    const int nR = st.nR, nT = st.nT;

    // ---- setup (setup.ipp:52-170)
    std::vector<IMat> constParm(N, IMat(N));
    IMat unit(N);
    for (int i = 0; i < N; i++) unit((i + 1) % N, i) = 1;
    constParm[0] = IMat::identity(N);
    for (int i = 1; i < N; i++) constParm[i] = unit * constParm[i - 1];

    std::vector<Trip> vT, cT;
    int currConstraint = 0;
    for (int v = 0; v < raw.nV; v++) {
        std::vector<IMat> permMats;
        std::vector<int> permIdcs;
        permMats.push_back(IMat::identity(N));
        permIdcs.push_back(v);
        int iVcut = -1;
        // find_first_seam_he (param.h:72): first seam half walking cw, else v.half()
        int start = raw.vert2half[v], bgn = start;
        {
            int h = start;
            do {
                if (seam[raw.edge[h]]) { bgn = h; break; }
                h = raw.vnext_cw(h);
            } while (h != start);
        }
        int h = bgn;
        do {
            int f = raw.face[h];
            int jVcut = -1;
            for (int j = 0; j < 3; j++) if (raw.idx[f][j] == v) jVcut = cut.idx[f][j];
            if (jVcut != iVcut) {
                iVcut = jVcut;
                for (size_t i = 0; i < permIdcs.size(); i++) assign_block(vT, permMats[i], N * iVcut, N * permIdcs[i]);
            }
            int hn = raw.twin[raw.prev[h]];
            if (seam[raw.edge[hn]]) {
                const IMat& p = constParm[st.he2matching[hn]];
                int t = st.he2transidx[hn];
                if (t > 0) {
                    for (auto& mm : permMats) mm = p * mm;
                    permMats.push_back(IMat::identity(N));
                    permIdcs.push_back(raw.nV + t - 1);
                } else {
                    permMats.push_back(IMat::identity(N, -1));
                    permIdcs.push_back(raw.nV - t - 1);
                    for (auto& mm : permMats) mm = p * mm;
                }
            }
            h = hn;
        } while (h != bgn);
        // constraint / singularity-from-transition detection
        std::set<int> tmp(permIdcs.begin(), permIdcs.end());
        std::vector<int> idcs(tmp.begin(), tmp.end());
        std::vector<IMat> mats(idcs.size(), IMat(N));
        for (size_t j = 0; j < idcs.size(); j++) {
            for (size_t k = 0; k < permIdcs.size(); k++)
                if (idcs[j] == permIdcs[k]) for (size_t q = 0; q < mats[j].a.size(); q++) mats[j].a[q] += permMats[k].a[q];
            if (idcs[j] == v) for (int q = 0; q < N; q++) mats[j](q, q) -= 1;
        }
        bool any = false;
        for (auto& mm : mats) any = any || mm.any();
        if (any) {
            for (size_t j = 0; j < mats.size(); j++) assign_block(cT, mats[j], N * currConstraint, N * idcs[j]);
            currConstraint++;
        }
    }
    Csr vtrans2cut = csr_from_triplets(N * cut.nV, N * nR, vT, true, 1e-15);
    Csr constraint = csr_from_triplets(N * currConstraint, N * nR, cT, true, 1e-15);
    // uncompress = blockdiag(lreductor), lreductor = [I; -I] (param.h:84-90, setup.ipp:141-143)
    std::vector<Trip> uT;
    for (int i = 0; i < nR; i++)
        for (int j = 0; j < n; j++) {
            uT.push_back({N * i + j, n * i + j, 1.0});
            uT.push_back({N * i + n + j, n * i + j, -1.0});
        }
    Csr uncompress = csr_from_triplets(N * nR, n * nR, uT);

    std::vector<int> fixedIdcs(n), integerIdcs((size_t)nT * n), singularIdcs;
    if (st.nS == 0) for (int j = 0; j < n; j++) fixedIdcs[j] = j;
    else {
        int iV = 0;
        for (; iV < raw.nV; iV++) if (is_singular(iV)) break;
        for (int j = 0; j < n; j++) fixedIdcs[j] = n * iV + j;
    }
    for (int i = 0; i < nT; i++) for (int j = 0; j < n; j++) integerIdcs[(size_t)n * i + j] = n * (raw.nV + i) + j;
    for (int v = 0; v < raw.nV; v++) if (is_singular(v)) for (int j = 0; j < n; j++) singularIdcs.push_back(n * v + j);
    std::vector<double> fixedVals(n, 0.0);

    // ---- integ (integ.ipp:4-43)
    st.fullx2Nfunc = csr_matmul(vtrans2cut, uncompress, true, 1e-15);
    Csr CfullAll = csr_matmul(constraint, uncompress, true, 1e-15);
    // reduce_to_linearly_independent on the dense restriction to the non-zero columns
    std::vector<int> ucols;
    {
        std::set<int> u(CfullAll.col.begin(), CfullAll.col.end());
        ucols.assign(u.begin(), u.end());
    }
    std::vector<int> cmask(n * nR, -1);
    for (size_t i = 0; i < ucols.size(); i++) cmask[ucols[i]] = (int)i;
    std::vector<std::vector<double>> dense(CfullAll.rows, std::vector<double>(ucols.size(), 0.0));
    for (int r = 0; r < CfullAll.rows; r++)
        for (int64_t k = CfullAll.ptr[r]; k < CfullAll.ptr[r + 1]; k++) dense[r][cmask[CfullAll.col[k]]] = CfullAll.val[k];
    std::vector<int> keep = independent_rows(dense);
    std::vector<Trip> cfT;
    for (size_t j = 0; j < keep.size(); j++)
        for (int64_t k = CfullAll.ptr[keep[j]]; k < CfullAll.ptr[keep[j] + 1]; k++)
            cfT.push_back({(int)j, CfullAll.col[k], CfullAll.val[k]});
    Csr Cfull = csr_from_triplets((int)keep.size(), n * nR, cfT);

    // G2 (integ.ipp:9-28) on the cut mesh; face frames of the cut mesh equal the raw ones (A.7)
    std::vector<Trip> gT;
    gT.reserve((size_t)raw.nF * 6 * N);
    for (int f = 0; f < raw.nF; f++) {
        const V3 nrm = raw.faceNormal[f], bx = raw.faceBasisX[f], by = raw.faceBasisY[f];
        const double a2 = raw.faceArea[f] * 2;
        for (int j = 0; j < 3; j++) {
            V3 nv = cut.pos[cut.idx[f][(j + 2) % 3]] - cut.pos[cut.idx[f][(j + 1) % 3]];  // vec(next(h))
            V3 g = nrm.cross(nv) / a2;
            int tail = cut.idx[f][j];
            for (int k = 0; k < N; k++) {
                int c = N * tail + k, ir = 2 * N * f + 2 * k;
                gT.push_back({ir + 0, c, g.dot(bx)});
                gT.push_back({ir + 1, c, g.dot(by)});
            }
        }
    }
    Csr G2 = csr_from_triplets(2 * N * raw.nF, N * cut.nV, gT);
    Csr G2p = csr_matmul(G2, st.fullx2Nfunc);  // the matrix handed to iterative_rounding (integ.ipp:72)

    // rawF2Vec (integ.ipp:30-43) and length (integ.ipp:70)
    std::vector<double> F2((size_t)2 * N * raw.nF);
    double normsum = 0;
    for (int f = 0; f < raw.nF; f++)
        for (int i = 0; i < N; i++) {
            V3 e(fs.ext[(size_t)f * 3 * N + 3 * i], fs.ext[(size_t)f * 3 * N + 3 * i + 1], fs.ext[(size_t)f * 3 * N + 3 * i + 2]);
            normsum += e.norm();
            F2[(size_t)2 * N * f + 2 * i + 0] = e.dot(raw.faceBasisX[f]);
            F2[(size_t)2 * N * f + 2 * i + 1] = e.dot(raw.faceBasisY[f]);
        }
    double norm = normsum / ((double)N * raw.nF);
    V3 lo = cut.pos[0], hi = cut.pos[0];
    for (const V3& p : cut.pos) {
        lo = V3(std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z));
        hi = V3(std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z));
    }
    st.length = (hi - lo).norm() * gridscale / norm;

    // ---- UFull (ir_init.h:98-154)
    std::vector<int> uniqueJ;
    {
        std::set<int> u(Cfull.col.begin(), Cfull.col.end());
        uniqueJ.assign(u.begin(), u.end());
    }
    std::vector<int> JMask(n * nR, -1);
    for (size_t i = 0; i < uniqueJ.size(); i++) JMask[uniqueJ[i]] = (int)i;
    std::vector<int> nonPart;
    for (int i = 0; i < n * nR; i++) if (JMask[i] < 0) nonPart.push_back(i);
    std::vector<std::vector<double>> USmall;
    int ker = 0;
    if (Cfull.rows != 0) {
        std::vector<std::vector<double>> CSmall(Cfull.rows, std::vector<double>(uniqueJ.size(), 0.0));
        for (int r = 0; r < Cfull.rows; r++)
            for (int64_t k = Cfull.ptr[r]; k < Cfull.ptr[r + 1]; k++) CSmall[r][JMask[Cfull.col[k]]] = Cfull.val[k];
        USmall = fullpiv_kernel(CSmall, Cfull.rows, (int)uniqueJ.size());
        ker = USmall.empty() ? 0 : (int)USmall[0].size();
    }
    std::vector<Trip> ufT;
    for (size_t ci = 0; ci < nonPart.size(); ci++) ufT.push_back({nonPart[ci], (int)ci, 1.0});
    for (size_t i = 0; i < uniqueJ.size(); i++)
        for (int j = 0; j < ker; j++)
            if (USmall[i][j] != 0) ufT.push_back({uniqueJ[i], (int)nonPart.size() + j, USmall[i][j]});
    st.UFull = csr_from_triplets(n * nR, (int)nonPart.size() + ker, ufT);
    Csr X2F = csr_matmul(G2p, st.UFull, true, 0.0);  // (G2 * UFull).pruned(), without `length`

    put_csr(bag, "fullx2Nfunc", st.fullx2Nfunc);
    put_csr(bag, "Cfull", Cfull);
    put_csr(bag, "G2p", G2p);
    put_csr(bag, "UFull", st.UFull);
    put_csr(bag, "X2F", X2F);
    bag.put("F2", F2);
    bag.put("fixedIdcs", fixedIdcs);
    bag.put("fixedVals", fixedVals);
    bag.put("integerIdcs", integerIdcs);
    bag.put("singularIdcs", singularIdcs);
    bag.put_scalar("length", st.length);
    bag.put("cutF", cut.idx);
    bag.put("cut2raw", cut.cut2raw);
    bag.put("he2matching", st.he2matching);
    bag.put("he2transidx", st.he2transidx);
    std::vector<int> dims = {N, n, raw.nF, nR, nT, st.nS, cut.nV, st.UFull.cols, currConstraint, Cfull.rows, ker};
    bag.put("igm_dims", dims);
}

// integ.ipp:88-98 result scatter plus the reference's runtime self-checks used as test invariants
// (SURVEY.md §4): uv-triangle orientation (example_qex.cpp:78), seam transition integrality.
inline void run_igm_finish(const Hmesh& raw, const FieldStage& fs, Bag& bag, const IgmStage& st) {
    const int N = st.N;
    std::vector<double> fullx = bag.get_f64("fullx");
    if ((int)fullx.size() != st.fullx2Nfunc.cols) throw std::runtime_error("igm_finish: fullx has wrong size");
    std::vector<double> NF = csr_matvec(st.fullx2Nfunc, fullx);  // nfn: cut.nV x N
    std::vector<double> cfn((size_t)raw.nF * 3 * N);
    for (int f = 0; f < raw.nF; f++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < N; k++) cfn[(size_t)f * 3 * N + N * j + k] = NF[(size_t)N * st.cut.idx[f][j] + k];
    bag.put("nfn", NF);
    bag.put("cfn", cfn);
    int flipped = 0;
    for (int f = 0; f < raw.nF; f++) {
        const double* a = &NF[(size_t)N * st.cut.idx[f][0]];
        const double* b = &NF[(size_t)N * st.cut.idx[f][1]];
        const double* c = &NF[(size_t)N * st.cut.idx[f][2]];
        double d1x = b[0] - a[0], d1y = b[1] - a[1], d2x = c[0] - a[0], d2y = c[1] - a[1];
        if (d1x * d2y - d1y * d2x <= 0) flipped++;
    }
    bag.put_scalar_i("flipped_faces", flipped);
    // seam check: across every seam edge the two copies of the (u,v) edge vector differ by a rotation of
    // k*90 degrees and the endpoint coordinates by R*uv + t with integer t when fully rounded
    double max_rot_dev = 0, max_trans_frac = 0;
    for (int e = 0; e < raw.nE; e++) {
        if (!fs.seam[e]) continue;
        int h = 2 * e, t = 2 * e + 1;
        auto uv_at = [&](int half, bool at_tail) {
            int f = raw.face[half];
            int v = at_tail ? raw.tail[half] : raw.head[half];
            for (int j = 0; j < 3; j++)
                if (raw.idx[f][j] == v) return std::array<double, 2>{NF[(size_t)N * st.cut.idx[f][j]], NF[(size_t)N * st.cut.idx[f][j] + 1]};
            return std::array<double, 2>{0, 0};
        };
        auto a0 = uv_at(h, true), a1 = uv_at(h, false);  // in face(h)
        auto b0 = uv_at(t, false), b1 = uv_at(t, true);  // same vertices in face(t)
        cplx da(a1[0] - a0[0], a1[1] - a0[1]), db(b1[0] - b0[0], b1[1] - b0[1]);
        double best = 1e300;
        int bk = 0;
        for (int k = 0; k < 4; k++) {
            double d = std::abs(db - da * std::polar(1.0, k * M_PI / 2));
            if (d < best) { best = d; bk = k; }
        }
        max_rot_dev = std::max(max_rot_dev, best);
        cplx tr = cplx(b0[0], b0[1]) - cplx(a0[0], a0[1]) * std::polar(1.0, bk * M_PI / 2);
        max_trans_frac = std::max(max_trans_frac, std::fabs(tr.real() - std::round(tr.real())));
        max_trans_frac = std::max(max_trans_frac, std::fabs(tr.imag() - std::round(tr.imag())));
    }
    bag.put_scalar("seam_rot_dev", max_rot_dev);
    bag.put_scalar("seam_trans_frac", max_trans_frac);
}

}  // namespace hg
