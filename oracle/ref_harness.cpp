// ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/): a C entry for each piece of the REFERENCE'S OWN hot-path code that can be compiled in this
// image. The reference headers are included where they lie (-I /root/reference/include; nothing is copied); Eigen and libigl, which the image
// does not have, are the stand-ins of oracle/ref_shim/ (read eigen_standin.hpp for what that does and does not pin). oracle/Makefile builds
// this file into oracle/_ref/libmetriko_ref.so; only tests/ and tests/golden/make_golden_ref.py load it.
//
//   ref_iterative_rounding   igm/iterative_rounding/iter_rounding.h:33-113 (NaiveIntegration ir_init.h:74-232, SeamlessIntegration
//                            ir_loop.h:9-125, Integration ir_common.h:12-94, LMSolver / DiagonalDamping lm.h:13-231, InjectiveBarrier
//                            injective_barrier.h:9-109, sparse_block mops.h:32-57)
//   ref_naive_integration    the NaiveIntegration constructor alone: UFull (ir_init.h:98-154), x0 (ir_init.h:173-212), X2F
//   ref_barrier_update       InjectiveBarrier::update + gen_image_filed (injective_barrier.h:37-109)
//   ref_solve_smallest_eig   solver/sparse_linear_solver.h:29-54 (the no-SuiteSparse branch) on given L, M with an explicit u0
//   ref_field_create / get   Hmesh(V, F) (hmesh/hmesh.h:313-569; igl::edge_topology is the stand-in of ref_shim/igl/edge_topology.h) and
//                            FaceRosyField(mesh, N, FieldType::Smoothest) (vectorfield/face_rosy_field.h:17-44: connectionLaplacian frf.h:46-76,
//                            galerkinMassMatrix frf.h:78-84, solveSmallestEig, decode frf.h:37-43), computeMatching (frf.h:100-202),
//                            compute_combbed_field (frf.h:207-249, seam given), compute_extrinsic_field (frf.h:251-264)
//   ref_field_parameterize   the rest of example/example_igm.cpp:17-33 on that field: compute_cut_mesh (param.h:126-152), RosyParameterization's
//                            constructor (compute_he2transidx / compute_he2matching, setup.ipp:3-50), setup() (setup.ipp:52-170), integ()
//                            (integ.ipp:3-121: G2, rawF2Vec, locally_injectivity_check, iterative_rounding, nfn / cfn). compute_seam (frf.h:268-360)
//                            is libigl's cut graph + Dijkstra and is NOT compiled: the seam is an input
#include <cstring>

#include "metriko/core/igm/iterative_rounding/iter_rounding.h"
#include "metriko/core/solver/sparse_linear_solver.h"
#include "metriko/core/vectorfield/face_rosy_field.h"
#include "metriko/core/igm/parameterization.h"

using namespace metriko;

namespace {
SprsD from_coo(int rows, int cols, long nnz, const int* r, const int* c, const double* v) {
    std::vector<TripD> t;
    t.reserve(nnz);
    for (long k = 0; k < nnz; k++) t.emplace_back(r[k], c[k], v[k]);
    SprsD m(rows, cols);
    m.setFromTriplets(t.begin(), t.end());
    return m;
}
SprsC from_coo_c(int rows, int cols, long nnz, const int* r, const int* c, const double* v) {
    std::vector<TripC> t;
    t.reserve(nnz);
    for (long k = 0; k < nnz; k++) t.emplace_back(r[k], c[k], complex(v[2 * k], v[2 * k + 1]));
    SprsC m(rows, cols);
    m.setFromTriplets(t.begin(), t.end());
    return m;
}
VecXi ivec(int n, const int* p) { VecXi v(n); for (int i = 0; i < n; i++) v(i) = p[i]; return v; }
VecXd dvec(long n, const double* p) { VecXd v(n); for (long i = 0; i < n; i++) v(i) = p[i]; return v; }
// copy a sparse matrix out as COO (column-major order); returns nnz; arrays may be null to query
long to_coo(const SprsD& m, long cap, int* r, int* c, double* v) {
    long k = 0;
    for (long j = 0; j < m.outerSize(); j++)
        for (SprsD::InnerIterator it(m, j); it; ++it) {
            if (r && k < cap) { r[k] = it.row(); c[k] = it.col(); v[k] = it.value(); }
            k++;
        }
    return k;
}
}  // namespace

extern "C" {

void ref_set_solver_hooks(int (*factor)(int, int, int, const int*, const int*, const double*), int (*solve)(int, const double*, double*),
                          void (*release)(int)) {
    refshim::hooks.factor = factor;
    refshim::hooks.solve = solve;
    refshim::hooks.release = release;
}

long ref_factor_count(int reset) {
    const long c = refshim::factor_count;
    if (reset) refshim::factor_count = 0;
    return c;
}

// iterative_rounding(...) exactly as RosyParameterization::integ calls it (igm/parameterization_integ.ipp:65-83). Cfull and G2 (= G2 *
// fullx2Nfunc, integ.ipp:72) arrive as COO. iter_cb may be null (= never stop early). Returns the function's bool (1 / 0); fullx_len receives
// fullx.size(), fullx (capacity fullx_cap) its coefficients.
int ref_iterative_rounding(int n_fixed, const int* fixedIdcs, const double* fixedVals, int n_sing, const int* singularIdcs, int n_int,
                           const int* integerIdcs, double length, int c_rows, int c_cols, long c_nnz, const int* c_r, const int* c_c,
                           const double* c_v, int g_rows, int g_cols, long g_nnz, const int* g_r, const int* g_c, const double* g_v, int N, int n,
                           int nF, int seamless, int roundSeams, int localInjectivity, int verbose, const double* F2,
                           int (*iter_cb)(const double* x, long len), long fullx_cap, double* fullx, long* fullx_len) {
    const VecXi fi = ivec(n_fixed, fixedIdcs), si = ivec(n_sing, singularIdcs), ii = ivec(n_int, integerIdcs);
    const VecXd fv = dvec(n_fixed, fixedVals), f2 = dvec((long)2 * N * nF, F2);
    const SprsD Cfull = from_coo(c_rows, c_cols, c_nnz, c_r, c_c, c_v), G2 = from_coo(g_rows, g_cols, g_nnz, g_r, g_c, g_v);
    std::function<bool(const VecXd&)> cb = [iter_cb](const VecXd& x) { return iter_cb ? iter_cb(x.data(), x.size()) != 0 : false; };
    VecXd out;
    const bool ok = iterative_rounding(fi, fv, si, ii, length, Cfull, G2, N, n, nF, seamless != 0, roundSeams != 0, localInjectivity != 0,
                                       verbose != 0, f2, cb, out);
    *fullx_len = out.size();
    for (long i = 0; i < out.size() && i < fullx_cap; i++) fullx[i] = out(i);
    return ok ? 1 : 0;
}

// NaiveIntegration's constructor (ir_init.h:74-232) alone. Outputs: UFull and X2F as COO (pass cap 0 / nulls to query the counts), x0
// (UFull.rows()), XSmall (UFull.cols()). dims = {UFull.rows, UFull.cols, nnz(UFull), X2F.rows, nnz(X2F)}.
int ref_naive_integration(int n_fixed, const int* fixedIdcs, const double* fixedVals, double length, int c_rows, int c_cols, long c_nnz,
                          const int* c_r, const int* c_c, const double* c_v, int g_rows, int g_cols, long g_nnz, const int* g_r, const int* g_c,
                          const double* g_v, int N, int n, int nF, const double* F2, long* dims, long u_cap, int* u_r, int* u_c, double* u_v,
                          long x_cap, int* x_r, int* x_c, double* x_v, double* x0, double* xsmall) {
    const VecXi fi = ivec(n_fixed, fixedIdcs);
    const VecXd fv = dvec(n_fixed, fixedVals), f2 = dvec((long)2 * N * nF, F2);
    const SprsD Cfull = from_coo(c_rows, c_cols, c_nnz, c_r, c_c, c_v), G2 = from_coo(g_rows, g_cols, g_nnz, g_r, g_c, g_v);
    InjectiveBarrier barrier(N, nF, f2);
    NaiveIntegration ni(G2, Cfull, f2, fi, fv, length, false, nF, N, n, [](const VecXd&) { return false; }, &barrier);
    dims[0] = ni.UFull.rows();
    dims[1] = ni.UFull.cols();
    dims[2] = to_coo(ni.UFull, u_cap, u_r, u_c, u_v);
    dims[3] = ni.X2F.rows();
    dims[4] = to_coo(ni.X2F, x_cap, x_r, x_c, x_v);
    if (x0) for (long i = 0; i < ni.x0.size(); i++) x0[i] = ni.x0(i);
    if (xsmall) for (long i = 0; i < ni.UFull.cols(); i++) xsmall[i] = ni.XF_Small(i);
    return ni.x0.size() == ni.UFull.rows() ? 1 : 0;  // 0: "Solver failed..." (ir_init.h:203-206) left the members unset
}

// InjectiveBarrier(N, nF, F2).update(field, updateJ, fBarrier, gBarrier) (injective_barrier.h:54-109): fBarrier[N nF], the diagonal of gBarrier
// [N nF] and the 4 entries per row of gen_image_filed(field) (values S, columns J; injective_barrier.h:27-33, 44-52, 85-89).
int ref_barrier_update(int N, int nF, const double* F2, const double* field, int updateJ, double* fBarrier, double* gDiag, int* imgCols,
                       double* imgVals) {
    const VecXd f2 = dvec((long)2 * N * nF, F2), fld = dvec((long)2 * N * nF, field);
    InjectiveBarrier b(N, nF, f2);
    VecXd fB;
    SprsD gB;
    b.update(fld, updateJ != 0, fB, gB);
    for (long i = 0; i < fB.size(); i++) fBarrier[i] = fB(i);
    if (updateJ) {
        for (long i = 0; i < (long)N * nF; i++) gDiag[i] = 0.0;
        for (long j = 0; j < gB.outerSize(); j++)
            for (SprsD::InnerIterator it(gB, j); it; ++it)
                if (it.row() == it.col()) gDiag[it.row()] = it.value();
        // gen_image_filed sums duplicate (row, col) pairs; none occur for N >= 3, and the dense copy below follows the triplet order
        const SprsD img = b.gen_image_filed(fld);
        std::vector<int> fill((size_t)N * nF, 0);
        const SprsD imgT = img.transpose();  // column j of the transpose = row j of img, columns ascending
        for (long j = 0; j < imgT.outerSize(); j++)
            for (SprsD::InnerIterator it(imgT, j); it; ++it) {
                const int k = fill[j]++;
                if (k < 4) { imgCols[4 * j + k] = it.row(); imgVals[4 * j + k] = it.value(); }
            }
    }
    return 1;
}

// solveSmallestEig(SprsC& L, SprsC& M, nIter) (sls.h:29-54, Eigen::SparseLU branch). VecXc::Random is fed from u0 (2 n doubles).
int ref_solve_smallest_eig(int n, long l_nnz, const int* l_r, const int* l_c, const double* l_v, long m_nnz, const int* m_r, const int* m_c,
                           const double* m_v, const double* u0, int nIter, double* x) {
    SprsC L = from_coo_c(n, n, l_nnz, l_r, l_c, l_v), M = from_coo_c(n, n, m_nnz, m_r, m_c, m_v);
    refshim::random_source = u0;
    refshim::random_pos = 0;
    const VecXc r = solveSmallestEig(L, M, nIter);
    refshim::random_source = nullptr;
    for (long i = 0; i < r.size(); i++) { x[2 * i] = r(i).real(); x[2 * i + 1] = r(i).imag(); }
    return 1;
}


struct RefField {
    std::unique_ptr<Hmesh> mesh;
    std::unique_ptr<FaceRosyField> field, combed;
    SprsC L, M;
    MatXd ext, cext;
    std::vector<bool> seam;
    std::unique_ptr<Hmesh> cut;
    std::unique_ptr<RosyParameterization> rp;
    int integ_ok = -1;
};

// V[3 nV] and F[3 nF] row-major. u0: 2 nF doubles for VecXc::Random. matching_type 0 = Principal, 1 = Curl. seam[nE] (0 / 1) or null: with a
// seam the combed field (compute_combbed_field) is computed as well.
// field_type 1 = FieldType::Smoothest, 2 = FieldType::CurvatureAligned (frf.h:25-33: principalCurvatureDir frf.h:86-98, solveSquare sls.h:16-27).
void* ref_field_create2(int nV, const double* V, int nF, const int* F, int N, const double* u0, int matching_type, const int* seam, int field_type);
void* ref_field_create(int nV, const double* V, int nF, const int* F, int N, const double* u0, int matching_type, const int* seam) {
    return ref_field_create2(nV, V, nF, F, N, u0, matching_type, seam, 1);
}
void* ref_field_create2(int nV, const double* V, int nF, const int* F, int N, const double* u0, int matching_type, const int* seam, int field_type) {
    MatXd Vm(nV, 3);
    MatXi Fm(nF, 3);
    for (int i = 0; i < nV; i++) for (int j = 0; j < 3; j++) Vm(i, j) = V[3 * i + j];
    for (int i = 0; i < nF; i++) for (int j = 0; j < 3; j++) Fm(i, j) = F[3 * i + j];
    auto* r = new RefField;
    r->mesh = std::make_unique<Hmesh>(Vm, Fm);
    refshim::random_source = u0;
    refshim::random_pos = 0;
    r->field = std::make_unique<FaceRosyField>(*r->mesh, N, field_type == 2 ? FieldType::CurvatureAligned : FieldType::Smoothest);
    refshim::random_source = nullptr;
    r->L = r->field->connectionLaplacian();
    r->M = r->field->galerkinMassMatrix();
    r->field->computeMatching(matching_type ? MatchingType::Curl : MatchingType::Principal);
    if (N == 4) r->ext = compute_extrinsic_field(*r->mesh, *r->field, N);
    if (seam) {
        std::vector<bool> sm(r->mesh->nE);
        for (int e = 0; e < r->mesh->nE; e++) sm[e] = seam[e] != 0;
        r->combed = compute_combbed_field(*r->field, sm);
        r->seam = sm;
    }
    return r;
}

// example_igm.cpp:17-33 from `auto cutm = compute_cut_mesh(*mesh, seam)` on. Needs a handle created with a seam. Returns integ()'s bool.
int ref_field_parameterize(void* h, double gridscale, int seamless, int localInjectivity, int roundSeams) {
    RefField& r = *static_cast<RefField*>(h);
    if (!r.combed) return -1;
    r.cut = compute_cut_mesh(*r.mesh, r.seam);
    r.cext = compute_extrinsic_field(*r.mesh, *r.combed, r.field->rosyN);
    r.rp = std::make_unique<RosyParameterization>(*r.mesh, *r.cut, r.cext, r.combed->singular, r.combed->matching, r.seam, r.field->rosyN, gridscale);
    r.rp->seamless = seamless != 0;
    r.rp->localInjectivity = localInjectivity != 0;
    r.rp->roundSeams = roundSeams != 0;
    r.rp->verbose = false;
    r.rp->setup();
    r.integ_ok = r.rp->integ() ? 1 : 0;
    return r.integ_ok;
}

void ref_field_destroy(void* h) { delete static_cast<RefField*>(h); }

// every array as doubles (integers converted; complex interleaved; matrices row-major); returns the length, fills up to cap
long ref_field_get(void* h, const char* name, long cap, double* out) {
    const RefField& r = *static_cast<RefField*>(h);
    const Hmesh& m = *r.mesh;
    std::vector<double> v;
    const std::string k(name);
    auto ints = [&](const std::vector<int>& a) { v.assign(a.begin(), a.end()); };
    auto vecd = [&](const VecXd& a) { for (long i = 0; i < a.size(); i++) v.push_back(a(i)); };
    auto veci = [&](const VecXi& a) { for (long i = 0; i < a.size(); i++) v.push_back(a(i)); };
    auto vecc = [&](const VecXc& a) { for (long i = 0; i < a.size(); i++) { v.push_back(a(i).real()); v.push_back(a(i).imag()); } };
    auto matd = [&](const MatXd& a) { for (long i = 0; i < a.rows(); i++) for (long j = 0; j < a.cols(); j++) v.push_back(a(i, j)); };
    auto mati = [&](const MatXi& a) { for (long i = 0; i < a.rows(); i++) for (long j = 0; j < a.cols(); j++) v.push_back(a(i, j)); };
    auto matc = [&](const MatXc& a) { for (long i = 0; i < a.rows(); i++) for (long j = 0; j < a.cols(); j++) { v.push_back(a(i, j).real()); v.push_back(a(i, j).imag()); } };
    auto coo = [&](const SprsC& a, int what) {
        for (long j = 0; j < a.outerSize(); j++)
            for (SprsC::InnerIterator it(a, j); it; ++it) {
                if (what == 0) v.push_back(it.row());
                else if (what == 1) v.push_back(it.col());
                else { v.push_back(it.value().real()); v.push_back(it.value().imag()); }
            }
    };
    if (k == "dims") v = {(double)m.nV, (double)m.nF, (double)m.nE, (double)m.nH, (double)m.nEularChars};
    else if (k == "twin") ints(m.twin);
    else if (k == "next") ints(m.next);
    else if (k == "prev") ints(m.prev);
    else if (k == "face") ints(m.face);
    else if (k == "tail") ints(m.tail);
    else if (k == "head") ints(m.head);
    else if (k == "edge") ints(m.edge);
    else if (k == "crnr") ints(m.crnr);
    else if (k == "face2half") ints(m.face2half);
    else if (k == "vert2half") ints(m.vert2half);
    else if (k == "heArgOnFace") vecd(m.heArgOnFace);
    else if (k == "faceArea") vecd(m.faceArea);
    else if (k == "angleDefect") vecd(m.angleDefect);
    else if (k == "dihedralArg") vecd(m.dihedralArg);
    else if (k == "heLen") { for (const Half& hh : m.halfs) v.push_back(hh.len()); }
    else if (k == "faceBasisX") matd(m.faceBasisX);
    else if (k == "faceBasisY") matd(m.faceBasisY);
    else if (k == "faceNormal") matd(m.faceNormal);
    else if (k == "edge2face") mati(m.edge2face);
    else if (k == "edge2vert") mati(m.edge2vert);
    else if (k == "face2edge") mati(m.face2edge);
    else if (k == "L_rows") coo(r.L, 0);
    else if (k == "L_cols") coo(r.L, 1);
    else if (k == "L_vals") coo(r.L, 2);
    else if (k == "M_rows") coo(r.M, 0);
    else if (k == "M_vals") coo(r.M, 2);
    else if (k == "compressed") vecc(r.field->compressed);
    else if (k == "principalCurvatureDir") vecc(r.field->principalCurvatureDir());
    else if (k == "field") matc(r.field->field);
    else if (k == "connection") vecc(r.field->connection);
    else if (k == "matching") veci(r.field->matching);
    else if (k == "singular") veci(r.field->singular);
    else if (k == "ext") matd(r.ext);
    else if (k == "cfield" && r.combed) matc(r.combed->field);
    else if (k == "cmatching" && r.combed) veci(r.combed->matching);
    else if (k == "cext" && r.rp) matd(r.cext);
    else if (k == "cutF" && r.cut) mati(r.cut->idx);
    else if (k == "cutPos" && r.cut) matd(r.cut->pos);
    else if (k == "fullx" && r.rp) vecd(r.rp->fullx);
    else if (k == "nfn" && r.rp) matd(r.rp->nfn);
    else if (k == "cfn" && r.rp) matd(r.rp->cfn);
    else if (k == "fixedIdcs" && r.rp) veci(r.rp->fixedIdcs);
    else if (k == "fixedVals" && r.rp) vecd(r.rp->fixedVals);
    else if (k == "integerIdcs" && r.rp) veci(r.rp->integerIdcs);
    else if (k == "singularIdcs" && r.rp) veci(r.rp->singularIdcs);
    else if (k == "he2matching" && r.rp) veci(r.rp->he2matching);
    else if (k == "he2transidx" && r.rp) veci(r.rp->he2transidx);
    else if (r.rp && (k.rfind("vtrans2cut_", 0) == 0 || k.rfind("uncompress_", 0) == 0 || k.rfind("constraint_", 0) == 0)) {
        const SprsD& a = k[0] == 'v' ? r.rp->vtrans2cut : (k[0] == 'u' ? r.rp->uncompress : r.rp->constraint);
        const std::string what = k.substr(k.find('_') + 1);
        if (what == "shape") v = {(double)a.rows(), (double)a.cols()};
        else for (long j = 0; j < a.outerSize(); j++)
            for (SprsD::InnerIterator it(a, j); it; ++it) v.push_back(what == "rows" ? it.row() : what == "cols" ? it.col() : it.value());
    }
    else return -1;
    for (long i = 0; i < (long)v.size() && i < cap; i++) out[i] = v[i];
    return (long)v.size();
}

}  // extern "C"
