// ref_overlay_test.cpp -- the overlay include/metriko_b200/metriko_cuda_shims.hpp compiled against THE REFERENCE'S OWN HEADERS (included in place
// from /root/reference/include; Eigen / libigl are the stand-ins of oracle/ref_shim/), with the reference's own call sites redirected to it:
//
//   B1  `compressed = solveSmallestEig(L, M);`                      vectorfield/face_rosy_field.h:22   ->  metriko::cuda::solveSmallestEig
//   B4  `bool success = iterative_rounding(fixedIdcs, ..., fullx);`  igm/parameterization_integ.ipp:65   ->  metriko::cuda::iterative_rounding
//
// The reference cannot be edited here, so the two one-word edits INTEGRATION.md describes are made by the preprocessor: sls.h and iter_rounding.h
// are included FIRST (their definitions keep their names), then the names are #defined to the overlay's while face_rosy_field.h and
// parameterization.h -- the callers -- are read. What comes out is example/example_igm.cpp:14-33 with both solver seams on the B200.
// tests/test_shim.py builds this file into a shared library on the CPU (compile + link against libmetriko_b200.so + symbol check); running it needs
// a GPU: scripts/run_ref_overlay.py calls ref_overlay_example_igm on the reference's icosphere.obj (round 2, B200: integ() true, integers identical
// to the oracle golden, fullx within 7.2e-9 of the golden and of the reference's own CPU run; profiles/overlay_r02_reference_example_igm_on_b200.log).
#include "metriko/core/solver/sparse_linear_solver.h"
#include "metriko/core/igm/iterative_rounding/iter_rounding.h"

#include "metriko_b200/metriko_cuda_shims.hpp"

namespace metriko::cuda {
inline const double* g_u0 = nullptr;  // VecXc::Random stand-in of the B1 call site: the explicit start vector
inline VecXc start_vector(long n) {
    VecXc u(n);
    for (long i = 0; i < n; i++) u(i) = complex(g_u0[2 * i], g_u0[2 * i + 1]);
    return u;
}
}  // namespace metriko::cuda

#define solveSmallestEig(L, M) cuda::solveSmallestEig(L, M, 50, cuda::start_vector((L).rows()))
#define iterative_rounding cuda::iterative_rounding
#include "metriko/core/vectorfield/face_rosy_field.h"
#include "metriko/core/igm/parameterization.h"
#undef iterative_rounding
#undef solveSmallestEig

using namespace metriko;

extern "C" {

// example_igm.cpp:14-33 (compute_seam aside: libigl; the seam is an input) with B1 and B4 on the device. V[3 nV], F[3 nF] row-major, u0[2 nF],
// seam[nE]. Outputs: fullx (capacity fullx_cap, length returned through fullx_len), cfn[nF x 3N] row-major. Returns integ()'s bool, -1 on an
// exception (e.g. no sm_100a device: the overlay's Context throws).
int ref_overlay_example_igm(int nV, const double* V, int nF, const int* F, int N, const double* u0, const int* seam, double gridscale, int seamless,
                            int localInjectivity, int roundSeams, long fullx_cap, double* fullx, long* fullx_len, double* cfn) {
    try {
        MatXd Vm(nV, 3);
        MatXi Fm(nF, 3);
        for (int i = 0; i < nV; i++) for (int j = 0; j < 3; j++) Vm(i, j) = V[3 * i + j];
        for (int i = 0; i < nF; i++) for (int j = 0; j < 3; j++) Fm(i, j) = F[3 * i + j];
        cuda::g_u0 = u0;
        auto mesh = std::make_unique<Hmesh>(Vm, Fm);
        auto rawf = std::make_unique<FaceRosyField>(*mesh, N, FieldType::Smoothest);  // B1 inside
        rawf->computeMatching(MatchingType::Curl);
        std::vector<bool> sm(mesh->nE);
        for (int e = 0; e < mesh->nE; e++) sm[e] = seam[e] != 0;
        auto cutm = compute_cut_mesh(*mesh, sm);
        auto cmbf = compute_combbed_field(*rawf, sm);
        MatXd extf = compute_extrinsic_field(*mesh, *cmbf, N);
        RosyParameterization rp(*mesh, *cutm, extf, cmbf->singular, cmbf->matching, sm, N, gridscale);
        rp.seamless = seamless != 0;
        rp.localInjectivity = localInjectivity != 0;
        rp.roundSeams = roundSeams != 0;
        rp.verbose = false;
        rp.setup();
        const bool ok = rp.integ();  // B4 inside
        *fullx_len = rp.fullx.size();
        for (long i = 0; i < rp.fullx.size() && i < fullx_cap; i++) fullx[i] = rp.fullx(i);
        if (cfn) for (long i = 0; i < rp.cfn.rows(); i++) for (long j = 0; j < rp.cfn.cols(); j++) cfn[i * rp.cfn.cols() + j] = rp.cfn(i, j);
        return ok ? 1 : 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_overlay_example_igm: %s\n", e.what());
        return -1;
    }
}

}  // extern "C"
