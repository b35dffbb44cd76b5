// Compiles include/metriko_b200/metriko_cuda_shims.hpp (seams B1, B2) against a minimal stand-in for the handful of Eigen
// accessors it uses (Eigen itself is not in this image) and links it with libmetriko_b200.so. Without a GPU the context
// constructor must throw -- that is the "fails loudly, no CPU fallback" contract; with a GPU (argv[1] == "run") a 4-face
// problem goes through both seams.
#include <complex>
#include <cstdio>
#include <cstring>
#include <vector>

namespace metriko {
template <class T>
struct MockVec {
    std::vector<T> v;
    MockVec() = default;
    explicit MockVec(int n) : v(n) {}
    T* data() { return v.data(); }
    const T* data() const { return v.data(); }
    long size() const { return (long)v.size(); }
    void resize(long n) { v.resize(n); }
    T& operator()(int i) { return v[i]; }
};
template <class T>
struct MockSparse {  // compressed column storage, as Eigen::SparseMatrix<T> holds it
    int n = 0;
    std::vector<int> outer, inner;
    std::vector<T> val;
    long rows() const { return n; }
    void makeCompressed() {}
    const int* outerIndexPtr() const { return outer.data(); }
    const int* innerIndexPtr() const { return inner.data(); }
    const T* valuePtr() const { return val.data(); }
};
using VecXc = MockVec<std::complex<double>>;
using VecXd = MockVec<double>;
using SprsC = MockSparse<std::complex<double>>;
using SprsD = MockSparse<double>;
}  // namespace metriko

#define METRIKO_B200_SHIM_TYPES_PROVIDED
#define METRIKO_B200_SHIM_NO_B4
#include "metriko_b200/metriko_cuda_shims.hpp"

int main(int argc, char** argv) {
    using namespace metriko;
    const bool run = argc > 1 && !std::strcmp(argv[1], "run");
    try {
        cuda::Context ctx(0);
        if (!run) { std::puts("context created"); return 0; }
        // tetrahedron-like 4x4 Hermitian matrix: 3 I - (ones - I) * phase-free couplings, M = I
        SprsC L, M;
        L.n = M.n = 4;
        for (int c = 0; c < 4; c++) {
            L.outer.push_back((int)L.inner.size());
            M.outer.push_back((int)M.inner.size());
            for (int r = 0; r < 4; r++) { L.inner.push_back(r); L.val.push_back(r == c ? 3.0 + 1e-9 : -0.9); }
            M.inner.push_back(c); M.val.push_back(1.0);
        }
        L.outer.push_back((int)L.inner.size());
        M.outer.push_back((int)M.inner.size());
        VecXc u0(4);
        for (int i = 0; i < 4; i++) u0(i) = {0.3 + 0.1 * i, -0.2 * i};
        VecXc x = cuda::solveSmallestEig(L, M, 50, u0, nullptr, ctx);
        double nrm = 0;
        for (int i = 0; i < 4; i++) nrm += std::norm(x(i));
        std::printf("B1 ok |x|^2 = %.12f\n", nrm);
        SprsD A;
        A.n = 3;
        A.outer = {0, 2, 5, 7};
        A.inner = {0, 1, 0, 1, 2, 1, 2};
        A.val = {4, -1, -1, 4, -1, -1, 4};
        cuda::CholeskyWrapper w(ctx);
        VecXd b(3), y;
        b(0) = 1; b(1) = 2; b(2) = 3;
        if (!w.factorize(A) || !w.solve(b, y)) { std::puts("B2 failed"); return 2; }
        std::printf("B2 ok x = %.9f %.9f %.9f\n", y(0), y(1), y(2));
        return (nrm > 0.999999 && nrm < 1.000001) ? 0 : 3;
    } catch (const std::runtime_error& e) {
        std::printf("runtime_error: %s\n", e.what());
        return run ? 4 : 0;  // on a CPU-only box this is the expected outcome
    }
}
