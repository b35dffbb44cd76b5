// Seam B4 of include/metriko_b200/metriko_cuda_shims.hpp through a compiler and, with a GPU, through the library: detail::kernel_basis
// (UFull from Cfull, ir_init.h:98-154), X2F = (G2 UFull).pruned() (ir_init.h:173), detail::CsrCopy and cuda::iterative_rounding with the
// reference's own argument list (ir.h:33-51). Eigen is not in this image: tests/cpp/mini_eigen.hpp supplies the accessors the overlay uses.
//
//   shim_b4_test host <in.bin> <out.bin>   UFull and X2F as the overlay builds them (CPU only; compared with the host glue by the test)
//   shim_b4_test run  <in.bin> <out.bin>   + cuda::iterative_rounding on device 0: success flag and fullx
//
// in.bin : int32 N, n, nF, seamless, roundSeams, localInjectivity; f64 length; Cfull, G2 as (int32 rows, cols; int64 nnz; int32 r[], c[]; f64 v[]);
//          then int64-counted arrays fixedIdcs (i32), fixedVals (f64), singularIdcs (i32), integerIdcs (i32), intrinsicField (f64)
// out.bin: UFull, X2F in the same matrix layout; run mode appends int32 success and the int64-counted f64 array fullx
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "mini_eigen.hpp"
#define METRIKO_B200_SHIM_TYPES_PROVIDED
#include "metriko_b200/metriko_cuda_shims.hpp"

using namespace metriko;

static void rd(FILE* f, void* p, size_t bytes) {
    if (bytes && std::fread(p, 1, bytes, f) != bytes) throw std::runtime_error("short read");
}
static SprsD read_matrix(FILE* f) {
    int32_t rc[2];
    int64_t nnz;
    rd(f, rc, sizeof(rc));
    rd(f, &nnz, sizeof(nnz));
    std::vector<int32_t> r(nnz), c(nnz);
    std::vector<double> v(nnz);
    rd(f, r.data(), nnz * 4);
    rd(f, c.data(), nnz * 4);
    rd(f, v.data(), nnz * 8);
    std::vector<Eigen::Triplet<double>> T;
    T.reserve(nnz);
    for (int64_t k = 0; k < nnz; k++) T.emplace_back(r[k], c[k], v[k]);
    SprsD A(rc[0], rc[1]);
    A.setFromTriplets(T.begin(), T.end());
    return A;
}
template <class V, class T>
static V read_vec(FILE* f) {
    int64_t n;
    rd(f, &n, sizeof(n));
    V v(n);
    rd(f, v.data(), n * sizeof(T));
    return v;
}
static void write_matrix(FILE* f, const SprsD& A) {
    const int32_t rc[2] = {(int32_t)A.rows(), (int32_t)A.cols()};
    const int64_t nnz = A.nonZeros();
    std::vector<int32_t> r, c;
    std::vector<double> v;
    for (int k = 0; k < A.outerSize(); ++k)
        for (SprsD::InnerIterator it(A, k); it; ++it) { r.push_back(it.row()); c.push_back(it.col()); v.push_back(it.value()); }
    std::fwrite(rc, sizeof(rc), 1, f);
    std::fwrite(&nnz, sizeof(nnz), 1, f);
    std::fwrite(r.data(), 4, nnz, f);
    std::fwrite(c.data(), 4, nnz, f);
    std::fwrite(v.data(), 8, nnz, f);
}

int main(int argc, char** argv) {
    if (argc < 4) { std::puts("usage: shim_b4_test host|run in.bin out.bin"); return 64; }
    const bool run = !std::strcmp(argv[1], "run");
    try {
        FILE* f = std::fopen(argv[2], "rb");
        if (!f) throw std::runtime_error("cannot open input");
        int32_t hdr[6];
        double length;
        rd(f, hdr, sizeof(hdr));
        rd(f, &length, sizeof(length));
        const SprsD Cfull = read_matrix(f), G2 = read_matrix(f);
        const VecXi fixedIdcs = read_vec<VecXi, int32_t>(f);
        const VecXd fixedVals = read_vec<VecXd, double>(f);
        const VecXi singularIdcs = read_vec<VecXi, int32_t>(f), integerIdcs = read_vec<VecXi, int32_t>(f);
        const VecXd intrinsicField = read_vec<VecXd, double>(f);
        std::fclose(f);

        const SprsD UFull = cuda::detail::kernel_basis(Cfull);
        const SprsD X2F = (G2 * UFull).pruned();
        cuda::detail::CsrCopy Xc(X2F);  // the arrays mk_igm_create receives
        if (Xc.rows != X2F.rows() || (long)Xc.col.size() != X2F.nonZeros() || Xc.ptr.back() != (int64_t)X2F.nonZeros()) throw std::runtime_error("CsrCopy inconsistent");
        for (int r = 0; r < Xc.rows; r++)
            for (int64_t k = Xc.ptr[r] + 1; k < Xc.ptr[r + 1]; k++)
                if (Xc.col[k - 1] >= Xc.col[k]) throw std::runtime_error("CsrCopy rows not strictly sorted");
        FILE* o = std::fopen(argv[3], "wb");
        if (!o) throw std::runtime_error("cannot open output");
        write_matrix(o, UFull);
        write_matrix(o, X2F);
        std::printf("host ok: UFull %ld x %ld (%ld), X2F %ld x %ld (%ld)\n", UFull.rows(), UFull.cols(), UFull.nonZeros(), X2F.rows(), X2F.cols(), X2F.nonZeros());
        if (run) {
            cuda::Context ctx(0);
            VecXd fullx;
            const bool ok = cuda::iterative_rounding(fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, Cfull, G2, hdr[0], hdr[1], hdr[2], hdr[3] != 0,
                                                     hdr[4] != 0, hdr[5] != 0, false, intrinsicField, fullx, ctx);
            const int32_t okf = ok ? 1 : 0;
            const int64_t n = fullx.size();
            std::fwrite(&okf, 4, 1, o);
            std::fwrite(&n, 8, 1, o);
            std::fwrite(fullx.data(), 8, n, o);
            std::printf("B4 %s: fullx %ld entries\n", ok ? "ok" : "returned false", (long)n);
        }
        std::fclose(o);
        return 0;
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 4;
    }
}
