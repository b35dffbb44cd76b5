// RosyParameterization::integ() on the device (igm/parameterization_integ.ipp:3-98): the per-face gradient stencil G2 (integ.ipp:9-28),
// the field in face coordinates rawF2Vec (integ.ipp:30-43), the operator X2F = (G2 * fullx2Nfunc * UFull).pruned() handed to the
// rounding loop (integ.ipp:72, iter_rounding_init.h:173) and the result scatter NF = fullx2Nfunc * fullx -> nfn / cfn (integ.ipp:88-98).
//
// In the reference these are triplet lists + three Eigen SpGEMMs on the host (50 M entries at 1M vertices); round 1 of this library
// still took X2F from the host (0.96 GB of H2D per step). Here the host sends what integ() itself reads -- the cut-mesh corner table
// and positions, fullx2Nfunc from setup(), UFull -- and every row of X2F is written by the thread that owns it:
//     T   = fullx2Nfunc * UFull                  (N cut.nV x cols; a signed near-permutation: device SpGEMM, 4 M entries)
//     row (f, k, c) of X2F = sum_{j<3} g_c(f, j) * T[N cutF(f, j) + k, :]      c = 0, 1 (basisX / basisY component)
// i.e. a 3-way merge of three short sorted rows -- the "precomputed, atomic-free scatter map" is the corner table itself.
#pragma once
#include "csr.cuh"
#include "devcsr.cuh"

namespace mk {

// Face frame exactly as Hmesh builds it (hmesh/hmesh.h:461-489): basisX = normalize(p1 - p0), normal = normalize((p1 - p0) x (p2 - p0)),
// basisY = normal x basisX, area = |(p1 - p0) x (p2 - p0)| / 2; gradient of the hat function of corner j (integ.ipp:13):
// g = normal x (p_{j+2} - p_{j+1}) / (2 area). Output: grad[(3 f + j) * 2 + c] = g . basis_c. One thread per face.
static __global__ void __launch_bounds__(kThreads) k_face_gradients(int nF, const int* __restrict__ cutF, const double* __restrict__ pos,
                                                             double* __restrict__ grad) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        double p[3][3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int v = cutF[3 * f + j];
            p[j][0] = pos[3 * (size_t)v]; p[j][1] = pos[3 * (size_t)v + 1]; p[j][2] = pos[3 * (size_t)v + 2];
        }
        const double e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
        const double e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
        const double l1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
        const double bx[3] = {e1[0] / l1, e1[1] / l1, e1[2] / l1};
        const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        const double nn[3] = {n[0] / ln, n[1] / ln, n[2] / ln};
        const double by[3] = {nn[1] * bx[2] - nn[2] * bx[1], nn[2] * bx[0] - nn[0] * bx[2], nn[0] * bx[1] - nn[1] * bx[0]};
        const double a2 = ln * 0.5 * 2.0;  // 2 * area
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int ja = (j + 1) % 3, jb = (j + 2) % 3;
            const double nv[3] = {p[jb][0] - p[ja][0], p[jb][1] - p[ja][1], p[jb][2] - p[ja][2]};
            const double g[3] = {(nn[1] * nv[2] - nn[2] * nv[1]) / a2, (nn[2] * nv[0] - nn[0] * nv[2]) / a2, (nn[0] * nv[1] - nn[1] * nv[0]) / a2};
            grad[((size_t)3 * f + j) * 2 + 0] = g[0] * bx[0] + g[1] * bx[1] + g[2] * bx[2];
            grad[((size_t)3 * f + j) * 2 + 1] = g[0] * by[0] + g[1] * by[1] + g[2] * by[2];
        }
    }
}

// One thread per row r = 2N f + 2k + c of X2F: merges rows N cutF(f, j) + k of T (sorted columns) scaled by grad(f, j, c).
// FILL = false: len[r] = number of entries that survive pruning (exact zeros are dropped like Eigen's .pruned(), ir_init.h:173);
// FILL = true : writes them at ptr[r]. Both passes run the same arithmetic, so they agree on what is pruned.
template <bool FILL>
static __global__ void __launch_bounds__(kThreads) k_x2f_rows(int N, int nF, const int* __restrict__ cutF, const double* __restrict__ grad,
                                                       const int* __restrict__ Tp, const int* __restrict__ Tc, const double* __restrict__ Tv,
                                                       const int* __restrict__ Xp, int* __restrict__ len_or_col, double* __restrict__ Xv) {
    const int64_t rows = (int64_t)2 * N * nF;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(r / (2 * N));
        const int rem = (int)(r - (int64_t)f * 2 * N);
        const int k = rem >> 1, c = rem & 1;
        int b[3], e[3];
        double w[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int tr = N * cutF[3 * f + j] + k;
            b[j] = Tp[tr];
            e[j] = Tp[tr + 1];
            w[j] = grad[((size_t)3 * f + j) * 2 + c];
        }
        int cnt = 0;
        int64_t o = FILL ? Xp[r] : 0;
        while (true) {
            int col = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (b[j] < e[j]) col = min(col, Tc[b[j]]);
            if (col == 0x7fffffff) break;
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < 3; j++)
                if (b[j] < e[j] && Tc[b[j]] == col) { v += w[j] * Tv[b[j]]; b[j]++; }
            if (v != 0.0) {
                if (FILL) { len_or_col[o + cnt] = col; Xv[o + cnt] = v; }
                cnt++;
            }
        }
        if (!FILL) len_or_col[r] = cnt;
    }
}

// rawF2Vec (integ.ipp:30-43) from the device-resident eigen-solve result and the per-face combing turn (frf.h:207-249): the combed
// field of face f is rotate_left(field_f, turn_f), field(f, i) = e^{i arg(z_f) / N} e^{2 pi i i / N} (frf.h:37-43), and its extrinsic
// vector is unit (frf.h:258-261), so its face coordinates are (cos, sin) of the angle.
static __global__ void __launch_bounds__(kThreads) k_f2_from_field(int nF, int N, const cdouble* __restrict__ z, const int* __restrict__ turns,
                                                            double* __restrict__ F2) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
        const cdouble zf = ldc_nc(z + f);
        const double a = atan2(zf.y, zf.x) / (double)N;
        const int t = turns[f];
        for (int i = 0; i < N; i++) {
            double s, c;
            sincos(a + 2.0 * M_PI * (double)((i + t) % N) / (double)N, &s, &c);
            F2[(size_t)2 * N * f + 2 * i] = c;
            F2[(size_t)2 * N * f + 2 * i + 1] = s;
        }
    }
}

// cfn(f, N j + k) = nfn(cutF(f, j), k)  (integ.ipp:93-98)
static __global__ void __launch_bounds__(kThreads) k_corner_gather(int nF, int N, const int* __restrict__ cutF, const double* __restrict__ nfn,
                                                            double* __restrict__ cfn) {
    const int64_t tot = (int64_t)nF * 3 * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / (3 * N));
        const int rem = (int)(i - (int64_t)f * 3 * N);
        const int j = rem / N, k = rem - j * N;
        cfn[i] = nfn[(size_t)N * cutF[3 * f + j] + k];
    }
}

// X2F = (G2 * T).pruned() with T = fullx2Nfunc * UFull already on the device; rows have strictly increasing columns by construction.
inline void build_x2f(int N, int nF, const int* cutF, const double* pos, const DeviceCsr<double>& T, DeviceCsr<double>& X, cudaStream_t s) {
    const int64_t rows64 = (int64_t)2 * N * nF;
    if (rows64 > 2000000000LL) throw std::invalid_argument("build_x2f: too many rows for int32 CSR");
    const int rows = (int)rows64;
    DevBuf<double> grad((size_t)6 * nF);
    k_face_gradients<<<stream_grid(nF), kThreads, 0, s>>>(nF, cutF, pos, grad.p);
    DevBuf<int> len((size_t)rows + 1);
    len.zero(s);
    k_x2f_rows<false><<<stream_grid(rows64), kThreads, 0, s>>>(N, nF, cutF, grad.p, T.ptr.p, T.col.p, T.val.p, nullptr, len.p, nullptr);
    MK_CHECK_LAUNCH();
    X.rows = rows;
    X.cols = T.cols;
    X.ptr.alloc((size_t)rows + 1);
    dev_exclusive_scan(len.p, X.ptr.p, rows, s);
    const int64_t nnz = dev_read_int(X.ptr.p + rows, s);
    X.nnz = nnz;
    X.col.alloc((size_t)std::max<int64_t>(nnz, 1));
    X.val.alloc((size_t)std::max<int64_t>(nnz, 1));
    k_x2f_rows<true><<<stream_grid(rows64), kThreads, 0, s>>>(N, nF, cutF, grad.p, T.ptr.p, T.col.p, T.val.p, X.ptr.p, X.col.p, X.val.p);
    MK_CHECK_LAUNCH();
}

}  // namespace mk
