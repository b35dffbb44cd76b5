// C-ABI: multigrid hierarchy set-up exposed for tests -- the host mirror (no device needed) and the device build, whose levels are
// copied back into the same container so that one set of accessors serves both.
#include <complex>
#include <memory>

#include "amg_dev.cuh"
#include "capi_util.cuh"
#include "metriko_b200.h"

using namespace mk;

struct mk_amg_host {
    bool is_complex = false;
    std::vector<AmgHostLevel<double>> re;
    std::vector<AmgHostLevel<std::complex<double>>> cx;
};

template <class S>
static const HostCsr<S>* pick(const std::vector<AmgHostLevel<S>>& lv, int level, int which) {
    if (level < 0 || level >= (int)lv.size()) return nullptr;
    if (which == 0) return &lv[level].A;
    if (which == 1 && level + 1 < (int)lv.size()) return &lv[level].P;
    return nullptr;
}

extern "C" {

int mk_amg_host_build_typed(int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out);

int mk_amg_host_build(int n, const int* ptr, const int* col, const double* val, int is_complex, const double* cand, mk_amg_host** out) {
    if (n <= 0 || !ptr || !col || !val || !out) return MK_ERR_ARG;
    if (!cand) return mk_amg_host_build_typed(n, ptr, col, val, is_complex, nullptr, out);
    try {
        auto h = std::make_unique<mk_amg_host>();
        h->is_complex = is_complex != 0;
        const AmgParams prm = amg_params_from_env(!is_complex);
        const int64_t nnz = ptr[n];
        if (is_complex) {
            HostCsr<std::complex<double>> A;
            A.rows = A.cols = n;
            A.ptr.assign(ptr, ptr + n + 1);
            A.col.assign(col, col + nnz);
            A.val.resize(nnz);
            for (int64_t k = 0; k < nnz; k++) A.val[k] = {val[2 * k], val[2 * k + 1]};
            h->cx = build_hierarchy<std::complex<double>>(std::move(A), {}, false, prm);
        } else {
            HostCsr<double> A;
            A.rows = A.cols = n;
            A.ptr.assign(ptr, ptr + n + 1);
            A.col.assign(col, col + nnz);
            A.val.assign(val, val + nnz);
            std::vector<double> c;
            if (cand) c.assign(cand, cand + n);
            h->re = build_hierarchy<double>(std::move(A), std::move(c), cand != nullptr, prm);
        }
        *out = h.release();
        return MK_OK;
    } catch (...) {
        return MK_ERR_INTERNAL;
    }
}

}  // extern "C"

template <class T, class S>
static void device_levels_to_host(cudaStream_t s, AmgPrecond<T>& amg, std::vector<AmgHostLevel<S>>& out) {
    out.resize(amg.lv.size());
    for (size_t l = 0; l < amg.lv.size(); l++) {
        out[l].A = download_csr<T, S>(amg.lv[l].A, s);
        if (l + 1 < amg.lv.size()) out[l].P = download_csr<T, S>(amg.lv[l].P, s);
        out[l].dinv.resize(amg.lv[l].n);
        MK_CUDA(cudaMemcpyAsync(out[l].dinv.data(), amg.lv[l].dinv.p, sizeof(T) * amg.lv[l].n, cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
        out[l].lmax = amg.lv[l].lmax;
    }
}

template <class T>
static void upload_abi_csr(DeviceCsr<T>& A, int rows, int cols, const int* ptr, const int* col, const double* val, cudaStream_t s) {
    upload_csr_plain<T>(A, rows, cols, ptr, col, (const T*)val, (int64_t)ptr[rows], s);
    MK_CUDA(cudaStreamSynchronize(s));
}

extern "C" {

int mk_amg_dev_build_typed(mk_ctx* ctx, int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out);
int mk_amg_dev_build(mk_ctx* ctx, int n, const int* ptr, const int* col, const double* val, int is_complex, mk_amg_host** out) {
    return mk_amg_dev_build_typed(ctx, n, ptr, col, val, is_complex, nullptr, out);
}

int mk_amg_dev_build_typed(mk_ctx* ctx, int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out) {
    if (!ctx || n <= 0 || !ptr || !col || !val || !out) return MK_ERR_ARG;
    MK_API_BEGIN(ctx)
    MK_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    auto h = std::make_unique<mk_amg_host>();
    h->is_complex = is_complex != 0;
    const AmgParams prm = amg_params_from_env(!is_complex);
    if (is_complex) {
        DeviceCsr<cdouble> A;
        upload_abi_csr<cdouble>(A, n, n, ptr, col, val, s);
        if (!rows_sorted_strict(A, s)) throw std::invalid_argument("mk_amg_dev_build: rows must have strictly increasing columns");
        DevBuf<int> ty;
        if (types) { ty.alloc(n); ty.upload(types, n, s); MK_CUDA(cudaStreamSynchronize(s)); }
        auto amg = build_amg_device<cdouble>(s, std::move(A), nullptr, prm, std::move(ty));
        device_levels_to_host<cdouble, std::complex<double>>(s, *amg, h->cx);
    } else {
        DeviceCsr<double> A;
        upload_abi_csr<double>(A, n, n, ptr, col, val, s);
        if (!rows_sorted_strict(A, s)) throw std::invalid_argument("mk_amg_dev_build: rows must have strictly increasing columns");
        DevBuf<int> ty;
        if (types) { ty.alloc(n); ty.upload(types, n, s); MK_CUDA(cudaStreamSynchronize(s)); }
        auto amg = build_amg_device<double>(s, std::move(A), nullptr, prm, std::move(ty));
        device_levels_to_host<double, double>(s, *amg, h->re);
    }
    *out = h.release();
    MK_API_END(ctx)
}

// op 0: C = A B (device SpGEMM); op 1: C = A^T; op 2: C = A^H. The result is level 0 of the returned container.
int mk_dev_sparse_op(mk_ctx* ctx, int op, int is_complex, int a_rows, int a_cols, const int* a_ptr, const int* a_col, const double* a_val, int b_rows,
                     int b_cols, const int* b_ptr, const int* b_col, const double* b_val, mk_amg_host** out) {
    if (!ctx || !out || !a_ptr || !a_col || !a_val || op < 0 || op > 2 || (op == 0 && (!b_ptr || !b_col || !b_val))) return MK_ERR_ARG;
    MK_API_BEGIN(ctx)
    MK_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    auto h = std::make_unique<mk_amg_host>();
    h->is_complex = is_complex != 0;
    if (is_complex) {
        DeviceCsr<cdouble> A, B, C;
        upload_abi_csr<cdouble>(A, a_rows, a_cols, a_ptr, a_col, a_val, s);
        if (op == 0) {
            upload_abi_csr<cdouble>(B, b_rows, b_cols, b_ptr, b_col, b_val, s);
            spgemm<cdouble>(A, B, C, s);
        } else {
            transpose_dev<cdouble>(A, C, op == 2, s);
        }
        h->cx.resize(1);
        h->cx[0].A = download_csr<cdouble, std::complex<double>>(C, s);
    } else {
        DeviceCsr<double> A, B, C;
        upload_abi_csr<double>(A, a_rows, a_cols, a_ptr, a_col, a_val, s);
        if (op == 0) {
            upload_abi_csr<double>(B, b_rows, b_cols, b_ptr, b_col, b_val, s);
            spgemm<double>(A, B, C, s);
        } else {
            transpose_dev<double>(A, C, op == 2, s);
        }
        h->re.resize(1);
        h->re[0].A = download_csr<double, double>(C, s);
    }
    *out = h.release();
    MK_API_END(ctx)
}

int mk_amg_host_build_typed(int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out) {
    if (n <= 0 || !ptr || !col || !val || !out) return MK_ERR_ARG;
    try {
        auto h = std::make_unique<mk_amg_host>();
        h->is_complex = is_complex != 0;
        const AmgParams prm = amg_params_from_env(!is_complex);
        const int64_t nnz = ptr[n];
        std::vector<int> ty;
        if (types) ty.assign(types, types + n);
        if (is_complex) {
            HostCsr<std::complex<double>> A;
            A.rows = A.cols = n;
            A.ptr.assign(ptr, ptr + n + 1);
            A.col.assign(col, col + nnz);
            A.val.resize(nnz);
            for (int64_t k = 0; k < nnz; k++) A.val[k] = {val[2 * k], val[2 * k + 1]};
            h->cx = build_hierarchy<std::complex<double>>(std::move(A), {}, false, prm, std::move(ty));
        } else {
            HostCsr<double> A;
            A.rows = A.cols = n;
            A.ptr.assign(ptr, ptr + n + 1);
            A.col.assign(col, col + nnz);
            A.val.assign(val, val + nnz);
            h->re = build_hierarchy<double>(std::move(A), {}, false, prm, std::move(ty));
        }
        *out = h.release();
        return MK_OK;
    } catch (...) {
        return MK_ERR_INTERNAL;
    }
}

int mk_amg_host_num_levels(mk_amg_host* h) { return !h ? 0 : (int)(h->is_complex ? h->cx.size() : h->re.size()); }

int mk_amg_host_level_info(mk_amg_host* h, int level, int which, int* rows, int* cols, int64_t* nnz, double* lmax) {
    if (!h) return MK_ERR_ARG;
    int r, c;
    int64_t z;
    double lm;
    if (h->is_complex) {
        auto* M = pick(h->cx, level, which);
        if (!M) return MK_ERR_ARG;
        r = M->rows; c = M->cols; z = M->nnz(); lm = h->cx[level].lmax;
    } else {
        auto* M = pick(h->re, level, which);
        if (!M) return MK_ERR_ARG;
        r = M->rows; c = M->cols; z = M->nnz(); lm = h->re[level].lmax;
    }
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (nnz) *nnz = z;
    if (lmax) *lmax = lm;
    return MK_OK;
}

int mk_amg_host_level_get(mk_amg_host* h, int level, int which, int* ptr, int* col, double* val) {
    if (!h || !val) return MK_ERR_ARG;
    if (which == 2) {  // the smoother's inverse diagonal of level `level` (rows values; ptr / col unused)
        if (level < 0 || level >= mk_amg_host_num_levels(h)) return MK_ERR_ARG;
        if (h->is_complex) {
            const auto& d = h->cx[level].dinv;
            for (size_t k = 0; k < d.size(); k++) { val[2 * k] = d[k].real(); val[2 * k + 1] = d[k].imag(); }
        } else {
            const auto& d = h->re[level].dinv;
            std::copy(d.begin(), d.end(), val);
        }
        return MK_OK;
    }
    if (!ptr || !col) return MK_ERR_ARG;
    if (h->is_complex) {
        auto* M = pick(h->cx, level, which);
        if (!M) return MK_ERR_ARG;
        std::copy(M->ptr.begin(), M->ptr.end(), ptr);
        std::copy(M->col.begin(), M->col.end(), col);
        for (size_t k = 0; k < M->val.size(); k++) { val[2 * k] = M->val[k].real(); val[2 * k + 1] = M->val[k].imag(); }
    } else {
        auto* M = pick(h->re, level, which);
        if (!M) return MK_ERR_ARG;
        std::copy(M->ptr.begin(), M->ptr.end(), ptr);
        std::copy(M->col.begin(), M->col.end(), col);
        std::copy(M->val.begin(), M->val.end(), val);
    }
    return MK_OK;
}

int mk_amg_host_destroy(mk_amg_host* h) {
    delete h;
    return MK_OK;
}

}  // extern "C"
