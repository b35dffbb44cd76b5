// C-ABI: host-side multigrid hierarchy set-up exposed for CPU tests (no device work in this file).
#include <complex>
#include <memory>

#include "amg.cuh"
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

int mk_amg_host_build(int n, const int* ptr, const int* col, const double* val, int is_complex, const double* cand, mk_amg_host** out) {
    if (n <= 0 || !ptr || !col || !val || !out) return MK_ERR_ARG;
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
