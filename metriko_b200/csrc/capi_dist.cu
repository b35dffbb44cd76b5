// C-ABI entry points of the partitioned multi-GPU face operator (see include/metriko_b200.h, "partitioned" section).
#include <chrono>
#include <complex>
#include <cstring>

#include "capi_util.cuh"
#include "dist.cuh"
#include "metriko_b200.h"

using namespace mk;

struct mk_dist {
    CtxHold hold;
    mk_ctx* ctx = nullptr;
    DistComm c;
};

struct mk_dfield {
    CtxHold hold;                   // very first member: released last
    std::shared_ptr<Arena> arena;   // destroyed after every buffer below
    mk_dist* d = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
    DistFaceOp op;
    std::unique_ptr<FaceOp> blk_op;               // level 0 of the local hierarchy
    std::unique_ptr<AmgPrecond<cdouble>> amg;     // multigrid on the diagonal block (block-Jacobi across ranks; opts.precond = 2)
    std::unique_ptr<DistAmgPrecond> gamg;         // global hierarchy, level 0 distributed, coarse levels replicated (opts.precond = 1)
    std::unique_ptr<DistPcg> pcg;
    DevBuf<cdouble> u, x_ext, b;
    std::vector<int> h_nbr;                       // host copies for the hierarchy set-up
    std::vector<std::complex<double>> h_val;
};

extern "C" {

int mk_dist_unique_id(char* id, int id_bytes) {
    if (!id || id_bytes < (int)sizeof(ncclUniqueId)) return MK_ERR_ARG;
    try {
        ncclUniqueId u;
        if (NcclApi::get().GetUniqueId(&u) != ncclSuccess) return MK_ERR_CUDA;
        std::memset(id, 0, id_bytes);
        std::memcpy(id, &u, sizeof(u));
        return MK_OK;
    } catch (...) {
        return MK_ERR_INTERNAL;
    }
}

int mk_dist_create(mk_ctx* ctx, int rank, int world, const char* id, int id_bytes, mk_dist** out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world) return MK_ERR_ARG;
    if (world > 1 && (!id || id_bytes < (int)sizeof(ncclUniqueId))) return MK_ERR_ARG;
    MK_API_BEGIN(ctx)
    MK_CUDA(cudaSetDevice(ctx->c.device));
    auto d = std::make_unique<mk_dist>();
    d->ctx = ctx;
    d->hold.attach(ctx);
    d->c.rank = rank;
    d->c.world = world;
    MK_CUDA(cudaStreamCreateWithFlags(&d->c.comm_stream, cudaStreamNonBlocking));
    if (world > 1) {
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        MK_NCCL(NcclApi::get().CommInitRank(&d->c.comm, world, u, rank));
    }
    *out = d.release();
    MK_API_END(ctx)
}

int mk_dist_destroy(mk_dist* d) {
    delete d;
    return MK_OK;
}

int mk_dfield_create(mk_dist* d, int n_local, int n_interior, int n_halo, const int* nbr_local, const double* val, const double* mass,
                     const int* send_counts, const int* recv_counts, const int* send_idx, mk_dfield** out) {
    if (!d || !out || n_local <= 0 || n_interior < 0 || n_interior > n_local || n_halo < 0 || !nbr_local || !val || !mass || !send_counts || !recv_counts)
        return MK_ERR_ARG;
    MK_API_BEGIN(d->ctx)
    MK_CUDA(cudaSetDevice(d->ctx->c.device));
    cudaStream_t s = d->ctx->c.stream;
    auto f = std::make_unique<mk_dfield>();
    f->arena = d->ctx->c.arena;
    f->d = d;
    f->hold.attach(d->ctx);
    f->stream = s;
    f->device = d->ctx->c.device;
    DistFaceOp& op = f->op;
    op.dc = &d->c;
    op.nloc = n_local; op.nint = n_interior; op.nhalo = n_halo;
    const int W = d->c.world;
    op.send_counts.assign(send_counts, send_counts + W);
    op.recv_counts.assign(recv_counts, recv_counts + W);
    op.send_off.assign(W, 0);
    op.recv_off.assign(W, 0);
    int so = 0, ro = 0;
    for (int p = 0; p < W; p++) { op.send_off[p] = so; op.recv_off[p] = ro; so += op.send_counts[p]; ro += op.recv_counts[p]; }
    if (ro != n_halo || (so > 0 && !send_idx)) throw std::invalid_argument("mk_dfield_create: halo plan inconsistent");
    op.nsend = so;
    for (size_t k = 0; k < (size_t)3 * n_local; k++)
        if (nbr_local[k] < 0 || nbr_local[k] >= n_local + n_halo) throw std::invalid_argument("mk_dfield_create: neighbour index out of range");
    op.S.nF = n_local;
    op.S.nbr.alloc((size_t)3 * n_local); op.S.val.alloc((size_t)3 * n_local); op.S.mass.alloc(n_local);
    op.S.nbr.upload(nbr_local, (size_t)3 * n_local, s);
    op.S.val.upload((const cdouble*)val, (size_t)3 * n_local, s);
    op.S.mass.upload(mass, n_local, s);
    if (so) { op.send_idx.alloc(so); op.send_idx.upload(send_idx, so, s); op.sendbuf.alloc(so); }
    // diagonal block: halo couplings dropped (value 0, column = the row itself)
    f->h_nbr.assign(nbr_local, nbr_local + (size_t)3 * n_local);
    f->h_val.resize((size_t)3 * n_local);
    std::memcpy(f->h_val.data(), val, sizeof(double) * 6 * (size_t)n_local);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < n_local; i++) {
            const size_t k = (size_t)j * n_local + i;
            if (f->h_nbr[k] >= n_local) { f->h_nbr[k] = i; f->h_val[k] = 0.0; }
        }
    op.Sblk.nF = n_local;
    op.Sblk.nbr.alloc((size_t)3 * n_local); op.Sblk.val.alloc((size_t)3 * n_local);
    op.Sblk.nbr.upload(f->h_nbr.data(), f->h_nbr.size(), s);
    op.Sblk.val.upload((const cdouble*)f->h_val.data(), f->h_val.size(), s);
    MK_CUDA(cudaEventCreateWithFlags(&op.ev_pack, cudaEventDisableTiming));
    MK_CUDA(cudaEventCreateWithFlags(&op.ev_halo, cudaEventDisableTiming));
    f->blk_op = std::make_unique<FaceOp>(&op.Sblk);
    f->pcg = std::make_unique<DistPcg>(n_local, n_local + n_halo, &d->c, s);
    f->u.alloc(n_local); f->x_ext.alloc((size_t)n_local + n_halo); f->b.alloc(n_local);
    f->x_ext.zero(s);
    MK_CUDA(cudaStreamSynchronize(s));
    *out = f.release();
    MK_API_END(d->ctx)
}

// Global multigrid hierarchy for the partitioned operator (DistAmgPrecond, dist.cuh): every rank passes the same global stencil
// (nbr / val [3][nF], global face ids, as mk_field_get_operator lays it out) and the global row of each of its local rows.
int mk_dfield_set_global(mk_dfield* f, int nF, const int* nbr, const double* val, const int* local_to_global) {
    if (!f || nF <= 0 || !nbr || !val || !local_to_global) return MK_ERR_ARG;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    const int nloc = f->op.nloc;
    for (int i = 0; i < nloc; i++)
        if (local_to_global[i] < 0 || local_to_global[i] >= nF) throw std::invalid_argument("mk_dfield_set_global: row map out of range");
    auto g = std::make_unique<DistAmgPrecond>();
    g->A = &f->op;
    g->dc = &f->d->c;
    g->prm = amg_params_from_env();
    g->nloc = nloc;
    g->next = nloc + f->op.nhalo;
    {
        DevBuf<int> dn((size_t)3 * nF), dup(1);
        DevBuf<cdouble> dv((size_t)3 * nF);
        dn.upload(nbr, (size_t)3 * nF, s);
        dv.upload((const cdouble*)val, (size_t)3 * nF, s);
        dup.zero(s);
        DeviceCsr<cdouble> A;
        A.rows = A.cols = nF;
        A.nnz = (int64_t)4 * nF;
        A.ptr.alloc((size_t)nF + 1);
        A.col.alloc((size_t)4 * nF);
        A.val.alloc((size_t)4 * nF);
        k_face_to_csr<<<(nF + kThreads) / kThreads, kThreads, 0, s>>>(nF, f->op.S.diag, dn.p, dv.p, A.ptr.p, A.col.p, A.val.p, dup.p);
        MK_CHECK_LAUNCH();
        if (dev_read_int(dup.p, s)) throw std::invalid_argument("mk_dfield_set_global: two faces share two edges (degenerate stencil)");
        g->G = build_amg_device<cdouble>(s, std::move(A), &g->null_op, g->prm);
    }
    if (g->ready()) {
        AmgLevelDev<cdouble>& L0 = g->G->lv[0];
        DevBuf<int> l2g(nloc);
        l2g.upload(local_to_global, nloc, s);
        csr_gather_rows<cdouble>(L0.P, l2g.p, nloc, g->P0, s);
        transpose_dev<cdouble>(g->P0, g->R0, true, s);
        build_sell(g->P0, s);
        build_sell(g->R0, s);
        g->dinv.alloc(nloc);
        k_gather_values<cdouble><<<(nloc + 255) / 256, 256, 0, s>>>(nloc, l2g.p, L0.dinv.p, g->dinv.p);
        MK_CHECK_LAUNCH();
        g->lmax = L0.lmax;
        g->n1 = g->G->lv[1].n;
        g->r.alloc(nloc); g->q.alloc(nloc); g->d_ext.alloc(g->next);
        g->d_ext.zero(s);
        MK_CUDA(cudaStreamSynchronize(s));
        // the replicated level-0 pieces are no longer needed
        L0.P = DeviceCsr<cdouble>();
        L0.R = DeviceCsr<cdouble>();
        L0.r.release(); L0.d.release(); L0.q.release(); L0.dinv.release();
        g->G->first_level = 1;   // level 0 runs distributed (DistAmgPrecond): no replicated work vectors of global length
        g->G->prepare(s);
        if (precond_fp32()) {
            // the replicated levels >= 1 in fp32 storage and arithmetic, like the single-GPU cycle (MK_PRECOND_FP32=0: the fp64 levels)
            g->Glo = lower_tail<cdouble>(s, *g->G, 1);
            g->Glo->prepare(s);
            const int W = f->d->c.world;
            const size_t padded = (size_t)((g->n1 + W - 1) / W) * W;   // reduce-scatter / all-gather want equal blocks
            g->b1_lo.alloc(padded);
            g->x1_lo.alloc(padded);
            g->b1_lo.zero(s);
            g->x1_lo.zero(s);
            const char* e = getenv("MK_DIST_L1");
            // (tiny hierarchies stay replicated: every rank must own a non-trivial block of level 1)
            if (W > 1 && g->Glo->lv.size() >= 3 && g->n1 >= 256 * W && !(e && std::strcmp(e, "0") == 0)) {
                g->L1 = std::make_unique<DistLevel1>();
                g->L1->build(&f->d->c, g->Glo->lv[0], g->Glo->lv[1].n, s);
            }
        }
        f->gamg = std::move(g);
    }
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END(f->d->ctx)
}

static Precond<cdouble>* dfield_precond(mk_dfield* f, const mk_solve_opts& o);

int mk_dfield_destroy(mk_dfield* f) {
    if (!f) return MK_OK;
    AllocStreamScope sc(f->stream);
    cudaSetDevice(f->device);
    delete f;
    return MK_OK;
}

int mk_dfield_spmv(mk_dfield* f, const double* x_local, double* y_local) {
    if (!f || !x_local || !y_local) return MK_ERR_ARG;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    f->x_ext.upload((const cdouble*)x_local, f->op.nloc, s);
    f->op.apply(f->x_ext.p, f->b.p, nullptr, s);
    f->b.download((cdouble*)y_local, f->op.nloc, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END(f->d->ctx)
}

int mk_dfield_spmv_bench(mk_dfield* f, int reps, int overlap, double* ms_per_launch) {
    if (!f || reps <= 0 || !ms_per_launch) return MK_ERR_ARG;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    const bool keep = f->op.overlap;
    f->op.overlap = overlap != 0;
    cudaEvent_t e0, e1;
    MK_CUDA(cudaEventCreate(&e0));
    MK_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) f->op.apply(f->x_ext.p, f->b.p, nullptr, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) f->op.apply(f->x_ext.p, f->b.p, nullptr, s);
    MK_CUDA(cudaEventRecord(e1, s));
    MK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    f->op.overlap = keep;
    *ms_per_launch = (double)ms / reps;
    MK_API_END(f->d->ctx)
}

// local multigrid hierarchy on the diagonal block
static void dfield_setup_precond(mk_dfield* f) {
    if (f->amg) return;
    const int n = f->op.nloc;
    HostCsr<std::complex<double>> A;
    A.rows = A.cols = n;
    A.ptr.resize(n + 1);
    A.col.reserve((size_t)4 * n);
    A.val.reserve((size_t)4 * n);
    for (int i = 0; i < n; i++) {
        A.ptr[i] = (int)A.col.size();
        std::pair<int, std::complex<double>> e[4];
        int m = 0;
        e[m++] = {i, {f->op.S.diag, 0.0}};
        for (int j = 0; j < 3; j++) {
            const size_t k = (size_t)j * n + i;
            if (f->h_nbr[k] != i) e[m++] = {f->h_nbr[k], f->h_val[k]};
        }
        std::sort(e, e + m, [](const auto& a, const auto& b) { return a.first < b.first; });
        for (int t = 0; t < m; t++) {
            if (t > 0 && e[t].first == e[t - 1].first) { A.val.back() += e[t].second; continue; }
            A.col.push_back(e[t].first);
            A.val.push_back(e[t].second);
        }
    }
    A.ptr[n] = (int)A.col.size();
    f->amg = build_amg<cdouble>(f->d->ctx->c.stream, std::move(A), {}, false, f->blk_op.get(), amg_params_from_env());
}

// opts.precond: 0 none, 1 the global hierarchy when mk_dfield_set_global was called (else the local one), 2 local block-Jacobi hierarchy
static Precond<cdouble>* dfield_precond(mk_dfield* f, const mk_solve_opts& o) {
    if (o.precond == 0) return nullptr;
    if (o.precond == 1 && f->gamg && f->gamg->ready()) return f->gamg.get();
    dfield_setup_precond(f);
    return f->amg.get();
}

// Inverse iteration (sls.h:44-51) on the partitioned operator: the same recurrence as the single-GPU driver, every scalar an
// all-reduce of local partial sums. u0_local / x_local: this rank's rows in the plan's local order, host, complex interleaved.
int mk_dfield_solve(mk_dfield* f, const double* u0_local, int n_iter, const mk_solve_opts* opts, double* x_local, mk_stats* stats) {
    if (!f || !u0_local || !x_local || n_iter < 0) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    mk_solve_opts o;
    if (opts) o = *opts; else mk_default_opts(&o);
    auto t0 = std::chrono::steady_clock::now();
    Precond<cdouble>* M = dfield_precond(f, o);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t1 = std::chrono::steady_clock::now();
    DistPcg& pcg = *f->pcg;
    DistFaceOp& A = f->op;
    const int n = A.nloc, grid = pcg.grid;
    f->u.upload((const cdouble*)u0_local, n, s);
    MK_CUDA(cudaMemcpyAsync(f->x_ext.p, f->u.p, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
    EigStats es;
    int hint = 0;
    double* part = pcg.part.p;
    for (int it = 0; it < n_iter; it++) {
        // b = M u ; theta = <u, b> / <u, L u> ; x = theta u
        k_dist_scale<<<grid, kThreads, 0, s>>>(n, A.S.mass.p, f->u.p, f->b.p, nullptr, 0);
        MK_CUDA(cudaMemcpyAsync(f->x_ext.p, f->u.p, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
        A.apply(f->x_ext.p, pcg.q.p, nullptr, s);
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, f->u.p, f->b.p, f->u.p, pcg.q.p, nullptr, part + 5, pcg.ws(), nullptr);
        f->d->c.allreduce_sum(part + 5, 2, s);
        double h[2];
        MK_CUDA(cudaMemcpyAsync(h, part + 5, sizeof(h), cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
        const double theta = h[1] > 0 ? h[0] / h[1] : 0.0;
        es.lambda = h[0] > 0 ? h[1] / h[0] : 0.0;
        MK_CUDA(cudaMemcpyAsync(part + 7, &theta, sizeof(double), cudaMemcpyHostToDevice, s));
        k_dist_scale<<<grid, kThreads, 0, s>>>(n, nullptr, f->u.p, f->x_ext.p, part + 7, 0);
        MK_CHECK_LAUNCH();
        SolveStats ss = pcg.solve(A, M, f->b.p, f->x_ext.p, o.tol, o.max_iter, 8, hint);
        hint = ss.iterations + 1;
        // x /= sqrt|x^H M x| ; u = x
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, f->x_ext.p, f->x_ext.p, nullptr, nullptr, A.S.mass.p, part + 5, pcg.ws(), nullptr);
        f->d->c.allreduce_sum(part + 5, 1, s);
        k_dist_scale<<<grid, kThreads, 0, s>>>(n, nullptr, f->x_ext.p, f->u.p, part + 5, 1);
        MK_CUDA(cudaMemcpyAsync(f->x_ext.p, f->u.p, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
        MK_CHECK_LAUNCH();
        es.outer++;
        es.inner_total += ss.iterations;
        es.inner_max = std::max(es.inner_max, ss.iterations);
        es.relres_max = std::max(es.relres_max, ss.relres);
        if (ss.status != 0) es.status = ss.status;
        es.launches += ss.launches + 8 + 2 * A.launches_per_apply();
    }
    f->x_ext.download((cdouble*)x_local, n, s);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t2 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->outer_iterations = es.outer;
        stats->inner_iterations = es.inner_total;
        stats->inner_max = es.inner_max;
        stats->relres_max = es.relres_max;
        stats->lambda = es.lambda;
        stats->kernel_launches = es.launches;
        stats->setup_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        stats->status = es.status == 0 ? MK_OK : (es.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    }
    rc = es.status == 0 ? MK_OK : (es.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    MK_API_END_RC(f->d->ctx, rc)
}

// One distributed PCG solve L x = b (host vectors, local rows): the "graph-partitioned SpMV/PCG" unit of BASELINE configs[3].
int mk_dfield_pcg(mk_dfield* f, const double* b_local, const double* x0_local, const mk_solve_opts* opts, double* x_local, mk_stats* stats) {
    if (!f || !b_local || !x_local) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    mk_solve_opts o;
    if (opts) o = *opts; else mk_default_opts(&o);
    auto t0 = std::chrono::steady_clock::now();
    Precond<cdouble>* M = dfield_precond(f, o);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t1 = std::chrono::steady_clock::now();
    const int n = f->op.nloc;
    f->b.upload((const cdouble*)b_local, n, s);
    if (x0_local) f->x_ext.upload((const cdouble*)x0_local, n, s);
    else f->x_ext.zero(s);
    SolveStats ss = f->pcg->solve(f->op, M, f->b.p, f->x_ext.p, o.tol, o.max_iter, 8, 0);
    f->x_ext.download((cdouble*)x_local, n, s);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t2 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->outer_iterations = 1;
        stats->inner_iterations = ss.iterations;
        stats->inner_max = ss.iterations;
        stats->relres_max = ss.relres;
        stats->kernel_launches = ss.launches;
        stats->setup_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        stats->status = ss.status == 0 ? MK_OK : (ss.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    }
    rc = ss.status == 0 ? MK_OK : (ss.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    MK_API_END_RC(f->d->ctx, rc)
}

int mk_dfield_pcg_dev(mk_dfield* f, const double* b_dev, double* x_dev, const mk_solve_opts* opts, mk_stats* stats) {
    if (!f || !b_dev || !x_dev) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->d->ctx)
    MK_CUDA(cudaSetDevice(f->d->ctx->c.device));
    cudaStream_t s = f->d->ctx->c.stream;
    mk_solve_opts o;
    if (opts) o = *opts; else mk_default_opts(&o);
    auto t0 = std::chrono::steady_clock::now();
    Precond<cdouble>* M = dfield_precond(f, o);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t1 = std::chrono::steady_clock::now();
    const int n = f->op.nloc;
    MK_CUDA(cudaMemcpyAsync(f->b.p, b_dev, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemcpyAsync(f->x_ext.p, x_dev, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
    SolveStats ss = f->pcg->solve(f->op, M, f->b.p, f->x_ext.p, o.tol, o.max_iter, 8, 0);
    MK_CUDA(cudaMemcpyAsync(x_dev, f->x_ext.p, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaStreamSynchronize(s));
    auto t2 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->outer_iterations = 1;
        stats->inner_iterations = ss.iterations;
        stats->inner_max = ss.iterations;
        stats->relres_max = ss.relres;
        stats->kernel_launches = ss.launches;
        stats->setup_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        stats->status = ss.status == 0 ? MK_OK : (ss.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    }
    rc = ss.status == 0 ? MK_OK : (ss.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    MK_API_END_RC(f->d->ctx, rc)
}

}  // extern "C"
