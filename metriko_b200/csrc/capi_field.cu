// C-ABI entry points: context + N-RoSy field (see include/metriko_b200.h for the reference seams they replace).
#include <chrono>
#include <complex>
#include <cstring>

#include "amg_dev.cuh"
#include "capi_util.cuh"
#include "capi_handles.cuh"
#include "field.cuh"
#include "metriko_b200.h"

using namespace mk;

extern "C" {

void mk_default_opts(mk_solve_opts* o) {
    o->tol = 1e-10;
    o->max_iter = 100000;
    o->precond = 1;
    o->use_graph = 1;
    o->recycle = 64;
}

int mk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mk_ctx_create(int device, void* stream, mk_ctx** out) {
    if (!out) return MK_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return MK_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return MK_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MK_ERR_CUDA;
    if (prop.major != 10) return MK_ERR_CUDA;  // sm_100a cubin only: no other architecture can run it
    // (device memory comes from the per-context arena, solver.cuh: the process-wide stream-ordered pool of the host application is left alone)
    mk_ctx* c = new mk_ctx();
    c->c.device = device;
    c->c.arena = std::make_shared<Arena>();
    if (stream) {
        c->c.stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->c.stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MK_ERR_CUDA; }
        c->c.owns_stream = true;
    }
    *out = c;
    return MK_OK;
}

int mk_ctx_destroy(mk_ctx* ctx) {
    if (!ctx) return MK_OK;
    if (ctx->c.handles > 0) {  // handles still refer to the stream and the error string: the last one to be destroyed finalizes the context
        ctx->c.zombie = true;
        return MK_OK;
    }
    mk_ctx_finalize(ctx);
    return MK_OK;
}

const char* mk_last_error(mk_ctx* ctx) { return ctx ? ctx->c.err.c_str() : "null context"; }

int mk_ctx_sync(mk_ctx* ctx) {
    MK_API_BEGIN(ctx)
    MK_CUDA(cudaStreamSynchronize(ctx->c.stream));
    MK_API_END(ctx)
}

int mk_field_create(mk_ctx* ctx, int nF, int nH, int N, const int* face2half, const int* next, const int* twin, const int* face,
                    const double* heArgOnFace, const double* faceArea, mk_field** out) {
    if (!ctx || !out || nF <= 0 || nH != 3 * nF || N <= 0) return MK_ERR_ARG;
    MK_API_BEGIN(ctx)
    MK_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.stream;
    auto f = std::make_unique<mk_field>();
    f->arena = ctx->c.arena;
    f->ctx = ctx;
    f->hold.attach(ctx);
    f->stream = s;
    f->device = ctx->c.device;
    f->nF = nF; f->nH = nH; f->N = N;
    f->face2half.alloc(nF); f->next.alloc(nH); f->twin.alloc(nH); f->face.alloc(nH); f->farg.alloc(nH);
    f->S.nF = nF;
    f->S.nbr.alloc((size_t)3 * nF); f->S.val.alloc((size_t)3 * nF); f->S.mass.alloc(nF);
    f->face2half.upload(face2half, nF, s); f->next.upload(next, nH, s); f->twin.upload(twin, nH, s);
    f->face.upload(face, nH, s); f->farg.upload(heArgOnFace, nH, s); f->S.mass.upload(faceArea, nF, s);
    f->u.alloc(nF); f->x.alloc(nF); f->b.alloc(nF);
    MK_CUDA(cudaStreamSynchronize(s));
    *out = f.release();
    MK_API_END(ctx)
}

int mk_field_destroy(mk_field* f) {
    if (!f) return MK_OK;
    AllocStreamScope sc(f->stream);
    cudaSetDevice(f->device);
    delete f;
    return MK_OK;
}

int mk_field_assemble(mk_field* f) {
    if (!f) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    k_field_assemble<<<stream_grid(f->nF), kThreads, 0, f->ctx->c.stream>>>(f->nF, f->N, f->face2half.p, f->next.p, f->twin.p, f->face.p,
                                                                              f->farg.p, f->S.nbr.p, f->S.val.p);
    MK_CHECK_LAUNCH();
    f->assembled = true;
    if (f->pcg) f->pcg->invalidate_graph();  // the captured iteration refers to the operator / hierarchy dropped below
    f->amg.reset();
    f->amg_lo.reset();
    f->op = std::make_unique<FaceOp>(&f->S);
    MK_API_END(f->ctx)
}

int mk_field_get_operator(mk_field* f, int* rowptr, int* col, double* val, double* mass) {
    if (!f || !f->assembled) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    cudaStream_t s = f->ctx->c.stream;
    const int nF = f->nF;
    std::vector<int> nbr((size_t)3 * nF);
    std::vector<cdouble> v((size_t)3 * nF);
    f->S.nbr.download(nbr.data(), nbr.size(), s);
    f->S.val.download(v.data(), v.size(), s);
    if (mass) f->S.mass.download(mass, nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < nF; i++) {
        rowptr[i] = 4 * i;
        col[4 * i] = i;
        val[8 * i] = f->S.diag;
        val[8 * i + 1] = 0.0;
        for (int j = 0; j < 3; j++) {
            col[4 * i + 1 + j] = nbr[(size_t)j * nF + i];
            val[8 * i + 2 + 2 * j] = v[(size_t)j * nF + i].x;
            val[8 * i + 3 + 2 * j] = v[(size_t)j * nF + i].y;
        }
    }
    rowptr[nF] = 4 * nF;
    MK_API_END(f->ctx)
}

// host set-up (the CPU mirror; also the fallback for degenerate stencils with repeated neighbours)
static std::unique_ptr<AmgPrecond<cdouble>> build_face_amg_host(mk_field* f) {
    cudaStream_t s = f->ctx->c.stream;
    const int nF = f->nF;
    StageTimer tm("field_amg");
    std::vector<int> nbr((size_t)3 * nF);
    std::vector<std::complex<double>> v((size_t)3 * nF);
    f->S.nbr.download(nbr.data(), nbr.size(), s);
    MK_CUDA(cudaMemcpyAsync(v.data(), f->S.val.p, v.size() * sizeof(cdouble), cudaMemcpyDeviceToHost, s));
    MK_CUDA(cudaStreamSynchronize(s));
    tm.lap("download stencil");
    HostCsr<std::complex<double>> A;
    A.rows = A.cols = nF;
    A.ptr.resize(nF + 1);
    A.col.resize((size_t)4 * nF);
    A.val.resize((size_t)4 * nF);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nF; i++) {
        std::pair<int, std::complex<double>> e[4] = {{i, {f->S.diag, 0.0}},
                                                     {nbr[i], v[i]},
                                                     {nbr[(size_t)nF + i], v[(size_t)nF + i]},
                                                     {nbr[2 * (size_t)nF + i], v[2 * (size_t)nF + i]}};
        std::sort(e, e + 4, [](const auto& a, const auto& b) { return a.first < b.first; });
        A.ptr[i] = 4 * i;
        for (int j = 0; j < 4; j++) { A.col[(size_t)4 * i + j] = e[j].first; A.val[(size_t)4 * i + j] = e[j].second; }
    }
    A.ptr[nF] = 4 * nF;
    tm.lap("host CSR");
    return build_amg<cdouble>(s, std::move(A), {}, false, f->op.get(), amg_params_from_env());
}

// Multigrid hierarchy for the face operator, built on the device from a CSR copy of the stencil; level 0 keeps applying the
// stencil kernel. MK_SETUP=host selects the host mirror.
static std::unique_ptr<AmgPrecond<cdouble>> build_face_amg(mk_field* f) {
    if (setup_on_host()) return build_face_amg_host(f);
    cudaStream_t s = f->ctx->c.stream;
    const int nF = f->nF;
    DeviceCsr<cdouble> A;
    A.rows = A.cols = nF;
    A.nnz = (int64_t)4 * nF;
    A.ptr.alloc((size_t)nF + 1);
    A.col.alloc((size_t)4 * nF);
    A.val.alloc((size_t)4 * nF);
    DevBuf<int> dup(1);
    dup.zero(s);
    k_face_to_csr<<<(nF + kThreads) / kThreads, kThreads, 0, s>>>(nF, f->S.diag, f->S.nbr.p, f->S.val.p, A.ptr.p, A.col.p, A.val.p, dup.p);
    MK_CHECK_LAUNCH();
    if (dev_read_int(dup.p, s)) return build_face_amg_host(f);
    return build_amg_device<cdouble>(s, std::move(A), f->op.get(), amg_params_from_env());
}

static void field_setup_solver(mk_field* f, const mk_solve_opts& o) {
    if (!f->pcg) f->pcg = std::make_unique<Pcg<cdouble>>(f->nF, f->ctx->c.stream);
    if (o.precond == 1 && !f->amg && !f->amg_lo) {
        f->amg = build_face_amg(f);
        if (precond_fp32()) {
            cudaStream_t s = f->ctx->c.stream;
            f->amg_lo = lower_precision<cdouble>(s, std::move(f->amg), std::make_unique<FaceOpLo>(&f->S, s));
        }
    }
}

static int field_solve_impl(mk_field* f, const cdouble* u0_dev, int n_iter, const mk_solve_opts* opts, mk_stats* stats) {
    mk_solve_opts o;
    if (opts) o = *opts; else mk_default_opts(&o);
    auto t0 = std::chrono::steady_clock::now();
    field_setup_solver(f, o);
    cudaStream_t s = f->ctx->c.stream;
    MK_CUDA(cudaStreamSynchronize(s));
    auto t1 = std::chrono::steady_clock::now();
    MK_CUDA(cudaMemcpyAsync(f->u.p, u0_dev, sizeof(cdouble) * f->nF, cudaMemcpyDeviceToDevice, s));
    MK_CUDA(cudaMemcpyAsync(f->x.p, u0_dev, sizeof(cdouble) * f->nF, cudaMemcpyDeviceToDevice, s));
    f->rec.init(f->nF, std::min(o.recycle, n_iter), s);
    EigStats es = inverse_iteration(f->ctx->c, *f->op, o.precond == 1 ? f->precond() : nullptr, f->S.mass.p, f->nF, f->u.p, f->x.p,
                                    f->b.p, *f->pcg, n_iter, o.tol, o.max_iter, o.use_graph != 0, &f->rec);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t2 = std::chrono::steady_clock::now();
    f->has_solution = true;
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
    return es.status == 0 ? MK_OK : (es.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
}

int mk_field_solve_dev(mk_field* f, const double* u0_dev, int n_iter, const mk_solve_opts* opts, double* x_dev, mk_stats* stats) {
    if (!f || !f->assembled || !u0_dev || n_iter < 0) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    rc = field_solve_impl(f, (const cdouble*)u0_dev, n_iter, opts, stats);
    if (x_dev) MK_CUDA(cudaMemcpyAsync(x_dev, f->x.p, sizeof(cdouble) * f->nF, cudaMemcpyDeviceToDevice, f->ctx->c.stream));
    MK_CUDA(cudaStreamSynchronize(f->ctx->c.stream));
    MK_API_END_RC(f->ctx, rc)
}

int mk_field_solve(mk_field* f, const double* u0, int n_iter, const mk_solve_opts* opts, double* x, mk_stats* stats) {
    if (!f || !f->assembled || !u0 || !x || n_iter < 0) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    cudaStream_t s = f->ctx->c.stream;
    f->b.upload((const cdouble*)u0, f->nF, s);  // stage through b (overwritten by the first iteration)
    DevBuf<cdouble> u0d(f->nF);
    MK_CUDA(cudaMemcpyAsync(u0d.p, f->b.p, sizeof(cdouble) * f->nF, cudaMemcpyDeviceToDevice, s));
    rc = field_solve_impl(f, u0d.p, n_iter, opts, stats);
    f->x.download((cdouble*)x, f->nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END_RC(f->ctx, rc)
}

// One Hermitian solve L x = b on the device operator (b in f->b, result in f->x): what solveSquare(L - 0 M, rhs) does (sls.h:16-27)
static int field_square_impl(mk_field* f, const mk_solve_opts* opts, mk_stats* stats) {
    mk_solve_opts o;
    if (opts) o = *opts; else mk_default_opts(&o);
    auto t0 = std::chrono::steady_clock::now();
    field_setup_solver(f, o);
    cudaStream_t s = f->ctx->c.stream;
    MK_CUDA(cudaStreamSynchronize(s));
    auto t1 = std::chrono::steady_clock::now();
    f->x.zero(s);
    f->pcg->bind(f->op.get(), o.precond == 1 ? f->precond() : nullptr, nullptr);
    SolveStats ss = f->pcg->solve(f->b.p, f->x.p, o.tol, o.max_iter, false, 16, o.use_graph != 0);
    MK_CUDA(cudaStreamSynchronize(s));
    auto t2 = std::chrono::steady_clock::now();
    f->has_solution = true;
    const int rc = ss.status == 0 ? MK_OK : (ss.status == 2 ? MK_ERR_NOT_SPD : MK_ERR_NOT_CONVERGED);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->outer_iterations = 1;
        stats->inner_iterations = ss.iterations;
        stats->inner_max = ss.iterations;
        stats->relres_max = ss.relres;
        stats->kernel_launches = ss.launches;
        stats->setup_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats->solve_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
        stats->status = rc;
    }
    return rc;
}

int mk_field_solve_square(mk_field* f, const double* rhs, const mk_solve_opts* opts, double* x, mk_stats* stats) {
    if (!f || !f->assembled || !rhs || !x) return MK_ERR_ARG;
    int rc = MK_OK;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    cudaStream_t s = f->ctx->c.stream;
    f->b.upload((const cdouble*)rhs, f->nF, s);
    rc = field_square_impl(f, opts, stats);
    f->x.download((cdouble*)x, f->nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END_RC(f->ctx, rc)
}

int mk_field_solve_curvature_aligned(mk_field* f, const double* heLen, const double* dihedralArg, const mk_solve_opts* opts, double* rhs_out, double* x,
                                     mk_stats* stats) {
    if (!f || !f->assembled || !heLen || !dihedralArg || !x || (f->N != 2 && f->N != 4)) return MK_ERR_ARG;  // assert(rosyN == 2 || rosyN == 4), frf.h:26
    int rc = MK_OK;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    cudaStream_t s = f->ctx->c.stream;
    if (!f->pcg) f->pcg = std::make_unique<Pcg<cdouble>>(f->nF, s);
    DevBuf<double> len(f->nH), darg(f->nH);
    len.upload(heLen, f->nH, s);
    darg.upload(dihedralArg, f->nH, s);
    const int grid = stream_grid(f->nF);
    k_curvature_rhs<<<grid, kThreads, 0, s>>>(f->nF, f->N, f->face2half.p, f->next.p, f->farg.p, len.p, darg.p, f->S.mass.p, f->b.p, f->pcg->state(),
                                              f->pcg->ws());
    k_scale_by_aux0<<<grid, kThreads, 0, s>>>(f->nF, f->b.p, f->pcg->state());
    MK_CHECK_LAUNCH();
    if (rhs_out) f->b.download((cdouble*)rhs_out, f->nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    rc = field_square_impl(f, opts, stats);
    f->x.download((cdouble*)x, f->nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END_RC(f->ctx, rc)
}

int mk_field_matching(mk_field* f, int nV, int nE, int matching_type, const double* pos, const int* tail, const int* head, const int* edge2face,
                      const double* faceBasisX, const double* faceBasisY, const int* edge, const int* prev, const int* vert2half,
                      const double* angleDefect, int* matching, double* effort, int* singular) {
    if (!f || !f->has_solution || nV <= 0 || nE <= 0 || 2 * nE != f->nH || (matching_type != 0 && matching_type != 1) || !pos || !tail || !head ||
        !edge2face || !faceBasisX || !faceBasisY || !edge || !prev || !vert2half || !angleDefect || !matching || !effort || !singular)
        return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    cudaStream_t s = f->ctx->c.stream;
    const int nH = f->nH, nF = f->nF;
    DevBuf<double> d_pos((size_t)3 * nV), d_bx((size_t)3 * nF), d_by((size_t)3 * nF), d_def(nV), d_eff(nE);
    DevBuf<int> d_tail(nH), d_head(nH), d_e2f((size_t)2 * nE), d_edge(nH), d_prev(nH), d_v2h(nV), d_match(nE), d_sing(nV);
    d_pos.upload(pos, (size_t)3 * nV, s); d_bx.upload(faceBasisX, (size_t)3 * nF, s); d_by.upload(faceBasisY, (size_t)3 * nF, s);
    d_def.upload(angleDefect, nV, s); d_tail.upload(tail, nH, s); d_head.upload(head, nH, s); d_e2f.upload(edge2face, (size_t)2 * nE, s);
    d_edge.upload(edge, nH, s); d_prev.upload(prev, nH, s); d_v2h.upload(vert2half, nV, s);
    k_edge_matching<<<stream_grid(nE), kThreads, 0, s>>>(nE, f->N, matching_type, f->x.p, d_pos.p, d_tail.p, d_head.p, d_e2f.p, d_bx.p, d_by.p, d_match.p,
                                                         d_eff.p);
    k_vertex_singular<<<stream_grid(nV), kThreads, 0, s>>>(nV, f->N, d_v2h.p, f->twin.p, d_prev.p, d_edge.p, d_def.p, d_eff.p, d_sing.p);
    MK_CHECK_LAUNCH();
    d_match.download(matching, nE, s);
    d_eff.download(effort, nE, s);
    d_sing.download(singular, nV, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END(f->ctx)
}

int mk_field_decode(mk_field* f, double* field) {
    if (!f || !f->has_solution || !field) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    cudaStream_t s = f->ctx->c.stream;
    DevBuf<cdouble> out((size_t)f->nF * f->N);
    k_decode<<<stream_grid(f->nF), kThreads, 0, s>>>(f->nF, f->N, f->x.p, out.p);
    MK_CHECK_LAUNCH();
    out.download((cdouble*)field, out.n, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END(f->ctx)
}

int mk_field_spmv(mk_field* f, const double* x, double* y) {
    if (!f || !f->assembled || !x || !y) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    cudaStream_t s = f->ctx->c.stream;
    DevBuf<cdouble> dx(f->nF), dy(f->nF);
    dx.upload((const cdouble*)x, f->nF, s);
    f->op->apply(dx.p, dy.p, nullptr, ReduceWs{nullptr, nullptr}, s);
    dy.download((cdouble*)y, f->nF, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_API_END(f->ctx)
}

int mk_field_spmv_bench(mk_field* f, int reps, int fused_dot, double* ms_per_launch) {
    if (!f || !f->assembled || reps <= 0 || !ms_per_launch) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    cudaStream_t s = f->ctx->c.stream;
    if (!f->pcg) f->pcg = std::make_unique<Pcg<cdouble>>(f->nF, s);
    DevBuf<cdouble> dx(f->nF), dy(f->nF);
    std::vector<cdouble> hx(f->nF);
    for (int i = 0; i < f->nF; i++) hx[i] = make_c(std::sin(0.37 * i), std::cos(0.11 * i));
    dx.upload(hx.data(), f->nF, s);
    PcgState init;
    std::memset(&init, 0, sizeof(init));
    init.rho = 1.0;
    MK_CUDA(cudaMemcpyAsync(f->pcg->state(), &init, sizeof(init), cudaMemcpyHostToDevice, s));
    *ms_per_launch = time_op(*f->op, dx.p, dy.p, fused_dot ? f->pcg->state() : nullptr, f->pcg->ws(), reps, s);
    MK_API_END(f->ctx)
}

int mk_field_profile_cycle(mk_field* f, int reps, int cap, int* n_out, double* us) {
    if (!f || reps <= 0 || !n_out || !us) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    MK_CUDA(cudaSetDevice(f->ctx->c.device));
    std::vector<double> v;
    if (f->amg_lo) v = profile_cycle<cdouble, cfloat>(*f->amg_lo, *f->amg_lo->lo, f->nF, reps, f->ctx->c.stream);
    else if (f->amg) v = profile_cycle<cdouble, cdouble>(*f->amg, *f->amg, f->nF, reps, f->ctx->c.stream);
    *n_out = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < cap; i++) us[i] = v[i];
    MK_API_END(f->ctx)
}

int mk_field_hierarchy(mk_field* f, int cap, int* n_levels, int64_t* rows, int64_t* nnz_a, int64_t* nnz_p) {
    if (!f || !n_levels) return MK_ERR_ARG;
    MK_API_BEGIN(f->ctx)
    std::vector<int64_t> r, a, p;
    if (f->amg_lo) f->amg_lo->lo->describe(r, a, p);
    else if (f->amg) f->amg->describe(r, a, p);
    if (!r.empty() && a[0] == 0) a[0] = (int64_t)4 * f->nF;  // the face stencil: diagonal + three neighbours per row
    *n_levels = (int)r.size();
    for (int l = 0; l < (int)r.size() && l < cap; l++) {
        if (rows) rows[l] = r[l];
        if (nnz_a) nnz_a[l] = a[l];
        if (nnz_p) nnz_p[l] = p[l];
    }
    MK_API_END(f->ctx)
}

}  // extern "C"
