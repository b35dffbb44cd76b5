// Shared by the capi_*.cu translation units: context struct, exception -> status translation, op timing.
#pragma once
#include "common.cuh"
#include "field.cuh"
#include "metriko_b200.h"

struct mk_ctx {
    mk::Ctx c;
};

// Handles keep a raw mk_ctx* (stream, error string). mk_ctx_destroy with live handles therefore only marks the context; the object and its
// stream go when the last handle is destroyed (one host thread per context, so a plain counter is enough).
inline void mk_ctx_finalize(mk_ctx* ctx) {
    if (ctx->c.owns_stream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}
struct CtxHold {
    mk_ctx* c = nullptr;
    CtxHold() = default;
    CtxHold(const CtxHold&) = delete;
    CtxHold& operator=(const CtxHold&) = delete;
    void attach(mk_ctx* x) { c = x; x->c.handles++; }
    ~CtxHold() {
        if (c && --c->c.handles == 0 && c->c.zombie) mk_ctx_finalize(c);
    }
};

#define MK_API_BEGIN(ctxp) try { mk::AllocStreamScope _alloc_scope((ctxp)->c.stream, (ctxp)->c.arena.get());
#define MK_API_END_RC(ctxp, rc)                                                     \
    return (rc);                                                                    \
    } catch (const mk::CudaError& e) { (ctxp)->c.err = e.what(); return MK_ERR_CUDA; } \
    catch (const std::invalid_argument& e) { (ctxp)->c.err = e.what(); return MK_ERR_ARG; } \
    catch (const std::exception& e) { (ctxp)->c.err = e.what(); return MK_ERR_INTERNAL; } \
    catch (...) { (ctxp)->c.err = "unknown error"; return MK_ERR_INTERNAL; }
#define MK_API_END(ctxp) MK_API_END_RC(ctxp, MK_OK)

namespace mk {
// average ms per launch of op.apply over `reps` launches (after 3 warm-ups), CUDA events on the launching stream
template <class T>
inline double time_op(LinOp<T>& op, const T* x, T* y, PcgState* st, ReduceWs ws, int reps, cudaStream_t s) {
    cudaEvent_t e0, e1;
    MK_CUDA(cudaEventCreate(&e0));
    MK_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) op.apply(x, y, st, ws, s);
    MK_CUDA(cudaStreamSynchronize(s));
    MK_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) op.apply(x, y, st, ws, s);
    MK_CUDA(cudaEventRecord(e1, s));
    MK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    MK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return (double)ms / reps;
}
}  // namespace mk
