// Handle structs shared by the capi_*.cu translation units (the field handle is also read by the IGM entry points: the eigen-solve
// result stays on the device between the two stages).
#pragma once
#include <memory>

#include "amg.cuh"
#include "capi_util.cuh"
#include "field.cuh"

using namespace mk;  // internal header of the capi_*.cu units only

struct mk_field {
    CtxHold hold;                   // very first member: released last (may finish a deferred mk_ctx_destroy)
    std::shared_ptr<Arena> arena;   // destroyed after every buffer below went back to it
    mk_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;  // copies, so that destruction does not depend on the context object
    int device = 0;
    int nF = 0, nH = 0, N = 4;
    DevBuf<int> face2half, next, twin, face;
    DevBuf<double> farg;
    FaceStencil S;
    bool assembled = false;
    std::unique_ptr<FaceOp> op;
    std::unique_ptr<Pcg<cdouble>> pcg;
    std::unique_ptr<AmgPrecond<cdouble>> amg;       // fp64 hierarchy (MK_PRECOND_FP32=0, or the host mirror path)
    std::unique_ptr<MixedPrecond<cdouble>> amg_lo;  // fp32 V-cycle (default)
    Precond<cdouble>* precond() { return amg_lo ? (Precond<cdouble>*)amg_lo.get() : (Precond<cdouble>*)amg.get(); }
    DevBuf<cdouble> u, x, b;
    RecycleSpace rec;
    bool has_solution = false;
};

