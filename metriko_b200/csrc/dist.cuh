// Partitioned (one process per GPU) face operator: halo exchange over NCCL overlapped with the interior-row SpMV, and a
// distributed PCG / inverse iteration whose scalars stay on the device (NCCL all-reduce of 1-2 doubles per reduction).
// The reference has no distributed code; the host-side plan comes from metriko_b200/partition.py.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, reusing the copy a host application such as PyTorch already loaded),
// so libmetriko_b200.so itself has no link-time dependency on it and single-GPU users never touch it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <memory>
#include <vector>

#include "amg_dev.cuh"
#include "field.cuh"

namespace mk {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    static NcclApi& get() {
        static NcclApi api;
        if (!api.handle) {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy already in the process, if any
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
            auto sym = [&](const char* n) {
                void* p = dlsym(h, n);
                if (!p) throw std::runtime_error(std::string("libnccl: missing symbol ") + n);
                return p;
            };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
            api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
            api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
            api.Send = (decltype(api.Send))sym("ncclSend");
            api.Recv = (decltype(api.Recv))sym("ncclRecv");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            api.handle = h;
        }
        return api;
    }
};

#define MK_NCCL(expr)                                                                                               \
    do {                                                                                                            \
        ncclResult_t _r = (expr);                                                                                   \
        if (_r != ncclSuccess)                                                                                      \
            throw mk::CudaError(std::string(#expr) + " failed: " + mk::NcclApi::get().GetErrorString(_r));          \
    } while (0)

struct DistComm {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;  // halo exchange runs here, next to the interior SpMV on the main stream
    ~DistComm() {
        if (comm) NcclApi::get().CommDestroy(comm);
        if (comm_stream) cudaStreamDestroy(comm_stream);
    }
    void allreduce_sum(double* buf, int n, cudaStream_t s) {
        if (world > 1) MK_NCCL(NcclApi::get().AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, comm, s));
    }
    void allreduce_sum_f32(float* buf, int n, cudaStream_t s) {
        if (world > 1) MK_NCCL(NcclApi::get().AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, comm, s));
    }
};

// ---- kernels ----

// out[k] = x[idx[k]]  (boundary values each peer needs, grouped by peer)
static __global__ void __launch_bounds__(kThreads) k_halo_pack(int m, const int* __restrict__ idx, const cdouble* __restrict__ x, cdouble* __restrict__ out,
                                                        const PcgState* st) {
    if (st && st->done) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) stc(out + k, ldc_nc(x + __ldg(idx + k)));
}

// y = L x on the rows [r0, r1) of the local block; x = [owned | halo]; stencil arrays strided by `stride` (= owned rows)
static __global__ void __launch_bounds__(kThreads) k_spmv_face_rows(int stride, int r0, int r1, double diag, const int* __restrict__ nbr,
                                                             const cdouble* __restrict__ val, const cdouble* __restrict__ x, cdouble* __restrict__ y,
                                                             const PcgState* st) {
    if (st && st->done) return;
    for (int f = r0 + blockIdx.x * blockDim.x + threadIdx.x; f < r1; f += gridDim.x * blockDim.x) {
        const int n0 = __ldg(nbr + f), n1 = __ldg(nbr + (size_t)stride + f), n2 = __ldg(nbr + 2 * (size_t)stride + f);
        const cdouble v0 = ldc_nc(val + f), v1 = ldc_nc(val + (size_t)stride + f), v2 = ldc_nc(val + 2 * (size_t)stride + f);
        const cdouble xf = ldc_nc(x + f);
        const cdouble x0 = ldc_nc(x + n0), x1 = ldc_nc(x + n1), x2 = ldc_nc(x + n2);
        cdouble acc = make_c(diag * xf.x, diag * xf.y);
        acc = ST<cdouble>::fma_(v0, x0, acc);
        acc = ST<cdouble>::fma_(v1, x1, acc);
        acc = ST<cdouble>::fma_(v2, x2, acc);
        stc(y + f, acc);
    }
}

// out[0] = Re<a, b>, out[1] = Re<c, d> (c == nullptr: only the first), optionally weighted by w (mass): local partial sums
static __global__ void __launch_bounds__(kThreads) k_dist_dots(int n, const cdouble* __restrict__ a, const cdouble* __restrict__ b,
                                                        const cdouble* __restrict__ c, const cdouble* __restrict__ d, const double* __restrict__ w,
                                                        double* __restrict__ out, ReduceWs ws, const PcgState* st) {
    __shared__ double red[32];
    if (st && st->done) return;
    double s0 = 0.0, s1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double wi = w ? __ldg(w + i) : 1.0;
        const cdouble ai = ldc_nc(a + i), bi = ldc_nc(b + i);
        s0 += wi * (ai.x * bi.x + ai.y * bi.y);
        if (c) {
            const cdouble ci = ldc_nc(c + i), di = ldc_nc(d + i);
            s1 += wi * (ci.x * di.x + ci.y * di.y);
        }
    }
    double v[2] = {s0, s1};
    grid_reduce<2>(v, ws.partials, ws.ticket, red, [&](const double* s) { out[0] = s[0]; out[1] = s[1]; });
}

// x += alpha p ; r -= alpha q with alpha = rho / <p, q> (part[0], already all-reduced); part[1] = local <r, r>
static __global__ void __launch_bounds__(kThreads) k_dist_update_xr(int n, cdouble* __restrict__ x, cdouble* __restrict__ r, const cdouble* __restrict__ p,
                                                             const cdouble* __restrict__ q, const PcgState* st, double* __restrict__ part, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    const double alpha = st->rho / part[0];
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const cdouble pi = ldc_nc(p + i), qi = ldc_nc(q + i);
        cdouble xi = ldc(x + i), ri = ldc(r + i);
        xi = make_c(xi.x + alpha * pi.x, xi.y + alpha * pi.y);
        ri = make_c(ri.x - alpha * qi.x, ri.y - alpha * qi.y);
        stc(x + i, xi);
        stc(r + i, ri);
        acc += ri.x * ri.x + ri.y * ri.y;
    }
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) { part[1] = s[0]; });
}

// p = z + beta p with beta = <r, z>_new / rho_old (part[2] all-reduced; st->rho still holds the old value)
static __global__ void __launch_bounds__(kThreads) k_dist_update_p(int n, cdouble* __restrict__ p, const cdouble* __restrict__ z, const PcgState* st,
                                                            const double* __restrict__ part) {
    if (st->done) return;
    const double beta = part[2] / st->rho;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const cdouble zi = ldc_nc(z + i), pi = ldc(p + i);
        stc(p + i, make_c(zi.x + beta * pi.x, zi.y + beta * pi.y));
    }
}

// one thread: fold the all-reduced partials of an iteration into the state (after every kernel that read the old values)
static __global__ void k_dist_state(PcgState* st, const double* part) {
    if (st->done) return;
    st->pq = part[0];
    st->alpha = st->rho / part[0];
    st->rr = part[1];
    st->beta = part[2] / st->rho;
    st->rho = part[2];
    st->iter += 1;
    if (!(part[0] > 0.0) || !(part[1] == part[1])) st->done = 2;
    else if (part[1] <= st->thresh) st->done = 1;
    else if (st->iter >= st->maxit) st->done = 3;
}

// start of a solve: part = {<r,r>, <b,b>, <r,z>} all-reduced
static __global__ void k_dist_start(PcgState* st, const double* part, double tol2, int maxit, int keep_bnorm) {
    st->rr = part[0];
    if (!keep_bnorm) {
        st->bnorm2 = part[1];
        st->thresh = tol2 * part[1];
        st->iter = 0;
    }
    st->rho = part[2];
    st->maxit = maxit;
    st->done = (part[0] <= st->thresh) ? 1 : 0;
}

// r = b - q
static __global__ void __launch_bounds__(kThreads) k_dist_residual(int n, const cdouble* __restrict__ b, const cdouble* __restrict__ q, cdouble* __restrict__ r) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) stc(r + i, ldc_nc(b + i) - ldc_nc(q + i));
}

// y = a * m .* x (a read from the device: scale[0], or 1 / sqrt|scale[0]| when inv_sqrt)
static __global__ void __launch_bounds__(kThreads) k_dist_scale(int n, const double* __restrict__ m, const cdouble* __restrict__ x, cdouble* __restrict__ y,
                                                         const double* __restrict__ scale, int inv_sqrt) {
    double a = scale ? scale[0] : 1.0;
    if (scale && inv_sqrt) a = 1.0 / sqrt(fabs(a));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double mi = m ? a * __ldg(m + i) : a;
        const cdouble xi = ldc_nc(x + i);
        stc(y + i, make_c(mi * xi.x, mi * xi.y));
    }
}

// ---- the partitioned operator ----
struct DistFaceOp {
    DistComm* dc = nullptr;
    int nloc = 0, nint = 0, nhalo = 0;
    FaceStencil S;      // nbr indexes x_ext = [owned | halo]
    FaceStencil Sblk;   // the diagonal block only (halo couplings dropped): level 0 of the local preconditioner
    DevBuf<int> send_idx;
    DevBuf<cdouble> sendbuf;
    std::vector<int> send_counts, recv_counts, send_off, recv_off;
    int nsend = 0;
    cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;
    bool overlap = true;

    ~DistFaceOp() {
        if (ev_pack) cudaEventDestroy(ev_pack);
        if (ev_halo) cudaEventDestroy(ev_halo);
    }

    // y[0:nloc) = (L x)[owned rows]; x_ext's halo part is refreshed from the peers
    void apply(cdouble* x_ext, cdouble* y, const PcgState* st, cudaStream_t s) {
        NcclApi& nc = NcclApi::get();
        const bool comm = dc->world > 1 && (nsend > 0 || nhalo > 0);
        cudaStream_t cs = overlap ? dc->comm_stream : s;
        if (comm) {
            if (nsend) k_halo_pack<<<stream_grid(nsend), kThreads, 0, s>>>(nsend, send_idx.p, x_ext, sendbuf.p, st);
            if (overlap) {
                MK_CUDA(cudaEventRecord(ev_pack, s));
                MK_CUDA(cudaStreamWaitEvent(cs, ev_pack, 0));
            }
            MK_NCCL(nc.GroupStart());
            for (int p = 0; p < dc->world; p++) {
                if (p == dc->rank) continue;
                if (send_counts[p]) MK_NCCL(nc.Send(sendbuf.p + send_off[p], 2 * (size_t)send_counts[p], ncclDouble, p, dc->comm, cs));
                if (recv_counts[p]) MK_NCCL(nc.Recv(x_ext + nloc + recv_off[p], 2 * (size_t)recv_counts[p], ncclDouble, p, dc->comm, cs));
            }
            MK_NCCL(nc.GroupEnd());
            if (overlap) MK_CUDA(cudaEventRecord(ev_halo, cs));
        }
        if (nint > 0) k_spmv_face_rows<<<stream_grid(nint), kThreads, 0, s>>>(nloc, 0, nint, S.diag, S.nbr.p, S.val.p, x_ext, y, st);
        if (comm && overlap) MK_CUDA(cudaStreamWaitEvent(s, ev_halo, 0));
        if (nloc > nint) k_spmv_face_rows<<<stream_grid(nloc - nint), kThreads, 0, s>>>(nloc, nint, nloc, S.diag, S.nbr.p, S.val.p, x_ext, y, st);
        MK_CHECK_LAUNCH();
    }
    int launches_per_apply() const { return 3; }
};

// Multigrid preconditioner of the partitioned operator with the GLOBAL hierarchy (the same one the single-GPU path builds, so the
// iteration counts are those of one GPU, not of block-Jacobi):
//   level 0   : distributed -- Chebyshev smoothing with the partitioned operator (halo exchange per SpMV); this rank holds its rows
//               of P_0 and their conjugate transpose;
//   restrict  : every rank forms its partial sum R_0[:, mine] r_mine, one NCCL all-reduce of the level-1 vector adds them up;
//   levels>=1 : replicated -- every rank runs the same coarse V-cycle on the full level-1 vector (bit-identical results, no traffic);
//   prolong   : local, x_mine += P_0[mine, :] x_1.
// The hierarchy is built redundantly on every rank from the global stencil by the deterministic device set-up. The replicated part is
// ~20% of a V-cycle's work, which bounds the speed-up (Amdahl) -- distributing level 1 as well is the next step.
struct NullOp : LinOp<cdouble> {
    void apply(const cdouble*, cdouble*, PcgState*, ReduceWs, cudaStream_t, bool) override { throw std::logic_error("NullOp applied"); }
    bool applies() const override { return false; }
};

// ---- level 1 of the global hierarchy, distributed (round 2) ----------------------------------------------------------------------------
// The replicated coarse levels were the Amdahl limit of the partitioned solve (level 1 of the 8M-vertex hierarchy: 3.1 M rows, 33 M entries,
// four fused smoother SpMVs per cycle on EVERY rank). Level 1 is now row-partitioned as well: rank r owns the contiguous index block
// [r c, (r+1) c), c = ceil(n1 / W) (level-1 dofs are numbered in root-face order, so blocks are spatially compact), keeps its rows of A_1
// with LOCAL column numbers [owned | halo] -- the same x_ext layout as level 0, so every fused smoother kernel of csr.cuh runs unchanged --
// and exchanges the few thousand halo entries before each of the four SpMVs (NCCL send/recv, plan derived locally from the replicated
// operator: every rank sees all rows, no set-up communication). The restricted residual arrives by reduce-scatter instead of all-reduce, the
// correction leaves by all-gather (together the bytes of the all-reduce they replace); level 2 (1/13 of level 1) and below stay replicated.
static __global__ void k_l1_mark_halo(int nown, int lo, int hi, const int* __restrict__ ptr, const int* __restrict__ col, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nown) return;
    for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) {
        const int c = col[k];
        if (c < lo || c >= hi) flag[c] = 1;  // racing writes of the same value
    }
}
static __global__ void k_l1_renumber(int nown, int lo, int hi, const int* __restrict__ ptr, int* __restrict__ col, const int* __restrict__ pos) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nown) return;
    for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) {
        const int c = col[k];
        col[k] = (c >= lo && c < hi) ? c - lo : nown + pos[c];
    }
}
static __global__ void k_l1_compact(int n, const int* __restrict__ flag, const int* __restrict__ pos, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = i;
}
// need[(owner of row) * nown + (c - lo)] = 1 for every entry (row, c) of the GLOBAL operator with c in my block and the row in another
static __global__ void k_l1_mark_needed(int n1, int chunk, int me, int lo, int nown, const int* __restrict__ ptr, const int* __restrict__ col,
                                        unsigned char* __restrict__ need) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n1) return;
    const int ro = i / chunk;
    if (ro == me) return;
    for (int k = ptr[i] + lane; k < ptr[i + 1]; k += 32) {
        const int c = col[k];
        if (c >= lo && c < lo + nown) need[(size_t)ro * nown + (c - lo)] = 1;
    }
}
static __global__ void __launch_bounds__(kThreads) k_l1_pack(int m, const int* __restrict__ idx, const cfloat* __restrict__ x, cfloat* __restrict__ out,
                                                      const PcgState* st) {
    if (st && st->done) return;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) out[k] = x[__ldg(idx + k)];
}
static __global__ void __launch_bounds__(kThreads) k_l1_copy(int n, const cfloat* __restrict__ in, cfloat* __restrict__ out, const PcgState* st) {
    if (st && st->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i];
}

struct DistLevel1 {
    DistComm* dc = nullptr;
    int n1 = 0, n2 = 0, chunk = 0, lo = 0, hi = 0, nown = 0, nhalo = 0, nsend = 0;
    DeviceCsr<cfloat> A, P, R;   // own rows of A_1 (local columns), own rows of P_1 (n_own x n2), its conjugate transpose (n2 x n_own)
    DevBuf<cfloat> dinv, b, r, x_ext, d_ext, q_ext, xo, b2, x2, sendbuf;
    DevBuf<int> send_idx;
    std::vector<int> send_counts, recv_counts, send_off, recv_off;
    double lmax = 2.0;

    // builds the slice of rank dc->rank from the replicated fp32 level (its A, P, dinv, lmax); `lv` is level 0 of the fp32 tail hierarchy
    void build(DistComm* dc_, AmgLevelDev<cfloat>& lv, int n2_, cudaStream_t s) {
        dc = dc_;
        const int W = dc->world, me = dc->rank;
        n1 = lv.n; n2 = n2_;
        chunk = (n1 + W - 1) / W;
        lo = std::min(me * chunk, n1);
        hi = std::min(lo + chunk, n1);
        nown = hi - lo;
        lmax = lv.lmax;
        DevBuf<int> rows(std::max(nown, 1));
        {
            std::vector<int> h(std::max(nown, 1));
            for (int i = 0; i < nown; i++) h[i] = lo + i;
            rows.upload(h.data(), nown, s);
        }
        csr_gather_rows<cfloat>(lv.A, rows.p, nown, A, s);
        csr_gather_rows<cfloat>(lv.P, rows.p, nown, P, s);
        transpose_dev<cfloat>(P, R, true, s);
        // halo columns of my rows, local numbering
        DevBuf<int> flag((size_t)n1 + 1), pos((size_t)n1 + 1);
        flag.zero(s);
        const int gw = std::max(1, (nown * 32 + 255) / 256);
        k_l1_mark_halo<<<gw, 256, 0, s>>>(nown, lo, hi, A.ptr.p, A.col.p, flag.p);
        dev_exclusive_scan(flag.p, pos.p, n1, s);
        nhalo = dev_read_int(pos.p + n1, s);
        DevBuf<int> halo_glob(std::max(nhalo, 1));
        k_l1_compact<<<(n1 + 255) / 256, 256, 0, s>>>(n1, flag.p, pos.p, halo_glob.p);
        k_l1_renumber<<<gw, 256, 0, s>>>(nown, lo, hi, A.ptr.p, A.col.p, pos.p);
        MK_CHECK_LAUNCH();
        A.cols = nown + nhalo;
        std::vector<int> hg(std::max(nhalo, 1));
        halo_glob.download(hg.data(), nhalo, s);
        // what the peers need of my block: the same rule applied to THEIR rows of the replicated operator
        DevBuf<unsigned char> need((size_t)W * std::max(nown, 1));
        MK_CUDA(cudaMemsetAsync(need.p, 0, (size_t)W * std::max(nown, 1), s));
        k_l1_mark_needed<<<(int)(((int64_t)n1 * 32 + 255) / 256), 256, 0, s>>>(n1, chunk, me, lo, nown, lv.A.ptr.p, lv.A.col.p, need.p);
        MK_CHECK_LAUNCH();
        std::vector<unsigned char> hn((size_t)W * std::max(nown, 1));
        need.download(hn.data(), hn.size(), s);
        MK_CUDA(cudaStreamSynchronize(s));
        send_counts.assign(W, 0); recv_counts.assign(W, 0); send_off.assign(W, 0); recv_off.assign(W, 0);
        for (int k = 0; k < nhalo; k++) recv_counts[std::min(hg[k] / chunk, W - 1)]++;   // sorted by global index: contiguous per owner
        std::vector<int> sidx;
        for (int p = 0; p < W; p++) {
            send_off[p] = (int)sidx.size();
            if (p != me)
                for (int j = 0; j < nown; j++)
                    if (hn[(size_t)p * nown + j]) sidx.push_back(j);   // ascending local index = ascending global index = the peer's halo order
            send_counts[p] = (int)sidx.size() - send_off[p];
        }
        for (int p = 1; p < W; p++) recv_off[p] = recv_off[p - 1] + recv_counts[p - 1];
        nsend = (int)sidx.size();
        send_idx.alloc(std::max(nsend, 1));
        send_idx.upload(sidx.data(), nsend, s);
        sendbuf.alloc(std::max(nsend, 1));
        build_sell(A, s);
        build_sell(P, s);
        build_sell(R, s);
        dinv.alloc(std::max(nown, 1));
        if (nown) MK_CUDA(cudaMemcpyAsync(dinv.p, lv.dinv.p + lo, sizeof(cfloat) * nown, cudaMemcpyDeviceToDevice, s));
        const size_t ne = (size_t)std::max(nown + nhalo, 1);
        b.alloc(std::max(chunk, 1)); r.alloc(std::max(nown, 1)); xo.alloc(std::max(chunk, 1));
        x_ext.alloc(ne); d_ext.alloc(ne); q_ext.alloc(ne);
        x_ext.zero(s); d_ext.zero(s); q_ext.zero(s); xo.zero(s); b.zero(s);
        b2.alloc(std::max(n2, 1)); x2.alloc(std::max(n2, 1));
        MK_CUDA(cudaStreamSynchronize(s));
        if (getenv("MK_VERBOSE"))
            fprintf(stderr, "[mk] distributed level 1: rank %d owns %d of %d rows, halo %d, sends %d\n", me, nown, n1, nhalo, nsend);
    }
    // refresh the halo part of v_ext = [owned | halo] from the owners
    void exchange(cfloat* v_ext, const PcgState* st, cudaStream_t s) {
        NcclApi& nc = NcclApi::get();
        if (nsend) k_l1_pack<<<stream_grid(nsend), kThreads, 0, s>>>(nsend, send_idx.p, v_ext, sendbuf.p, st);
        MK_NCCL(nc.GroupStart());
        for (int p = 0; p < dc->world; p++) {
            if (p == dc->rank) continue;
            if (send_counts[p]) MK_NCCL(nc.Send(sendbuf.p + send_off[p], 2 * (size_t)send_counts[p], ncclFloat, p, dc->comm, s));
            if (recv_counts[p]) MK_NCCL(nc.Recv(v_ext + nown + recv_off[p], 2 * (size_t)recv_counts[p], ncclFloat, p, dc->comm, s));
        }
        MK_NCCL(nc.GroupEnd());
    }
    // b1_pad: every rank's partial sums of the restricted residual, W * chunk entries (zero padded); x1_pad: the correction, same length.
    // `tail` = the replicated fp32 hierarchy whose level 0 is this level (its levels >= 1 are used as they are).
    void cycle(const cfloat* b1_pad, cfloat* x1_pad, AmgPrecond<cfloat>& tail, int deg, double ratio, const PcgState* st, cudaStream_t s) {
        NcclApi& nc = NcclApi::get();
        MK_NCCL(nc.ReduceScatter(b1_pad, b.p, 2 * (size_t)chunk, ncclFloat, ncclSum, dc->comm, s));
        const double lmin = lmax / ratio, theta = 0.5 * (lmax + lmin), de = 0.5 * (lmax - lmin), sigma = theta / de;
        const int g = stream_grid(std::max(nown, 1));
        cfloat* d_cur = deg == 1 ? x_ext.p : d_ext.p;
        if (nown) k_start_dir<cfloat><<<g, kThreads, 0, s>>>(nown, b.p, dinv.p, d_cur, 1.0 / theta, st);
        {
            double rho = 1.0 / sigma;
            for (int k = 0; k < deg; k++) {
                const double rho1 = 1.0 / (2.0 * sigma - rho);
                cfloat* d_next = (k == deg - 1) ? nullptr : (d_cur == d_ext.p ? q_ext.p : d_ext.p);
                exchange(d_cur, st, s);
                if (nown) spmv_cheb_step<cfloat>(A, 0.0, d_cur, d_next, r.p, x_ext.p, b.p, dinv.p, rho1 * rho, 2.0 * rho1 / de, k == 0 ? 1 : 0, st, s);
                d_cur = d_next;
                rho = rho1;
            }
        }
        // restriction: partial sums over my columns of R_1, summed over the ranks; levels >= 2 replicated
        if (nown) spmv_plain(R, r.p, b2.p, false, st, s);
        else MK_CUDA(cudaMemsetAsync(b2.p, 0, sizeof(cfloat) * n2, s));
        dc->allreduce_sum_f32(reinterpret_cast<float*>(b2.p), 2 * n2, s);
        tail.vcycle(1, b2.p, x2.p, st, s);
        if (nown) spmv_plain(P, x2.p, x_ext.p, true, st, s);
        // post-smoothing
        exchange(x_ext.p, st, s);
        if (deg == 1) {
            if (nown) spmv_cheb_resid_last<cfloat>(A, 0.0, x_ext.p, b.p, xo.p, dinv.p, 1.0 / theta, nullptr, const_cast<PcgState*>(st), ReduceWs{nullptr, nullptr}, s);
        } else {
            if (nown) spmv_cheb_resid<cfloat>(A, 0.0, x_ext.p, b.p, r.p, d_ext.p, dinv.p, 1.0 / theta, st, s);
            cfloat* dc_ = d_ext.p;
            double rho = 1.0 / sigma;
            for (int k = 0; k + 1 < deg; k++) {
                const double rho1 = 1.0 / (2.0 * sigma - rho);
                exchange(dc_, st, s);
                if (k + 2 == deg) {
                    if (nown) spmv_cheb_step_last<cfloat>(A, 0.0, dc_, r.p, x_ext.p, xo.p, dinv.p, rho1 * rho, 2.0 * rho1 / de, nullptr, const_cast<PcgState*>(st),
                                                          ReduceWs{nullptr, nullptr}, s);
                } else {
                    cfloat* dn = dc_ == d_ext.p ? q_ext.p : d_ext.p;
                    if (nown) spmv_cheb_step<cfloat>(A, 0.0, dc_, dn, r.p, x_ext.p, nullptr, dinv.p, rho1 * rho, 2.0 * rho1 / de, 0, st, s);
                    dc_ = dn;
                }
                rho = rho1;
            }
        }
        MK_CHECK_LAUNCH();
        MK_NCCL(nc.AllGather(xo.p, x1_pad, 2 * (size_t)chunk, ncclFloat, dc->comm, s));
    }
};

// hand-over of the restricted residual to the fp32 replicated tail: b_lo = b / |r| (|r| = the PCG residual norm in the state: any factor of the
// right magnitude serves, the cycle is linear), and back: xo = |r| xo_lo. st == null (the start of a solve): no scaling.
static __global__ void __launch_bounds__(kThreads) k_tail_narrow(int n, const cdouble* __restrict__ b, cfloat* __restrict__ b_lo, const PcgState* st) {
    if (st && st->done) return;
    const double sc = (st && st->rr > 0.0) ? rsqrt(st->rr) : 1.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) b_lo[i] = to_low<cfloat, cdouble>(ldc_nc(b + i), sc);
}
static __global__ void __launch_bounds__(kThreads) k_tail_widen(int n, const cfloat* __restrict__ x_lo, cdouble* __restrict__ x, const PcgState* st) {
    if (st && st->done) return;
    const double sc = (st && st->rr > 0.0) ? sqrt(st->rr) : 1.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) stc(x + i, to_high<cdouble, cfloat>(x_lo[i], sc));
}

struct DistAmgPrecond : Precond<cdouble> {
    DistFaceOp* A = nullptr;
    DistComm* dc = nullptr;
    NullOp null_op;
    std::unique_ptr<AmgPrecond<cdouble>> G;  // global hierarchy; level 0 only contributes lmax / dinv / P_0
    std::unique_ptr<AmgPrecond<cfloat>> Glo;  // fp32 copy of its levels >= 1 (MK_PRECOND_FP32, default): the replicated part of every rank
    DevBuf<cfloat> b1_lo, x1_lo;              // level-1 right-hand side (summed over the ranks in fp32: half the all-reduce bytes) and result
    std::unique_ptr<DistLevel1> L1;           // world > 1: level 1 row-partitioned as well (MK_DIST_L1=0: replicated like the levels below it)
    DeviceCsr<cdouble> P0, R0;               // this rank's rows of P_0 (local row order) and their conjugate transpose (n1 x n_local)
    DevBuf<cdouble> dinv, r, d_ext, q;
    double lmax = 2.0;
    int nloc = 0, next = 0, n1 = 0;
    AmgParams prm;

    bool ready() const { return G && G->lv.size() >= 2; }

    void smooth(const cdouble* b, cdouble* x_ext, bool zero_guess, const PcgState* st, cudaStream_t s) {
        const int grid = stream_grid(nloc);
        const double lmin = lmax / prm.cheb_ratio;
        const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
        double rho = 1.0 / sigma;
        if (zero_guess) {
            k_cheb_init<cdouble><<<grid, kThreads, 0, s>>>(nloc, b, (const cdouble*)nullptr, dinv.p, r.p, d_ext.p, 1.0 / theta, st);
        } else {
            A->apply(x_ext, q.p, st, s);
            k_cheb_init<cdouble><<<grid, kThreads, 0, s>>>(nloc, b, q.p, dinv.p, r.p, d_ext.p, 1.0 / theta, st);
        }
        for (int k = 0; k < prm.cheb_degree; k++) {
            A->apply(d_ext.p, q.p, st, s);
            const double rho1 = 1.0 / (2.0 * sigma - rho);
            k_cheb_step<cdouble><<<grid, kThreads, 0, s>>>(nloc, q.p, dinv.p, r.p, x_ext, d_ext.p, rho1 * rho, 2.0 * rho1 / delta,
                                                          (zero_guess && k == 0) ? 1 : 0, st);
            rho = rho1;
        }
        MK_CHECK_LAUNCH();
    }

    // z must have room for the halo slots ([owned | halo] layout): it is the level-0 iterate the partitioned operator is applied to
    void apply(const cdouble* rin, cdouble* z, const PcgState* st, cudaStream_t s) override {
        smooth(rin, z, true, st, s);
        AmgLevelDev<cdouble>& C = G->lv[1];
        spmv_plain(R0, r.p, C.b.p, false, st, s);
        if (Glo) {
            k_tail_narrow<<<stream_grid(n1), kThreads, 0, s>>>(n1, C.b.p, b1_lo.p, st);
            if (L1) {
                L1->cycle(b1_lo.p, x1_lo.p, *Glo, Glo->degree(0), Glo->prm.cheb_ratio, st, s);
            } else {
                dc->allreduce_sum_f32(reinterpret_cast<float*>(b1_lo.p), 2 * n1, s);
                Glo->vcycle(0, b1_lo.p, x1_lo.p, st, s);
            }
            k_tail_widen<<<stream_grid(n1), kThreads, 0, s>>>(n1, x1_lo.p, C.xo.p, st);
            MK_CHECK_LAUNCH();
        } else {
            dc->allreduce_sum(reinterpret_cast<double*>(C.b.p), 2 * n1, s);
            G->vcycle(1, C.b.p, C.xo.p, st, s);
        }
        spmv_plain(P0, C.xo.p, z, true, st, s);
        smooth(rin, z, false, st, s);
    }
    int launches_per_apply() const override {
        const int per_spmv = A->launches_per_apply();
        return (2 * prm.cheb_degree + 1) * (per_spmv + 1) + 1 + 2 + (Glo ? 3 + Glo->launches_from_level(0) : (G ? 1 + G->launches_from_level(1) : 0));
    }
};

// Distributed PCG: A = DistFaceOp, M = the global-hierarchy multigrid above (or the local V-cycle on the diagonal block, or identity).
// Two all-reduces per iteration (<p,Ap>; {<r,r>, <r,z>}); the host looks at the state every `chunk` iterations.
struct DistPcg {
    int n = 0, next = 0;
    cudaStream_t s = nullptr;
    DistComm* dc = nullptr;
    DevBuf<cdouble> r, z, p_ext, q;
    DevBuf<PcgState> st;
    DevBuf<double> part, partials;
    DevBuf<unsigned int> ticket;
    PcgState* h_st = nullptr;
    int grid = 1;

    DistPcg(int n_, int next_, DistComm* dc_, cudaStream_t s_) : n(n_), next(next_), s(s_), dc(dc_), r(n_), z(next_), p_ext(next_), q(n_), st(1), part(8),
                                                                  partials(kMaxPartials * 4), ticket(1) {
        ticket.zero(s);
        st.zero(s);
        part.zero(s);
        p_ext.zero(s);
        MK_CUDA(cudaMallocHost((void**)&h_st, sizeof(PcgState)));
        grid = stream_grid(n);
    }
    ~DistPcg() { if (h_st) cudaFreeHost(h_st); }
    ReduceWs ws() { return ReduceWs{partials.p, ticket.p}; }
    void fetch() {
        MK_CUDA(cudaMemcpyAsync(h_st, st.p, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
        MK_CUDA(cudaStreamSynchronize(s));
    }
    void precond(Precond<cdouble>* M, const cdouble* rr, cdouble* zz, bool unconditional = false) {
        if (M) M->apply(rr, zz, unconditional ? nullptr : st.p, s);
    }
    // part = {<r,r>, <b,b>, <r,z>} -> state; p = z
    void start(DistFaceOp& A, Precond<cdouble>* M, const cdouble* b, double tol, int maxit, int keep) {
        precond(M, r.p, z.p, true);  // the state still carries the previous solve's `done` flag here
        const cdouble* zz = M ? z.p : r.p;
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, r.p, r.p, b, b, nullptr, part.p, ws(), nullptr);
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, r.p, zz, nullptr, nullptr, nullptr, part.p + 2, ws(), nullptr);
        dc->allreduce_sum(part.p, 3, s);
        k_dist_start<<<1, 1, 0, s>>>(st.p, part.p, tol * tol, maxit, keep);
        MK_CUDA(cudaMemcpyAsync(p_ext.p, zz, sizeof(cdouble) * n, cudaMemcpyDeviceToDevice, s));
        MK_CHECK_LAUNCH();
    }
    void iteration(DistFaceOp& A, Precond<cdouble>* M, cdouble* x) {
        A.apply(p_ext.p, q.p, st.p, s);
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, p_ext.p, q.p, nullptr, nullptr, nullptr, part.p, ws(), st.p);
        dc->allreduce_sum(part.p, 1, s);
        k_dist_update_xr<<<grid, kThreads, 0, s>>>(n, x, r.p, p_ext.p, q.p, st.p, part.p, ws());
        precond(M, r.p, z.p);
        const cdouble* zz = M ? z.p : r.p;
        k_dist_dots<<<grid, kThreads, 0, s>>>(n, r.p, zz, nullptr, nullptr, nullptr, part.p + 2, ws(), st.p);
        dc->allreduce_sum(part.p + 1, 2, s);
        k_dist_update_p<<<grid, kThreads, 0, s>>>(n, p_ext.p, zz, st.p, part.p);
        k_dist_state<<<1, 1, 0, s>>>(st.p, part.p);
        MK_CHECK_LAUNCH();
    }
    // solves L x = b from the x passed in (x_ext layout: halo part is scratch). Returns stats like Pcg::solve.
    SolveStats solve(DistFaceOp& A, Precond<cdouble>* M, const cdouble* b, cdouble* x_ext, double tol, int maxit, int chunk = 8, int hint = 0) {
        SolveStats out;
        long launches = 0;
        A.apply(x_ext, q.p, nullptr, s);
        k_dist_residual<<<grid, kThreads, 0, s>>>(n, b, q.p, r.p);
        int keep = 0;
        bool breakdown = false;
        double prev_relres = 1e300;
        for (int attempt = 0; attempt < 4; attempt++) {
            start(A, M, b, tol, maxit, keep);
            keep = 1;
            int batch = hint > 0 ? std::min(std::max(hint, 1), 4 * chunk) : chunk;
            while (true) {
                for (int c = 0; c < batch; c++) iteration(A, M, x_ext);
                launches += (long)batch * (8 + (M ? M->launches_per_apply() : 0));
                fetch();
                if (h_st->done) break;
                batch = chunk;
            }
            breakdown = h_st->done == 2;
            // true residual
            A.apply(x_ext, q.p, nullptr, s);
            k_dist_residual<<<grid, kThreads, 0, s>>>(n, b, q.p, r.p);
            k_dist_dots<<<grid, kThreads, 0, s>>>(n, r.p, r.p, nullptr, nullptr, nullptr, part.p + 4, ws(), nullptr);
            dc->allreduce_sum(part.p + 4, 1, s);
            double rr = 0;
            MK_CUDA(cudaMemcpyAsync(&rr, part.p + 4, sizeof(double), cudaMemcpyDeviceToHost, s));
            fetch();
            out.iterations = h_st->iter;
            out.relres = h_st->bnorm2 > 0 ? std::sqrt(rr / h_st->bnorm2) : 0.0;
            out.restarts = attempt;
            launches += 6;
            if (out.relres <= tol * 1.0000001 || h_st->iter >= maxit || breakdown) break;
            if (attempt > 0 && out.relres > 0.5 * prev_relres) break;  // fp64 floor reached (see Pcg::solve)
            prev_relres = out.relres;
        }
        out.status = (out.relres <= tol * 1.0000001 || h_st->bnorm2 == 0) ? 0 : (breakdown ? 2 : 1);
        out.launches = launches;
        return out;
    }
};

}  // namespace mk
