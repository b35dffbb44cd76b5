// Preconditioned conjugate gradients with every scalar device-resident (no host sync inside an iteration),
// fused vector kernels with warp-shuffle reductions, and the iteration body replayed as a CUDA graph.
// Replaces the direct factorisations of the reference (Eigen::SparseLU sls.h:42-47, CHOLMOD LDL^T
// ss/suitesparse_pdefinite.h:26-41, Eigen::SimplicialLLT ir.h:14-20) with an HBM-bound iterative path.
#pragma once
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace mk {

// Device memory comes from a per-context arena: a few large slabs obtained from the driver once and sub-allocated on the host
// (best fit, coalescing free list). A solve step allocates and frees gigabytes of operators, hierarchies and work vectors; driver
// allocations -- cudaMalloc / cudaFree as well as growth or trimming of the stream-ordered pool -- cost 0.06-0.5 s per GB on this
// platform (profiles/alloc_probe_r01.txt) and showed up as sporadic 0.3-0.9 s outliers inside timed steps. With the arena a step makes
// no driver allocation call at all once the slabs exist. Every context issues all its work on ONE stream, so handing a freed block to
// the next request is ordered correctly without events; one host thread per context (include/metriko_b200.h).
// Buffers created outside any context scope fall back to cudaMallocAsync on the scope's stream.
struct Arena {
    static constexpr size_t kAlign = 512;
    // Slab sizes grow geometrically from 256 MiB to 2 GiB (a 2,562-vertex mesh should not pin 2 GiB; a 1M-vertex step ends up with a handful
    // of 2 GiB slabs after its first pass); MK_ARENA_SLAB_MB fixes the size instead. A request larger than the current slab size gets a
    // slab of its own.
    static constexpr size_t kSlabFirst = (size_t)256 << 20, kSlabMax = (size_t)2 << 30;
    size_t next_slab = slab_from_env() ? slab_from_env() : kSlabFirst;
    static size_t slab_from_env() {
        static const size_t v = [] { const char* e = getenv("MK_ARENA_SLAB_MB"); return e && atol(e) > 0 ? (size_t)atol(e) << 20 : (size_t)0; }();
        return v;
    }
    std::vector<void*> slabs;
    std::map<char*, size_t> free_by_addr;
    std::multimap<size_t, char*> free_by_size;
    std::unordered_map<void*, size_t> live;
    size_t reserved = 0, in_use = 0, peak = 0;
    Arena() = default;
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    ~Arena() {
        for (void* p : slabs) cudaFree(p);
    }
    bool is_slab_base(const char* p) const {
        for (void* b : slabs)
            if ((const char*)b == p) return true;
        return false;
    }
    void erase_free(char* p, size_t sz) {
        free_by_addr.erase(p);
        auto r = free_by_size.equal_range(sz);
        for (auto it = r.first; it != r.second; ++it)
            if (it->second == p) { free_by_size.erase(it); break; }
    }
    void insert_free(char* p, size_t sz) {
        // coalesce with the neighbours -- never across a slab boundary: a block must lie inside ONE driver allocation
        // (cudaMemcpy / cudaMemset reject ranges that span two, even when they are adjacent in the address space)
        auto nx = free_by_addr.find(p + sz);
        if (nx != free_by_addr.end() && !is_slab_base(p + sz)) { const size_t s2 = nx->second; erase_free(nx->first, s2); sz += s2; }
        auto it = free_by_addr.lower_bound(p);
        if (it != free_by_addr.begin() && !is_slab_base(p)) {
            --it;
            if (it->first + it->second == p) { char* q = it->first; const size_t s0 = it->second; erase_free(q, s0); p = q; sz += s0; }
        }
        free_by_addr[p] = sz;
        free_by_size.emplace(sz, p);
    }
    void* alloc(size_t bytes) {
        bytes = (bytes + kAlign - 1) / kAlign * kAlign;
        auto it = free_by_size.lower_bound(bytes);
        if (it == free_by_size.end()) {
            const size_t slab = bytes > next_slab ? bytes : next_slab;
            if (!slab_from_env() && next_slab < kSlabMax) next_slab *= 2;
            void* base = nullptr;
            MK_CUDA(cudaMalloc(&base, slab));
            slabs.push_back(base);
            reserved += slab;
            insert_free((char*)base, slab);
            it = free_by_size.lower_bound(bytes);
        }
        char* p = it->second;
        const size_t sz = it->first;
        erase_free(p, sz);
        if (sz > bytes) insert_free(p + bytes, sz - bytes);
        live[p] = bytes;
        in_use += bytes;
        if (in_use > peak) peak = in_use;
        return p;
    }
    void release(void* p) {
        auto it = live.find(p);
        if (it == live.end()) return;
        const size_t sz = it->second;
        live.erase(it);
        in_use -= sz;
        insert_free((char*)p, sz);
    }
};

inline cudaStream_t& alloc_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}
inline Arena*& alloc_arena() {
    static thread_local Arena* a = nullptr;
    return a;
}
struct AllocStreamScope {
    cudaStream_t prev;
    Arena* prev_arena;
    explicit AllocStreamScope(cudaStream_t s, Arena* a = nullptr) : prev(alloc_stream()), prev_arena(alloc_arena()) { alloc_stream() = s; alloc_arena() = a; }
    ~AllocStreamScope() { alloc_stream() = prev; alloc_arena() = prev_arena; }
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    Arena* owner = nullptr;  // the arena the block came from (null: cudaMallocAsync); handles keep their arena alive
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owner(o.owner) { o.p = nullptr; o.n = 0; o.owner = nullptr; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; owner = o.owner; o.p = nullptr; o.n = 0; o.owner = nullptr; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (owner) owner->release(p);
            else cudaFreeAsync(p, alloc_stream());
        }
        p = nullptr; n = 0; owner = nullptr;
    }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (!n) return;
        owner = alloc_arena();
        if (owner) p = (T*)owner->alloc(n * sizeof(T));
        else MK_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), alloc_stream()));
    }
    void zero(cudaStream_t s) { if (n) MK_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void upload(const T* h, size_t cnt, cudaStream_t s) { if (cnt) MK_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s)); }
    void download(T* h, size_t cnt, cudaStream_t s) const { if (cnt) MK_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, s)); }
};

// Device-resident solver state. One instance per PCG object; kernels read alpha/beta from here.
struct PcgState {
    double rho;       // <r, z>
    double pq;        // <p, A p>
    double rr;        // <r, r>
    double alpha, beta;
    double bnorm2;    // <b, b>
    double thresh;    // tol^2 * bnorm2
    double aux0, aux1, aux2;  // scratch reductions for drivers (e.g. <u, M u>)
    double zscale;    // mixed-precision hand-over: the fp32 cycle ran on r / zscale, its output is z / zscale (set by the hand-over kernel)
    int iter;
    int done;         // 1 = converged (kernels become no-ops), 2 = breakdown
    int maxit;
    int pad;
};

struct PcgState;
// Pinned host copies of the solver state: cudaMallocHost / cudaFreeHost cost 3-5 ms each on this platform and a step creates several
// solver objects, so the slots are recycled (per host thread; never returned to the driver).
struct PinnedStatePool {
    std::vector<void*> free_list;
    void* get(size_t bytes) {
        if (!free_list.empty()) { void* p = free_list.back(); free_list.pop_back(); return p; }
        void* p = nullptr;
        MK_CUDA(cudaMallocHost(&p, bytes < 256 ? 256 : bytes));
        return p;
    }
    void put(void* p) { if (p) free_list.push_back(p); }
    static PinnedStatePool& instance() { static thread_local PinnedStatePool pool; return pool; }
};

struct ReduceWs {
    double* partials;      // kMaxPartials * 4 doubles
    unsigned int* ticket;  // zero-initialised, reset by the last CTA
};

// ---------------------------------------------------------------------------------------------
// fused vector kernels (grid-stride, one pass over each vector, reductions finished in-kernel)
// ---------------------------------------------------------------------------------------------

// x += alpha p ; r -= alpha q ; rr = <r,r>.  Without a preconditioner z == r, so rho/beta are finished here too.
template <class T, bool PRECOND>
__global__ void __launch_bounds__(kThreads) k_update_xr(int n, T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p,
                                                        const T* __restrict__ q, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    const double alpha = st->alpha;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T pi = ld_ro(p + i), qi = ld_ro(q + i);
        T xi = ld_rw(x + i), ri = ld_rw(r + i);
        xi = xi + ST<T>::scale(alpha, pi);
        ri = ri - ST<T>::scale(alpha, qi);
        st_rw(x + i, xi);
        st_rw(r + i, ri);
        acc += ST<T>::abs2(ri);
    }
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
        st->rr = s[0];
        st->iter += 1;
        if (!PRECOND) {
            st->beta = s[0] / st->rho;
            st->rho = s[0];
        }
        if (s[0] <= st->thresh) st->done = 1;
        else if (st->iter >= st->maxit) st->done = 3;
        else if (!(s[0] == s[0])) st->done = 2;
    });
}

// rho_new = <r, z> ; beta = rho_new / rho
template <class T>
__global__ void __launch_bounds__(kThreads) k_dot_rz(int n, const T* __restrict__ r, const T* __restrict__ z, PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += ST<T>::hdot(ld_ro(r + i), ld_ro(z + i));
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) {
        st->beta = s[0] / st->rho;
        st->rho = s[0];
    });
}

// p = z + beta p
template <class T>
__global__ void __launch_bounds__(kThreads) k_update_p(int n, T* __restrict__ p, const T* __restrict__ z, const PcgState* st) {
    if (st->done) return;
    const double beta = st->beta;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T zi = ld_ro(z + i), pi = ld_rw(p + i);
        st_rw(p + i, zi + ST<T>::scale(beta, pi));
    }
}

// r = b - q (q = A x) ; rr = <r,r> ; bnorm2 = <b,b> ; resets the iteration state. If keep_bnorm, only rr is refreshed
// (used for the true-residual restart).
template <class T>
__global__ void __launch_bounds__(kThreads) k_residual(int n, const T* __restrict__ b, const T* __restrict__ q, T* __restrict__ r,
                                                       PcgState* st, ReduceWs ws, double tol2, int maxit, int keep_bnorm) {
    __shared__ double red[32];
    double a0 = 0.0, a1 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T bi = ld_ro(b + i), qi = ld_ro(q + i);
        T ri = bi - qi;
        st_rw(r + i, ri);
        a0 += ST<T>::abs2(ri);
        a1 += ST<T>::abs2(bi);
    }
    double v[2] = {a0, a1};
    grid_reduce<2>(v, ws.partials, ws.ticket, red, [&](const double* s) {
        st->rr = s[0];
        if (!keep_bnorm) {
            st->bnorm2 = s[1];
            st->thresh = tol2 * s[1];
            st->iter = 0;
        }
        st->maxit = maxit;
        st->done = (s[0] <= st->thresh) ? 1 : 0;
    });
}

// p = z ; rho = <r, z>   (start of a PCG run; with no preconditioner call with z == r)
template <class T>
__global__ void __launch_bounds__(kThreads) k_start_p(int n, const T* __restrict__ r, const T* __restrict__ z, T* __restrict__ p,
                                                      PcgState* st, ReduceWs ws) {
    __shared__ double red[32];
    if (st->done) return;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T zi = ld_ro(z + i);
        st_rw(p + i, zi);
        acc += ST<T>::hdot(ld_ro(r + i), zi);
    }
    double v[1] = {acc};
    grid_reduce<1>(v, ws.partials, ws.ticket, red, [&](const double* s) { st->rho = s[0]; });
}

// ---------------------------------------------------------------------------------------------
// operator / preconditioner / projector interfaces
// ---------------------------------------------------------------------------------------------
template <class T>
struct LinOp {
    int n = 0;
    virtual ~LinOp() = default;
    // y = A x. st != nullptr: the launch is skipped once st->done is set (graph replay); if additionally `dot`, the
    // kernel finishes the fused reduction pq = <x, y>, alpha = rho / pq (the PCG step).
    virtual void apply(const T* x, T* y, PcgState* st, ReduceWs ws, cudaStream_t s, bool dot = true) = 0;
    virtual int launches_per_apply() const { return 1; }
    virtual bool applies() const { return true; }  // false: a placeholder that only marks "level 0 lives elsewhere" (dist.cuh NullOp)
    // Chebyshev smoothing fused into the operator kernel (the multigrid V-cycle, amg.cuh; the epilogues are documented at SpmvMode, csr.cuh):
    //   cheb_step      : q = A d_in ; x (+)= d_in ; r -= q ; d_out = c1 d_in + c2 dinv .* r   (`first`: x = d_in, r = b - q; d_out may be null)
    //   cheb_resid     : r = b - A x ; d_out = inv_theta dinv .* r
    //   cheb_resid_last: xo = x + inv_theta dinv .* (b - A x)                        (degree-1 post-smoothing in one SpMV)
    //   cheb_step_last : xo = x + d_in + (c1 d_in + c2 dinv .* (r - A d_in))          (last post-smoothing step of degree >= 2)
    // The two *_last forms take the fp64 PCG residual r_hi (HiOf<T>::type*, may be null): with it they also finish the PCG reduction
    // rho_new = zscale <r_hi, xo>, beta = rho_new / rho in the same kernel.
    virtual void cheb_step(const T*, T*, T*, T*, const T*, const T*, double, double, int, const PcgState*, cudaStream_t) { throw std::logic_error("cheb_step"); }
    virtual void cheb_resid(const T*, const T*, T*, T*, const T*, double, const PcgState*, cudaStream_t) { throw std::logic_error("cheb_resid"); }
    virtual void cheb_resid_last(const T*, const T*, T*, const T*, double, const void*, PcgState*, ReduceWs, cudaStream_t) { throw std::logic_error("cheb_resid_last"); }
    virtual void cheb_step_last(const T*, T*, T*, T*, const T*, double, double, const void*, PcgState*, ReduceWs, cudaStream_t) { throw std::logic_error("cheb_step_last"); }
};

template <class T>
struct Precond {
    virtual ~Precond() = default;
    virtual void apply(const T* r, T* z, const PcgState* st, cudaStream_t s) = 0;  // z = M^{-1} r
    virtual int launches_per_apply() const = 0;
    // z = M^{-1} r and, fused into its last kernel, the PCG reduction rho_new = <r, z>, beta = rho_new / rho (what k_dot_rz does).
    // Return false when unsupported (the caller then runs apply + k_dot_rz).
    virtual bool apply_dot(const T*, T*, PcgState*, ReduceWs, cudaStream_t) { return false; }
    // Fused PCG body (solves without a projector): the preconditioner takes over the vector updates either side of its cycle, so that
    // an iteration is  q = A p (+ <p,q>)  ->  update_xr_handover  ->  cycle_dot  ->  update_p  with no stand-alone vector kernel.
    virtual bool fused_pcg() const { return false; }
    //   x += alpha p ; r -= alpha q ; rr = <r,r> (+ the iteration bookkeeping of k_update_xr) and the hand-over of r to the cycle
    virtual void update_xr_handover(int, T*, T*, const T*, const T*, PcgState*, ReduceWs, cudaStream_t) {}
    //   the cycle on the handed-over residual; its last kernel finishes rho_new = <r, z>, beta = rho_new / rho
    virtual void cycle_dot(const T*, PcgState*, ReduceWs, cudaStream_t) {}
    //   p = z + beta p, z read from the cycle's own output
    virtual void update_p(int, T*, const PcgState*, cudaStream_t) {}
    virtual int launches_fused() const { return 0; }  // kernels of the three calls above
    // Fused body for solves WITH a projector (the IGM re-solves): update_xr_handover as above, then cycle_to_z = the cycle on the handed-over
    // residual with z written in fp64 and rho_new = <r, z>, beta finished in its last kernel (no stand-alone hand-back / dot kernel); the
    // projector and p = z + beta p follow. The two hooks below let a wrapping preconditioner (igm.cuh ConstraintPrecond) finish the cycle itself.
    virtual bool fused_projected() const { return false; }
    virtual void cycle_to_z(const T*, T*, PcgState*, ReduceWs, cudaStream_t) {}
    virtual int launches_fused_projected() const { return 0; }  // kernels of update_xr_handover + cycle_to_z
    virtual void cycle_only(PcgState*, cudaStream_t) {}          // the cycle alone; result (fp32, scaled by 1 / zscale) at lo_result()
    virtual const void* lo_result() const { return nullptr; }
};

// In-place projection onto the feasible subspace (IGM constraint elimination); null for unconstrained solves.
template <class T>
struct Projector {
    virtual ~Projector() = default;
    virtual void apply(T* v, const PcgState* st, cudaStream_t s) = 0;
};

struct SolveStats {
    int iterations = 0;
    int restarts = 0;
    double relres = 0.0;  // true ||b - A x|| / ||b|| at exit
    int status = 0;       // 0 ok, 1 maxit, 2 breakdown
    long launches = 0;
    double iter_ms = 0.0;  // device time of the iteration batches alone (events; only measured when MK_VERBOSE is set)
};

template <class T>
class Pcg {
public:
    Pcg(int n, cudaStream_t stream) : n_(n), stream_(stream), r_(n), z_(n), p_(n), q_(n), st_(1), partials_(kMaxPartials * 4), ticket_(1) {
        ticket_.zero(stream_);
        st_.zero(stream_);
        h_st_ = (PcgState*)PinnedStatePool::instance().get(sizeof(PcgState));
        grid_ = stream_grid(n);
    }
    ~Pcg() {
        if (ev0_) { cudaEventDestroy(ev0_); cudaEventDestroy(ev1_); }
        if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
        PinnedStatePool::instance().put(h_st_);
    }
    PcgState* state() { return st_.p; }
    ReduceWs ws() { return ReduceWs{partials_.p, ticket_.p}; }
    T* q() { return q_.p; }
    T* r() { return r_.p; }
    int grid() const { return grid_; }

    void bind(LinOp<T>* A, Precond<T>* M, Projector<T>* P) {
        if (A != A_ || M != M_ || P != P_) invalidate_graph();
        A_ = A; M_ = M; P_ = P;
    }
    void invalidate_graph() {
        if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
    }

    // Solve A x = b starting from the x passed in. If residual_ready, r() already holds b - A x and the state
    // (rr, bnorm2, thresh, iter, maxit, done) was set by k_residual.
    // With a projector the right-hand side is projected in place first (r = P b - P A x stays in the feasible subspace).
    // `hint`: expected iteration count (e.g. that of the previous, similar solve) -- sizes the first batch of iterations
    // issued before the host looks at the state; later batches are `chunk` long. Iterations issued after convergence
    // are no-ops on the device (every kernel returns on st->done) but still cost launch latency.
    // The iteration stops on the RECURRENCE residual, the result is accepted on the TRUE one (b - A x, recomputed): close to the fp64 floor
    // (eps * cond(A): 5e-11 at 1M vertices against tol = 1e-10) the two differ by tens of percent, and 49 of the 50 field solves of a step
    // came back from the check just above the tolerance and restarted (a preconditioner application, two more SpMVs and two host syncs
    // each). The recurrence is therefore driven to kRecurrenceMargin * tol: one more iteration instead of a restart.
    static constexpr double kRecurrenceMargin = 0.7;
    static double recurrence_tol2(double tol) { return tol * tol * kRecurrenceMargin * kRecurrenceMargin; }  // for callers that start the state themselves
    SolveStats solve(T* b, T* x, double tol, int maxit, bool residual_ready, int chunk = 16, bool use_graph = true, int hint = 0) {
        NvtxRange nvtx_("pcg_solve");
        SolveStats out;
        long launches = 0;
        const double tol2 = tol * tol * kRecurrenceMargin * kRecurrenceMargin;
        if (!residual_ready) {
            if (P_) { P_->apply(b, nullptr, stream_); launches++; }
            A_->apply(x, q_.p, nullptr, ws(), stream_);
            if (P_) P_->apply(q_.p, nullptr, stream_);
            launch_residual(b, tol2, maxit, 0);
            launches += 1 + A_->launches_per_apply();
        }
        bool breakdown = false;
        double prev_relres = 1e300;
        static const bool timed = getenv("MK_VERBOSE") != nullptr;
        if (timed && !ev0_) { MK_CUDA(cudaEventCreate(&ev0_)); MK_CUDA(cudaEventCreate(&ev1_)); }
        for (int attempt = 0; attempt < 4; attempt++) {
            start_direction();
            launches += 1 + (M_ ? M_->launches_per_apply() : 0);
            // First batch: the expected count. After a restart from the true residual only a few iterations are missing, and every
            // iteration issued past convergence is ~26 empty launches (42 us at 1M vertices in the CUPTI trace, profiles/trace_r02_*):
            // round 1 re-issued `hint` of them after every restart, 10 % of the field solve.
            int batch = attempt > 0 ? std::max(2, chunk / 4) : (hint > 0 ? std::min(std::max(hint, 1), 8 * chunk) : chunk);
            if (timed) MK_CUDA(cudaEventRecord(ev0_, stream_));
            while (true) {
                for (int c = 0; c < batch; c++) iteration(x, use_graph);
                launches += (long)batch * launches_per_iteration();
                fetch_state();
                if (h_st_->done) break;
                batch = (hint > 0 || attempt > 0) ? std::max(2, chunk / 4) : chunk;
            }
            if (timed) {
                MK_CUDA(cudaEventRecord(ev1_, stream_));
                MK_CUDA(cudaEventSynchronize(ev1_));
                float ms = 0;
                MK_CUDA(cudaEventElapsedTime(&ms, ev0_, ev1_));
                out.iter_ms += ms;
            }
            breakdown = h_st_->done == 2;
            // true residual check (recurrence residuals drift at the 1e-10 level)
            A_->apply(x, q_.p, nullptr, ws(), stream_);
            if (P_) P_->apply(q_.p, nullptr, stream_);
            launch_residual(b, tol2, maxit, 1);
            launches += 1 + A_->launches_per_apply();
            fetch_state();
            out.iterations = h_st_->iter;
            out.relres = h_st_->bnorm2 > 0 ? std::sqrt(h_st_->rr / h_st_->bnorm2) : 0.0;
            out.restarts = attempt;
            if (out.relres <= tol || h_st_->bnorm2 == 0 || h_st_->iter >= maxit) break;  // accepted on the true residual against tol itself
            if (breakdown) break;
            // The recurrence residual met the threshold but the true one did not: restart from the true residual -- unless the
            // last restart gained less than 2x, i.e. fp64 is exhausted (attainable ||b - A x|| / ||b|| ~ eps * cond(A): ~5e-11 at 1M
            // vertices, ~4e-10 at 8M for right-hand sides along the smallest eigenvector, which inverse iteration produces)
            if (attempt > 0 && out.relres > 0.5 * prev_relres) break;
            prev_relres = out.relres;
        }
        // 0 ok; 1 tolerance not reached (iteration cap or fp64 floor; the result is still the best iterate); 2 breakdown (<p, A p> <= 0 or NaN)
        out.status = (out.relres <= tol * 1.0000001 || h_st_->bnorm2 == 0) ? 0 : (breakdown ? 2 : 1);
        out.launches = launches;
        return out;
    }

    bool fused() const { return M_ && !P_ && M_->fused_pcg(); }
    bool fused_projected() const { return M_ && P_ && M_->fused_projected(); }
    int launches_per_iteration() const {
        if (fused()) return A_->launches_per_apply() + M_->launches_fused();
        if (fused_projected()) return A_->launches_per_apply() + 2 + M_->launches_fused_projected() + 1;
        return A_->launches_per_apply() + (P_ ? (M_ ? 2 : 1) : 0) + 1 + (M_ ? M_->launches_per_apply() + 1 : 0) + 1;
    }

    void launch_residual(const T* b, double tol2, int maxit, int keep) {
        k_residual<T><<<grid_, kThreads, 0, stream_>>>(n_, b, q_.p, r_.p, st_.p, ws(), tol2, maxit, keep);
        MK_CHECK_LAUNCH();
    }
    void fetch_state() {
        MK_CUDA(cudaMemcpyAsync(h_st_, st_.p, sizeof(PcgState), cudaMemcpyDeviceToHost, stream_));
        MK_CUDA(cudaStreamSynchronize(stream_));
    }
    const PcgState& host_state() const { return *h_st_; }

private:
    void start_direction() {
        if (M_) {
            M_->apply(r_.p, z_.p, st_.p, stream_);
            if (P_) P_->apply(z_.p, st_.p, stream_);
            k_start_p<T><<<grid_, kThreads, 0, stream_>>>(n_, r_.p, z_.p, p_.p, st_.p, ws());
        } else {
            k_start_p<T><<<grid_, kThreads, 0, stream_>>>(n_, r_.p, r_.p, p_.p, st_.p, ws());
        }
        MK_CHECK_LAUNCH();
    }
    void iteration_body(T* x) {
        A_->apply(p_.p, q_.p, st_.p, ws(), stream_);
        if (fused()) {
            M_->update_xr_handover(n_, x, r_.p, p_.p, q_.p, st_.p, ws(), stream_);
            M_->cycle_dot(r_.p, st_.p, ws(), stream_);
            M_->update_p(n_, p_.p, st_.p, stream_);
            MK_CHECK_LAUNCH();
            return;
        }
        if (P_) P_->apply(q_.p, st_.p, stream_);
        if (fused_projected()) {
            M_->update_xr_handover(n_, x, r_.p, p_.p, q_.p, st_.p, ws(), stream_);
            M_->cycle_to_z(r_.p, z_.p, st_.p, ws(), stream_);
            P_->apply(z_.p, st_.p, stream_);  // <r, P z> = <P r, z> = <r, z>: rho is not touched by cleaning z
            k_update_p<T><<<grid_, kThreads, 0, stream_>>>(n_, p_.p, z_.p, st_.p);
            MK_CHECK_LAUNCH();
            return;
        }
        if (M_) {
            k_update_xr<T, true><<<grid_, kThreads, 0, stream_>>>(n_, x, r_.p, p_.p, q_.p, st_.p, ws());
            // without a projector the preconditioner may finish <r, z> in its own last kernel
            if (P_ || !M_->apply_dot(r_.p, z_.p, st_.p, ws(), stream_)) {
                M_->apply(r_.p, z_.p, st_.p, stream_);
                if (P_) P_->apply(z_.p, st_.p, stream_);
                k_dot_rz<T><<<grid_, kThreads, 0, stream_>>>(n_, r_.p, z_.p, st_.p, ws());
            }
            k_update_p<T><<<grid_, kThreads, 0, stream_>>>(n_, p_.p, z_.p, st_.p);
        } else {
            k_update_xr<T, false><<<grid_, kThreads, 0, stream_>>>(n_, x, r_.p, p_.p, q_.p, st_.p, ws());
            k_update_p<T><<<grid_, kThreads, 0, stream_>>>(n_, p_.p, r_.p, st_.p);
        }
        MK_CHECK_LAUNCH();
    }
    void iteration(T* x, bool use_graph) {
        if (!use_graph) { iteration_body(x); return; }
        if (!graph_exec_ || graph_x_ != x) {
            invalidate_graph();
            cudaGraph_t g;
            MK_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
            iteration_body(x);
            MK_CUDA(cudaStreamEndCapture(stream_, &g));
            MK_CUDA(cudaGraphInstantiate(&graph_exec_, g, 0));
            cudaGraphDestroy(g);
            graph_x_ = x;
        }
        MK_CUDA(cudaGraphLaunch(graph_exec_, stream_));
    }

    int n_;
    cudaStream_t stream_;
    DevBuf<T> r_, z_, p_, q_;
    DevBuf<PcgState> st_;
    DevBuf<double> partials_;
    DevBuf<unsigned int> ticket_;
    PcgState* h_st_ = nullptr;
    int grid_ = 1;
    LinOp<T>* A_ = nullptr;
    Precond<T>* M_ = nullptr;
    Projector<T>* P_ = nullptr;
    cudaGraphExec_t graph_exec_ = nullptr;
    T* graph_x_ = nullptr;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
};

}  // namespace mk
