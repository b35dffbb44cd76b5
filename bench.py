#!/usr/bin/env python
"""bench.py -- metriko hot path on B200: N-RoSy field eigen-solve + IGM Poisson / iterative rounding at 1M vertices.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's CPU algorithm on the box's host cores)

One "step" = one pass of the hot path over one mesh (BASELINE.json configs[2], octahedron subdivided 9x = 1,048,578
vertices, N = 4, example_igm flags seamless / !localInjectivity / !roundSeams):

    field : connectionLaplacian + galerkinMassMatrix assembly (frf.h:46-84), multigrid set-up (stands where the reference
            factorises, sls.h:42-43), 50 inverse-iteration steps (sls.h:44-52)
    IGM   : integ() (integ.ipp:3-98): G2 / F2 / X2F assembly, A0 = length^2 X2F^T X2F + hierarchy, Poisson solve
            (ir_init.h:173-212), NI closed form, SeamlessIntegration rounding loop with one warm-started projected PCG re-solve per
            rounded dof (ir.h:90-110), result scatter nfn / cfn (integ.ipp:88-98)

`value`  : steps timed with the step's inputs already resident in HBM (CUDA events, max over ranks): the half-edge arrays, u0,
           X2F and F2 as device arrays (mk_field_solve_dev, mk_igm_create_dev).
`e2e`    : the same step through the reference-facing API with HOST (pinned) buffers, everything integ() reads going up and
           everything it returns coming back inside the timed region: half-edge arrays + u0 up, the field down; cut-mesh corner
           table + positions, fullx2Nfunc, UFull and the combing turns up (X2F and F2 are assembled on the device from them:
           mk_igm_create_from_mesh), fullx, nfn and cfn down.
`roofline`: the face-stencil SpMV (the dominant kernel by bytes) timed live, plus `iteration` -- the bytes EVERY kernel of one PCG
           iteration has to move (all multigrid levels, stored formats) over the measured time per iteration, i.e. how far the
           whole solve is from the HBM roofline, and `cycle_us` -- the live per-level time of the V-cycle (graph replays).
N > 1    : batch mode -- every rank solves its own 1M-vertex mesh, no data-path collective (weak scaling) -- plus a
           `partitioned` object: BASELINE configs[3], ONE 8,388,610-vertex mesh row-partitioned over the N GPUs (halo exchange + global
           multigrid over NCCL), SpMV / PCG / inverse-iteration times.
`reference_sample`: the B200 arm timed on the SAME mesh the reference arm's step runs on, so that one same-config ratio exists.

Only the `cpu_baseline` leg and `--impl reference` touch oracle/ (the CPU restatement of the reference algorithm).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))  # tests/hostglue: the synthetic-input generator (host stages of the reference)

METRIC = "field+IGM solve throughput at 1M verts (ms_per_step = solve ms)"
UNIT = "verts/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
# dram__bytes_read.sum + dram__bytes_write.sum per launch of k_spmv_face<false> from the committed `ncu --set full` capture
# (profiles/ncu_full_spmv_kernels_r01b.csv: 159.5 MB read + 19.8 MB written), keyed by the face count it was taken at
NCU_TRAFFIC_BYTES = {2097152: 179.3e6}
NCU_TRAFFIC_SOURCE = ("profiles/ncu_full_spmv_kernels_r01b.csv (ncu --set full, per launch, k_spmv_face<false>: 159.5 MB read + 19.8 MB written); re-captured in round 2 on its "
                      "fused-dot twin k_spmv_face<true> (same loads and stores): 159.5 + 19.2 MB, profiles/ncu_full_kernels_r02.csv")
REFERENCE_BUDGET_S = 240.0   # the reference arm sizes its sample so that all K + W steps end within this
# seconds per step of the reference arm (oracle, ni_mode="lm", every factorisation the reference performs) measured on this pool's
# hosts, by octahedron subdivision level; only used to PICK the sample level, the run is timed live
REFERENCE_EST_S = {3: 0.6, 4: 2.5, 5: 17.0, 6: 110.0}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# workload generation (host glue = the reference's unchanged host stages; untimed)
# ---------------------------------------------------------------------------------------------------------------------
def make_session(levels: int, rank: int):
    from hostglue import Session
    return Session().make_sphere(8, levels, bump=0.08, phase=0.37 * rank)


def random_u0(nF: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)  # VecXc::Random range (sls.h:32), explicit because the reference's is unseeded
    return rng.uniform(-1.0, 1.0, nF) + 1j * rng.uniform(-1.0, 1.0, nF)


def clocks_sampler(path: str):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                stderr=subprocess.DEVNULL), f
    except Exception:
        return None, None


def clocks_summary(path: str, device: int) -> dict:
    sm, mx, reasons = [], 0, set()
    try:
        for line in open(path):
            p = [t.strip() for t in line.split(",")]
            if len(p) < 9 or not p[0].isdigit() or int(p[0]) != device:
                continue
            sm.append(int(float(p[1])))
            mx = max(mx, int(float(p[2])))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    busy = [c for c in sm if c >= 0.5 * mx] or sm  # samples taken while the GPU is clocked up (host stages idle it)
    return {"sm_mhz": int(statistics.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (CPU restatement of the reference algorithm with SuperLU direct solves)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(levels: int, seed: int = 0, ni_mode: str = "lm"):
    """One pass of the reference algorithm on a bounded sample (octahedron subdivided `levels` times), exactly as the reference runs
    it: NI-stage LM over [x; f] (`ni_mode="lm"`), every LM iteration re-factorising (no reuse_factor). Returns (seconds of the hot
    path, nV, per-stage dict). Only the hot path is timed; the host stages between the solves are the reference's unchanged code."""
    import oracle
    s = make_session(levels, 0)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    u0 = random_u0(nF, seed)
    t0 = time.perf_counter()
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    M = oracle.galerkin_mass_matrix(fi["faceArea"])
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    z = oracle.solve_smallest_eig(L, M, u0, 50)
    t_eig = time.perf_counter() - t0
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    t1 = time.perf_counter()
    st = {}
    out = oracle.iterative_rounding_x2f(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"],
                                        inp["UFull"], inp["X2F"], inp["N"], inp["n"], True, False, inp["F2"], ni_mode=ni_mode, stats=st)
    t_igm = time.perf_counter() - t1
    assert out["success"]
    det = dict(field_assemble_s=t_asm, field_lu_and_50_solves_s=t_eig, igm_poisson_ni_rounding_s=t_igm, n_solves=out["n_solves"],
               factorizations=st.get("factorizations", 0), lm_iterations=st.get("lm_iterations", 0), ni_mode=ni_mode, faces=nF)
    return t_asm + t_eig + t_igm, s.dims["nV"], det


def oracle_runs_at_baseline_sizes():
    """Timings of the oracle's full runs at the BASELINE sizes, recorded by tests/golden/make_golden_large.py when the committed golden
    fixtures were generated (the builder's container CPU, one thread, reuse_factor=True: identical re-factorisations skipped, so a LOWER
    bound of what the reference's LM loop costs). Context only: not measured in this run, not on this box."""
    out = []
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in ("icosphere", "octa7", "torus320", "octa8", "octa9"):
        p = os.path.join(gdir, f"large_{name}.npz")
        if not os.path.exists(p):
            continue
        try:
            g = np.load(p)
            tm = json.loads(str(g["timings"]))
            out.append({"mesh": name, "verts": int(g["dims"][0]), "faces": int(g["dims"][1]), "ni_mode": str(g["ni_mode"]),
                        "field_s": round(tm["field_assemble_s"] + tm["field_lu_and_50_solves_s"], 1), "igm_s": round(tm["igm_s"], 1),
                        "lm_factorizations_reference_would_run": int(g["lm_factorizations"]), "lm_factorizations_run": int(g["lm_factorizations_done"])})
        except Exception:
            pass
    return out


def reference_level(args) -> int:
    n = args.steps + min(args.warmup, 1)
    lv = 3
    for cand in sorted(REFERENCE_EST_S):
        if cand <= args.cpu_levels and n * REFERENCE_EST_S[cand] <= REFERENCE_BUDGET_S:
            lv = cand
    return lv


def run_reference(args, rank: int):
    if rank != 0:
        return
    lv = reference_level(args)
    for _ in range(min(args.warmup, 1)):   # a CPU direct solver has nothing to warm up beyond the page cache: at most one untimed pass
        cpu_reference_step(lv)
    ts, dets = [], []
    nV = 0
    for _ in range(args.steps):
        t, nV, det = cpu_reference_step(lv)
        ts.append(t)
        dets.append(det)
    per = statistics.median(ts)
    det = dets[ts.index(per)] if per in ts else dets[-1]
    val = nV / per
    nproc = os.cpu_count()
    sample = (f"octahedron x{lv} ({nV} verts, {det['faces']} faces) per step, the mesh family of the 1M-vertex workload at the largest size whose "
              f"{args.steps}+{min(args.warmup, 1)} steps fit {REFERENCE_BUDGET_S:.0f} s: assembly {det['field_assemble_s']:.2f}s, SuperLU complex LU + 50 solves "
              f"{det['field_lu_and_50_solves_s']:.2f}s, IGM KKT LU + NI-stage LM + {det['n_solves']} rounding re-solves through the reference's LM loop = "
              f"{det['factorizations']} LU factorisations {det['igm_poisson_ni_rounding_s']:.2f}s; median of {args.steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"4-RoSy field + IGM iterative rounding, bounded sample of the 1M-vertex config: {sample}",
                       "sample_levels": lv, "sample_verts": nV,
                       "note": "the reference (header-only C++ over Eigen/libigl) cannot be compiled in this image; this is the oracle port, single-threaded like "
                               "the reference (no OpenMP / threads anywhere in metriko; SuperLU is sequential). A sparse direct path is NOT size-invariant in "
                               "verts/s: the B200 arm prints `reference_sample` (its own time on this very mesh) for a same-config ratio, and "
                               "`oracle_runs_at_baseline_sizes` lists the oracle's full runs at 65k ... 1M vertices recorded when the golden fixtures were made"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores_available": nproc, "kind": "port", "sample": sample, "stages_s": det},
            "oracle_runs_at_baseline_sizes": oracle_runs_at_baseline_sizes(),
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------------
def pinned_like(a: np.ndarray):
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype if a.dtype != np.complex128 else torch.complex128,
                    pin_memory=True)
    n = t.numpy()
    n[...] = a
    return t, n


def iteration_bytes(hier, rows_bytes0: dict, real: bool, deg0: int = 1, degc: int = 2) -> dict:
    """Bytes every kernel of ONE PCG iteration has to move (DESIGN.md "Bytes per iteration"): the fp64 PCG kernels and level 0 of the fp32
    cycle from `rows_bytes0` (bytes per level-0 row of each kernel, stored formats), the coarser levels from the hierarchy
    [(rows, entries of A_l, entries of P_l)]: 2 deg_l SpMVs with A_l (entries x (value + 4 B index) + ~5 vector entries per row) plus
    restriction and prolongation (entries of P_l twice + a vector on either side)."""
    v = 4 if real else 8            # bytes of one fp32 vector / matrix value
    n0 = hier[0][0]
    level0 = sum(rows_bytes0.values()) * n0 + 2 * hier[0][2] * (v + 4)
    coarse = 0
    for l in range(1, len(hier)):
        rows, na, npp = hier[l]
        if l == len(hier) - 1:
            coarse += rows * rows * v + 2 * rows * v      # dense inverse
        else:
            coarse += 2 * degc * (na * (v + 4) + 5 * rows * v) + 2 * npp * (v + 4) + 4 * rows * v
    return {"level0_bytes": int(level0), "coarse_levels_bytes": int(coarse), "total_bytes": int(level0 + coarse)}


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from metriko_b200 import api, capi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = capi.Context(local_rank, stream=stream.cuda_stream)

    def workload(levels, seed_rank):
        """mesh, one field solve to derive the IGM inputs through the host stages (untimed)"""
        s = make_session(levels, seed_rank)
        fi = s.field_inputs()
        nF, nV = s.dims["nF"], s.dims["nV"]
        u0 = random_u0(nF, seed_rank)
        F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
        z = F.solveSmallestEig(u0, 50)
        s.field_post(z, 4, "curl")
        dims = s.igm_setup(0.02)
        inp = s.igm_inputs()
        mesh = s.igm_mesh_inputs()
        return s, fi, nF, nV, u0, F, inp, mesh, dims

    t0 = time.time()
    s, fi, nF, nV, u0, F, inp, mesh, dims = workload(args.levels, rank)
    log(f"[rank {rank}] workload ready in {time.time()-t0:.1f}s: nV={nV} nF={nF} igm={dims} X2F nnz={inp['X2F'].nnz}")
    X2F, UFull, F2N = inp["X2F"].tocsr(), inp["UFull"].tocsr(), mesh["fullx2Nfunc"].tocsr()
    for M in (X2F, UFull, F2N):
        M.sort_indices()
    N = inp["N"]

    # device-resident inputs for `value`
    u0_t = torch.from_numpy(u0).to(dev)
    x_t = torch.empty_like(u0_t)
    x2f_dev = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (X2F.indptr.astype(np.int64), X2F.indices.astype(np.int32),
                                                                            X2F.data.astype(np.float64), np.asarray(inp["F2"], np.float64))]
    dev_ptrs = tuple(int(t.data_ptr()) for t in x2f_dev)

    def host_buffers(fi, u0, UFull, F2N, mesh):
        keep, host = [], {}
        for k in ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea"):
            t, host[k] = pinned_like(np.ascontiguousarray(fi[k]))
            keep.append(t)
        for k, a in (("u0", u0),
                     ("uf_ptr", UFull.indptr.astype(np.int64)), ("uf_col", UFull.indices.astype(np.int32)), ("uf_val", UFull.data.astype(np.float64)),
                     ("f2n_ptr", F2N.indptr.astype(np.int64)), ("f2n_col", F2N.indices.astype(np.int32)), ("f2n_val", F2N.data.astype(np.float64)),
                     ("cutF", np.ascontiguousarray(mesh["cutF"], np.int32).reshape(-1)), ("cutPos", np.ascontiguousarray(mesh["cutPos"], np.float64).reshape(-1)),
                     ("turns", np.ascontiguousarray(mesh["comb_turns"], np.int32))):
            t, host[k] = pinned_like(a)
            keep.append(t)
        nF_, cnv, N_ = len(fi["faceArea"]), len(mesh["cutPos"]), 4
        for k, a in (("out_z", np.zeros(nF_, np.complex128)), ("out_nfn", np.zeros((cnv, N_), np.float64)), ("out_cfn", np.zeros((nF_, 3 * N_), np.float64))):
            t, host[k] = pinned_like(a)   # pinned result buffers (not counted as H2D)
            keep.append(t)
        return keep, host

    keep, host = host_buffers(fi, u0, UFull, F2N, mesh)
    # the X2F arrays of the resident path's UFull argument (host, read by the rounding loop)
    x2f_host = (X2F.indptr.astype(np.int64), X2F.indices.astype(np.int32), X2F.data.astype(np.float64))
    h2d = sum(host[k].nbytes for k in host if not k.startswith("out_"))
    cut_nV = len(mesh["cutPos"])
    d2h = nF * 16 + UFull.shape[0] * 8 + cut_nV * N * 8 + nF * 3 * N * 8   # field, fullx, nfn, cfn
    stage = {}
    e2e_stage = []

    def step_resident():
        """inputs in HBM: field arrays uploaded once (F), u0 / x device tensors, X2F and F2 device tensors (mk_igm_create_dev);
        UFull is read on the host by the rounding loop (candidate rows) and uploaded by the library"""
        ta = time.perf_counter()
        F.assemble()
        st_f = F.solve_dev(u0_t.data_ptr(), 50, x_t.data_ptr())
        tb = time.perf_counter()
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"], raw=(x2f_host[0], x2f_host[1], x2f_host[2],
                                                                             host["uf_ptr"], host["uf_col"], host["uf_val"], None), dev=dev_ptrs)
        out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
        prob.close()
        tc = time.perf_counter()
        assert out["success"]
        stage.update(field_setup_ms=st_f["setup_ms"], field_solve_ms=st_f["solve_ms"], field_inner=st_f["inner_iterations"],
                     field_ms=(tb - ta) * 1e3, igm_ms=(tc - tb) * 1e3, igm_setup_ms=out["stats"]["setup_ms"], igm_solve_ms=out["stats"]["solve_ms"],
                     igm_inner=out["stats"]["inner_iterations"], igm_resolves=out["n_solves"], relres_field=st_f["relres_max"],
                     relres_igm=out["stats"]["relres_max"])
        return 1 + st_f["kernel_launches"] + out["stats"]["kernel_launches"]

    def make_step_e2e(host, inp, N, UFull, F2N, nF, record):
        def step_e2e():
            """reference-facing API with host buffers: FaceRosyField ctor -> solveSmallestEig (field down) -> integ(): operator assembly on the
            device from the cut-mesh arrays + the resident field + the combing turns -> iterative_rounding -> result scatter (nfn, cfn down)"""
            t = [time.perf_counter()]
            Fe = api.FaceRosyField(ctx, host["face2half"], host["next"], host["twin"], host["face"], host["heArgOnFace"], host["faceArea"], 4)
            t.append(time.perf_counter())
            ze = Fe.solveSmallestEig(host["u0"], 50, out=host["out_z"])
            t.append(time.perf_counter())
            prob = api.IgmProblem.from_mesh(ctx, N, None, None, F2N, UFull, inp["length"], field=Fe,
                                            raw=(host["f2n_ptr"], host["f2n_col"], host["f2n_val"], host["uf_ptr"], host["uf_col"], host["uf_val"],
                                                 host["cutF"], host["cutPos"], host["turns"]))
            t.append(time.perf_counter())
            out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
            t.append(time.perf_counter())
            nfn, cfn = prob.scatter(out=(host["out_nfn"], host["out_cfn"]))
            t.append(time.perf_counter())
            prob.close()
            Fe.close()
            t.append(time.perf_counter())
            assert out["success"] and ze.shape[0] == nF and cfn.shape[0] == nF
            if record is not None:
                record.append(dict(zip(("field_create", "field_solve", "igm_create_from_mesh", "igm_round", "igm_scatter", "close"),
                                       [round((b - a) * 1e3, 1) for a, b in zip(t[:-1], t[1:])])))
            return out
        return step_e2e

    step_e2e = make_step_e2e(host, inp, N, UFull, F2N, nF, e2e_stage)

    def timed(fn, warm, steps, profile=False):
        launches = 0
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if profile:
                torch.cuda.profiler.start()  # ncu --profile-from-start off: only the timed region is captured
            e0.record(stream)
            for _ in range(steps):
                r = fn()
                launches += r if isinstance(r, int) else 0
            e1.record(stream)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            if profile:
                torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), launches

    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, fh = clocks_sampler(clk_path) if rank == 0 else (None, None)
    ms_total, launches = timed(step_resident, args.warmup, args.steps, profile=args.profile)
    if args.profile:  # a run under ncu is never a bench value: stop here
        log(f"profile run: {launches} launches in the profiled region; no bench line printed")
        return
    ms_e2e, _ = timed(step_e2e, 1, args.steps)
    if proc:
        proc.terminate()
        proc.wait()
        fh.close()
    clocks = clocks_summary(clk_path, local_rank) if rank == 0 else {}

    # ---- roofline of the dominant kernel: the face-stencil SpMV k_spmv_face (field PCG), timed live with CUDA events on
    # the launching stream; the working set (vals + nbrs + x + y = 92 B/face resident, 243 MB at 1M verts) exceeds L2 (126 MB)
    reps = 200
    spmv_ms = F.spmv_bench(reps, False)
    spmv_dot_ms = F.spmv_bench(reps, True)
    alg_bytes = 116 * nF  # CSR-equivalent algorithmic bytes per row (nnz*20 + 4 + 32)
    stored_bytes = 92 * nF  # what the ELL-3 stencil actually has to move (48 values + 12 indices + 16 x + 16 y)
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (spmv_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_spmv_face<false> (y = L x, complex fp64 face stencil)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC_BYTES.get(nF), "traffic_source": NCU_TRAFFIC_SOURCE if nF in NCU_TRAFFIC_BYTES else None,
                "peak_source": peak_src, "us_per_launch": spmv_ms * 1e3,
                "algorithmic_bytes_per_launch": alg_bytes, "stored_format_bytes_per_launch": stored_bytes,
                "achieved_stored_format_gbs": stored_bytes / (spmv_ms * 1e-3) / 1e9, "fused_dot_us_per_launch": spmv_dot_ms * 1e3}
    # whole PCG iteration against the roofline (all kernels, all levels; DESIGN.md "Bytes per iteration")
    try:
        fh_ = F.hierarchy()
        # per face: SpMV + <p,Lp> (92) | x, r update + hand-over (p 16, q 16, x 32, r 32, dinv 8, b 8, d0 8) | fp32 stencil step (36 + v 8, b 8, r 8)
        # | restriction reads r (8) | prolongation x (16) | fp32 stencil residual + last update + <r,z> (36 + x 8, b 8, dinv 8, r_hi 16, xo 8) | p update (8 + 32)
        fb = iteration_bytes(fh_, {"spmv_dot": 92, "update_xr_handover": 120, "cheb_step_first": 60, "restrict": 8, "prolong": 16,
                                   "resid_last_dot": 84, "update_p": 40}, real=False)
        us_it = stage["field_solve_ms"] * 1e3 / max(1, stage["field_inner"])
        roofline["iteration"] = {"field": {"hierarchy_rows_nnzA_nnzP": fh_, **fb, "us_per_iteration_incl_outer_loop": us_it,
                                            "achieved": fb["total_bytes"] / (us_it * 1e-6) / 1e9, "frac": fb["total_bytes"] / (us_it * 1e-6) / 1e9 / peak,
                                            "cycle_us_from_level": F.profile_cycle(100)}}
    except Exception as e:  # noqa: BLE001
        roofline["iteration"] = {"field": {"error": repr(e)}}
    try:
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"])
        a_ms = prob.spmv_bench(reps, False)
        A = prob.matrix()
        a_bytes = A.nnz * 12 + (A.shape[0] + 1) * 4 + A.shape[0] * 16
        roofline["igm_spmv"] = {"kernel": "k_spmv_sell<double, kPlain> (y = (A0 + 1e-4 I) x, SELL-32 + long-row CTAs in one launch)", "rows": A.shape[0], "nnz": int(A.nnz),
                                "us_per_launch": a_ms * 1e3, "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak}
        prob.poisson(inp["fixedIdcs"], inp["fixedVals"])   # one solve so that the hierarchy's work buffers exist
        ih = prob.hierarchy()
        npr = A.nnz / A.shape[0]
        # per row (real): fp64 SELL SpMV + dot (nnz 12 + 4 + x 8 + y 8) | x, r update fused with the hand-over (p 8, q 8, x 16, r 16, dinv 4, b 4, d0 4) |
        # fp32 SELL step (nnz 8 + 4 + v 4, b 4, r 4) | restrict 4 | prolong 8 | fp32 SELL residual + last (nnz 8 + 4 + x 4, b 4, dinv 4, xo 4) |
        # end of the constraint preconditioner fused with the hand-back and <r,z> (z_lo 4, W 4 k in fp32 with k ~ 7 on average, r 8, z 8) | p update (8 + 16)
        ib = iteration_bytes(ih, {"spmv_dot": npr * 12 + 20, "update_xr_handover": 60, "cheb_step_first": npr * 8 + 16, "restrict": 4, "prolong": 8,
                                  "resid_last": npr * 8 + 20, "cprec_finish": 20 + 4 * 7, "update_p": 24}, real=True)
        us_it = stage["igm_solve_ms"] * 1e3 / max(1, stage["igm_inner"])
        roofline["iteration"]["igm"] = {"hierarchy_rows_nnzA_nnzP": ih, **ib, "us_per_iteration_incl_rounding_loop": us_it,
                                        "achieved": ib["total_bytes"] / (us_it * 1e-6) / 1e9, "frac": ib["total_bytes"] / (us_it * 1e-6) / 1e9 / peak,
                                        "cycle_us_from_level": prob.profile_cycle(100)}
        prob.close()
    except Exception as e:  # noqa: BLE001
        roofline["igm_spmv"] = roofline.get("igm_spmv", {"error": repr(e)})
        roofline["iteration"]["igm"] = {"error": repr(e)}

    # ---- the B200 arm on the reference arm's own sample mesh (same config for one ratio), through the e2e path, median of 3
    ref_sample = None
    if world == 1:
        try:
            lv = reference_level(args)
            s2, fi2, nF2, nV2, u02, F2_, inp2, mesh2, _ = workload(lv, 0)
            U2, Fn2 = inp2["UFull"].tocsr(), mesh2["fullx2Nfunc"].tocsr()
            U2.sort_indices()
            Fn2.sort_indices()
            keep2, host2 = host_buffers(fi2, u02, U2, Fn2, mesh2)
            step2 = make_step_e2e(host2, inp2, inp2["N"], U2, Fn2, nF2, None)
            ts = []
            with torch.cuda.stream(stream):
                step2()
                for _ in range(3):
                    torch.cuda.synchronize()
                    ta = time.perf_counter()
                    step2()
                    torch.cuda.synchronize()
                    ts.append((time.perf_counter() - ta) * 1e3)
            ref_sample = {"mesh": f"octahedron x{lv}", "verts": nV2, "faces": nF2, "ms_per_step_e2e": statistics.median(ts), "verts_per_s": nV2 / (statistics.median(ts) * 1e-3),
                          "note": "the mesh `bench.py --impl reference` times with these --steps / --warmup; launch-latency-bound at this size"}
            F2_.close()
        except Exception as e:  # noqa: BLE001
            ref_sample = {"error": repr(e)}

    # ---- N > 1: BASELINE configs[3], one 8M-vertex mesh partitioned over the N GPUs
    partitioned = None
    if world > 1 and not args.no_partitioned:
        try:
            from scripts.bench_partitioned import run_partitioned
            F.close()
            del x2f_dev
            torch.cuda.empty_cache()
            partitioned = run_partitioned(ctx, rank, world, kind=4, levels=args.partitioned_levels, reps=50, outer=3)
        except Exception as e:  # noqa: BLE001
            partitioned = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    per = ms_total / args.steps
    per_e2e = ms_e2e / args.steps
    line = {"metric": METRIC, "value": world * nV / (per * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"4-RoSy field (50 inverse iterations) + IGM Poisson + iterative rounding on an octahedron subdivided "
                                   f"{args.levels}x ({nV} verts, {nF} faces) per GPU, example_igm flags (seamless, !localInjectivity, !roundSeams)",
                       "N": 4, "verts_per_gpu": nV, "faces_per_gpu": nF, "igm_unknowns": int(UFull.shape[1]), "inner_tol": 1e-10,
                       "parallelism": f"batch x{world} (one mesh per GPU, no data-path collective)",
                       "l2": "inputs larger than L2 (field stencil working set 92 B/face = %.0f MB)" % (92 * nF / 1e6)},
            "e2e": {"value": world * nV / (per_e2e * 1e-3), "unit": UNIT, "ms_per_step": per_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": "host buffers -> FaceRosyField / solveSmallestEig -> integ() from the cut-mesh arrays (X2F, F2 assembled on the device) -> "
                            "iterative_rounding -> nfn / cfn scatter -> host"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stage, "e2e_stages_ms": e2e_stage[-args.steps:]}
    if ref_sample is not None:
        line["reference_sample"] = ref_sample
    if partitioned is not None:
        line["partitioned"] = partitioned
    if world == 1 and not args.no_cpu:
        lv = reference_level(args)
        t, cv, det = cpu_reference_step(lv)
        line["cpu_baseline"] = {"value": cv / t, "unit": UNIT, "cores": 1, "host_cores_available": os.cpu_count(), "kind": "port", "ms": t * 1e3,
                                "sample": f"octahedron x{lv} ({cv} verts): oracle (scipy SuperLU LU + 50 solves; IGM KKT LU + NI-stage LM + "
                                          f"{det['n_solves']} re-solves through the reference's LM loop = {det['factorizations']} LU factorisations), "
                                          f"single thread like the reference; the same mesh as `reference_sample`", "stages_s": det}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--levels", type=int, default=9, help="octahedron subdivisions (9 -> 1,048,578 vertices)")
    ap.add_argument("--cpu-levels", type=int, default=5, help="largest bounded sample for the CPU reference / cpu_baseline (shrunk to fit the time budget)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-partitioned", action="store_true", help="N > 1: skip the partitioned 8M-vertex leg")
    ap.add_argument("--partitioned-levels", type=int, default=11, help="tetrahedron subdivisions of the partitioned mesh (11 -> 8,388,610 vertices)")
    ap.add_argument("--profile", action="store_true", help="bracket the timed region with cudaProfilerStart/Stop (for ncu) and exit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
