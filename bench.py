#!/usr/bin/env python
"""bench.py -- metriko hot path on B200: N-RoSy field eigen-solve + IGM Poisson / iterative rounding at 1M vertices.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference ...                      (the reference's CPU algorithm on the box's host cores)

One "step" = one pass of the hot path over one mesh (BASELINE.json configs[2], octahedron subdivided 9x = 1,048,578
vertices, N = 4, example_igm flags seamless / !localInjectivity / !roundSeams):

    field : connectionLaplacian + galerkinMassMatrix assembly (frf.h:46-84), multigrid set-up (stands where the reference
            factorises, sls.h:42-43), 50 inverse-iteration steps (sls.h:44-52)
    IGM   : A0 = length^2 X2F^T X2F assembly + hierarchy, Poisson solve (ir_init.h:173-212), NI closed form,
            SeamlessIntegration rounding loop with one warm-started projected PCG re-solve per rounded dof (ir.h:90-110)

`value`  : steps timed with the step's inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step through the reference-facing API with HOST buffers (pinned): H2D of every input and D2H of
           the field + fullx inside the timed region.
N > 1    : batch mode -- every rank solves its own 1M-vertex mesh, no data-path collective (weak scaling).

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
NCU_TRAFFIC_SOURCE = "profiles/ncu_full_spmv_kernels_r01b.csv (ncu --set full, per launch)"


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
def cpu_reference_step(levels: int, seed: int = 0):
    """One pass of the reference algorithm on a bounded sample (octahedron subdivided `levels` times).
    Returns (seconds_hot_path, nV). Only the hot path is timed: L/M assembly + LU + 50 solves, IGM Poisson + rounding."""
    import oracle
    s = make_session(levels, 0)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    u0 = random_u0(nF, seed)
    t0 = time.perf_counter()
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    M = oracle.galerkin_mass_matrix(fi["faceArea"])
    z = oracle.solve_smallest_eig(L, M, u0, 50)
    t_field = time.perf_counter() - t0
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    t1 = time.perf_counter()
    st = {}
    out = oracle.iterative_rounding_x2f(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"],
                                        inp["UFull"], inp["X2F"], inp["N"], inp["n"], True, False, inp["F2"], ni_mode="closed", stats=st)
    t_igm = time.perf_counter() - t1
    assert out["success"]
    return t_field + t_igm, s.dims["nV"], dict(field_s=t_field, igm_s=t_igm, n_solves=out["n_solves"], factorizations=st.get("factorizations", 0))


def run_reference(args, rank: int):
    if rank != 0:
        return
    lv = args.cpu_levels if args.steps + args.warmup <= 6 else min(args.cpu_levels, 4)  # keep the whole run within minutes
    for _ in range(args.warmup):
        cpu_reference_step(lv)
    ts = []
    nV = 0
    det = {}
    for _ in range(args.steps):
        t, nV, det = cpu_reference_step(lv)
        ts.append(t)
    per = sum(ts) / len(ts)
    val = nV / per
    sample = (f"octahedron x{lv} ({nV} verts) per step: scipy SuperLU complex LU + 50 solves (field {det['field_s']:.1f}s), IGM KKT LU + "
              f"{det['n_solves']} rounding re-solves through the reference's LM loop = {det['factorizations']} LU factorisations (igm {det['igm_s']:.1f}s)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"4-RoSy field + IGM iterative rounding, bounded sample of the 1M-vertex config: {sample}",
                       "note": "the reference (header-only C++ over Eigen/libigl) cannot be compiled in this image; this is the oracle port, "
                               "single-threaded like the reference (no OpenMP/threads anywhere in metriko)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
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

    # ---- workload (untimed): mesh, one field solve to derive the IGM inputs through the host stages
    t0 = time.time()
    s = make_session(args.levels, rank)
    fi = s.field_inputs()
    nF, nV = s.dims["nF"], s.dims["nV"]
    u0 = random_u0(nF, rank)
    F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
    z = F.solveSmallestEig(u0, 50)
    s.field_post(z, 4, "curl")
    dims = s.igm_setup(0.02)
    inp = s.igm_inputs()
    log(f"[rank {rank}] workload ready in {time.time()-t0:.1f}s: nV={nV} nF={nF} igm={dims} X2F nnz={inp['X2F'].nnz}")
    X2F, UFull = inp["X2F"].tocsr(), inp["UFull"].tocsr()
    X2F.sort_indices()
    UFull.sort_indices()

    # device-resident inputs for `value`
    u0_t = torch.from_numpy(u0).to(dev)
    x_t = torch.empty_like(u0_t)
    x2f_dev = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (X2F.indptr.astype(np.int64), X2F.indices.astype(np.int32),
                                                                            X2F.data.astype(np.float64), np.asarray(inp["F2"], np.float64))]
    dev_ptrs = tuple(int(t.data_ptr()) for t in x2f_dev)

    # pinned host copies for `e2e`
    keep = []
    host = {}
    for k in ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea"):
        t, host[k] = pinned_like(np.ascontiguousarray(fi[k]))
        keep.append(t)
    for k, a in (("u0", u0), ("F2", np.ascontiguousarray(inp["F2"], np.float64)),
                 ("x2f_ptr", X2F.indptr.astype(np.int64)), ("x2f_col", X2F.indices.astype(np.int32)), ("x2f_val", X2F.data.astype(np.float64)),
                 ("uf_ptr", UFull.indptr.astype(np.int64)), ("uf_col", UFull.indices.astype(np.int32)), ("uf_val", UFull.data.astype(np.float64))):
        t, host[k] = pinned_like(a)
        keep.append(t)
    h2d = sum(host[k].nbytes for k in host)
    d2h = nF * 16 + UFull.shape[0] * 8
    stage = {}
    e2e_stage = []

    def step_resident():
        """inputs in HBM: field arrays uploaded once (F), u0 / x device tensors, X2F and F2 device tensors (mk_igm_create_dev);
        UFull is read on the host by the rounding loop (candidate rows) and uploaded by the library"""
        ta = time.perf_counter()
        F.assemble()
        st_f = F.solve_dev(u0_t.data_ptr(), 50, x_t.data_ptr())
        tb = time.perf_counter()
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"], raw=(host["x2f_ptr"], host["x2f_col"], host["x2f_val"],
                                                                             host["uf_ptr"], host["uf_col"], host["uf_val"], host["F2"]), dev=dev_ptrs)
        out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
        prob.close()
        tc = time.perf_counter()
        assert out["success"]
        stage.update(field_setup_ms=st_f["setup_ms"], field_solve_ms=st_f["solve_ms"], field_inner=st_f["inner_iterations"],
                     field_ms=(tb - ta) * 1e3, igm_ms=(tc - tb) * 1e3, igm_setup_ms=out["stats"]["setup_ms"], igm_solve_ms=out["stats"]["solve_ms"],
                     igm_inner=out["stats"]["inner_iterations"], igm_resolves=out["n_solves"], relres_field=st_f["relres_max"],
                     relres_igm=out["stats"]["relres_max"])
        return 1 + st_f["kernel_launches"] + out["stats"]["kernel_launches"]

    def step_e2e():
        """reference-facing API with host buffers: create -> assemble -> solveSmallestEig -> iterative_rounding"""
        t = [time.perf_counter()]
        Fe = api.FaceRosyField(ctx, host["face2half"], host["next"], host["twin"], host["face"], host["heArgOnFace"], host["faceArea"], 4)
        t.append(time.perf_counter())
        ze = Fe.solveSmallestEig(host["u0"], 50)
        t.append(time.perf_counter())
        Fe.close()
        t.append(time.perf_counter())
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"], raw=(host["x2f_ptr"], host["x2f_col"], host["x2f_val"],
                                                                             host["uf_ptr"], host["uf_col"], host["uf_val"], host["F2"]))
        t.append(time.perf_counter())
        out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
        t.append(time.perf_counter())
        prob.close()
        t.append(time.perf_counter())
        assert out["success"] and ze.shape[0] == nF
        e2e_stage.append(dict(zip(("field_create", "field_solve", "field_close", "igm_create", "igm_round", "igm_close"),
                                  [round((b - a) * 1e3, 1) for a, b in zip(t[:-1], t[1:])])))
        return out

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
    try:
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"])
        a_ms = prob.spmv_bench(reps, False)
        A = prob.matrix()
        a_bytes = A.nnz * 12 + (A.shape[0] + 1) * 4 + A.shape[0] * 16
        roofline["igm_spmv"] = {"kernel": "k_spmv_sell<double, kPlain> (y = (A0 + 1e-4 I) x, SELL-32 + long-row CTAs in one launch)", "rows": A.shape[0], "nnz": int(A.nnz),
                                "us_per_launch": a_ms * 1e3, "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak}
        prob.close()
    except Exception as e:  # noqa: BLE001
        roofline["igm_spmv"] = {"error": repr(e)}

    if rank != 0:
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
            "e2e": {"value": world * nV / (per_e2e * 1e-3), "unit": UNIT, "ms_per_step": per_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "stages": stage, "e2e_stages_ms": e2e_stage[-args.steps:]}
    if world == 1 and not args.no_cpu:
        t, cv, det = cpu_reference_step(args.cpu_levels)
        line["cpu_baseline"] = {"value": cv / t, "unit": UNIT, "cores": 1, "kind": "port", "ms": t * 1e3,
                                "sample": f"octahedron x{args.cpu_levels} ({cv} verts): oracle (scipy SuperLU LU + 50 solves; IGM KKT LU + "
                                          f"{det['n_solves']} re-solves through the reference's LM loop = {det['factorizations']} LU factorisations), "
                                          f"single thread like the reference; field {det['field_s']:.1f}s igm {det['igm_s']:.1f}s"}
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
    ap.add_argument("--cpu-levels", type=int, default=5, help="bounded sample for the CPU reference / cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true")
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
