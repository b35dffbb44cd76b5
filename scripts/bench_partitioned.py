"""BASELINE configs[3]: graph-partitioned SpMV / PCG on a subdivided mesh across N B200s (torchrun, one rank per GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_partitioned.py --levels 11

Prints one JSON line (rank 0): per-launch time of the distributed SpMV with and without overlap (CUDA events on the launching
stream, max over ranks), achieved aggregate GB/s (116 B/row algorithmic), halo sizes, and one PCG solve L x = M u0
to 1e-10 with its iteration count and time. Kind 4 / levels 11 = tetrahedron x11 = 8,388,610 vertices."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metriko_b200 import capi  # noqa: E402
from metriko_b200.dist import DistContext, DistFaceField  # noqa: E402
from hostglue import Session  # noqa: E402


def run_partitioned(ctx, rank, world, kind=4, levels=11, reps=100, outer=0, both=False):
    """The partitioned-mode measurements on an already initialised process group / capi.Context; returns the result dict on every rank.
    Used by this script and by bench.py (--gpus N > 1: the `partitioned` object of the bench line)."""
    d = DistContext(ctx, rank, world)
    t0 = time.time()
    s = Session().make_sphere(kind, levels, bump=0.08)
    fi = s.field_inputs()
    nF, nV = s.dims["nF"], s.dims["nV"]
    t_mesh = time.time() - t0
    t0 = time.time()
    F = DistFaceField.from_mesh(d, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
    plan = F.plan
    ctx.sync()
    t_setup = time.time() - t0

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_ov = maxr(F.spmv_bench(reps, True))
    ms_no = maxr(F.spmv_bench(reps, False))
    rng = np.random.default_rng(0)
    u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
    rows = plan.owned_global()
    b = (fi["faceArea"] * u0)[rows]
    res = {}
    F.pcg(b, precond=1, max_iter=4)   # untimed warm-up: lazy NCCL connections for the level-1 all-reduce, kernel loading
    for precond in ((1, 2) if both else (1,)):   # 1 = global hierarchy (coarse levels replicated), 2 = block-Jacobi hierarchy
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.time()
        try:
            if os.environ.get("PART_TRACE") and rank == 0:   # CUPTI timeline of the timed solve (then NOT a timing run): kernel time by name
                from torch.profiler import ProfilerActivity, profile
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    F.pcg(b, precond=precond)
                    torch.cuda.synchronize()
                import collections
                path = os.environ["PART_TRACE"]
                prof.export_chrome_trace(path + ".json")
                ev = [e for e in json.load(open(path + ".json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
                os.remove(path + ".json")
                by = collections.defaultdict(lambda: [0, 0.0])
                for e in ev:
                    k = e["name"].split("(")[0].replace("void ", "").replace("mk::", "")
                    by[k][0] += 1
                    by[k][1] += e["dur"]
                tot = sum(v[1] for v in by.values())
                span = (max(e["ts"] + e["dur"] for e in ev) - min(e["ts"] for e in ev)) if ev else 0.0
                with open(path, "w") as fh:
                    fh.write(f"{len(ev)} device activities, span {span / 1e3:.2f} ms, busy {tot / 1e3:.2f} ms\n")
                    for k, (c, tt) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
                        fh.write(f"{tt / 1e3:9.2f} ms {100 * tt / max(tot, 1e-9):5.1f} %  {c:6d} x {tt / c:8.2f} us  {k}\n")
            else:
                F.pcg(b, precond=precond)
        except Exception as e:  # noqa: BLE001
            res[f"pcg_precond{precond}_error"] = repr(e)
        res[f"pcg_precond{precond}"] = dict(iterations=int(F.stats["inner_iterations"]), relres=F.stats["relres_max"], setup_ms=F.stats["setup_ms"],
                                            solve_ms=maxr(F.stats["solve_ms"]), wall_s=maxr(time.time() - t))
    # the same solve with b and x resident in HBM (what `value` means in bench.py): the host-buffer call above moves 2 x 16 B per owned row
    # through pageable memory inside its solve_ms
    try:
        bd = torch.from_numpy(np.ascontiguousarray(b).view(np.float64)).cuda()
        xd = torch.zeros_like(bd)
        torch.cuda.synchronize()
        F.pcg_dev(bd.data_ptr(), xd.data_ptr(), precond=1)
        xd.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = time.time()
        F.pcg_dev(bd.data_ptr(), xd.data_ptr(), precond=1)
        res["pcg_precond1_resident"] = dict(iterations=int(F.stats["inner_iterations"]), relres=F.stats["relres_max"], solve_ms=maxr(F.stats["solve_ms"]),
                                            wall_s=maxr(time.time() - t))
    except Exception as e:  # noqa: BLE001
        res["pcg_precond1_resident_error"] = repr(e)
    if outer:
        t = time.time()
        try:
            F.solveSmallestEig(u0[rows], outer)
        except Exception as e:  # noqa: BLE001
            res["inverse_iteration_error"] = repr(e)
        res["inverse_iteration"] = dict(outer=outer, inner=int(F.stats["inner_iterations"]), solve_ms=maxr(F.stats["solve_ms"]),
                                        relres=F.stats["relres_max"], wall_s=maxr(time.time() - t))
    halo = maxr(plan.n_halo)
    alg = 116 * nF
    out = {"mode": "partitioned", "n_gpus": world, "verts": nV, "faces": nF, "rows_per_gpu": plan.n_local, "halo_max": int(halo),
           "interior_frac": plan.n_interior / plan.n_local, "spmv_ms_overlap": ms_ov, "spmv_ms_no_overlap": ms_no,
           "spmv_aggregate_gbs_algorithmic": alg / (ms_ov * 1e-3) / 1e9, "halo_bytes_per_spmv_per_gpu": int(halo) * 16,
           "allreduces_per_pcg_iteration": 3, "mesh_s": t_mesh, "operator_and_hierarchy_setup_s": t_setup, **res}
    F.close()
    d.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", type=int, default=4)
    ap.add_argument("--levels", type=int, default=11)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--outer", type=int, default=0, help="also run this many inverse-iteration steps")
    ap.add_argument("--both", action="store_true", help="also time the block-Jacobi preconditioner")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    out = run_partitioned(ctx, rank, world, args.kind, args.levels, args.reps, args.outer, args.both)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
