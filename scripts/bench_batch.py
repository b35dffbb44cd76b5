"""BASELINE configs[4]: a batch of 64 synthetic ~50k-vertex meshes (224x224 tori with different bump phases) sharded over the GPUs of
one box, one mesh at a time per GPU, no data-path collective (SURVEY.md 8e "batch"). Works with or without torchrun:

  python scripts/bench_batch.py --meshes 64 [--gpus-total 8]         one process: this rank's share only (rank 0 of gpus-total)
  python -m torch.distributed.run --nproc-per-node N scripts/bench_batch.py --meshes 64

Per mesh: field (assemble + hierarchy + 50 inverse iterations) -> host glue (matching / seam / setup, untimed in `gpu_ms`) -> IGM
(operator + hierarchy + Poisson + rounding). Prints one JSON line: meshes/s over the GPU stages, and wall meshes/s including host glue.

--workers K: K host threads per GPU, each with its OWN context (stream + arena; "one host thread per context", include/metriko_b200.h),
meshes dealt round-robin. At 50k vertices every kernel is latency-bound (100k rows), so the streams of several meshes interleave on the
SMs and the host glue of one mesh (ctypes C++, GIL released) overlaps the GPU stages of another."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metriko_b200 import api, capi  # noqa: E402
from hostglue import Session  # noqa: E402

FIELD_KEYS = ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--meshes", type=int, default=64)
    ap.add_argument("--grid", type=int, default=224)
    ap.add_argument("--gpus-total", type=int, default=0, help="without torchrun: pretend to be rank 0 of this many GPUs")
    ap.add_argument("--workers", type=int, default=1, help="host threads (contexts) per GPU")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.gpus_total and world == 1:
        world = args.gpus_total
    import torch
    torch.cuda.set_device(local)
    mine = list(range(rank, args.meshes, world))
    wall0 = time.perf_counter()

    def work(ids):
        ctx = capi.Context(local)
        gpu_ms, inner, nv = 0.0, 0, 0
        for i in ids:
            rng = np.random.default_rng(i)
            s = Session().make_torus(args.grid, args.grid, bump=0.1, phase=float(rng.uniform(0, 2 * np.pi)))
            fi = s.field_inputs()
            nF = s.dims["nF"]
            u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
            t = time.perf_counter()
            F = api.FaceRosyField(ctx, *[fi[k] for k in FIELD_KEYS], 4)
            z = F.solveSmallestEig(u0, 50)
            inner += F.stats["inner_iterations"]
            F.close()
            gpu_ms += (time.perf_counter() - t) * 1e3
            s.field_post(z, 4, "curl")
            s.igm_setup(0.02)
            inp = s.igm_inputs()
            t = time.perf_counter()
            prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"])
            out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
            prob.close()
            gpu_ms += (time.perf_counter() - t) * 1e3
            assert out["success"], i
            inner += out["stats"]["inner_iterations"]
            nv += s.dims["nV"]
        ctx.close()
        return gpu_ms, inner, nv

    k = max(1, min(args.workers, len(mine)))
    if k == 1:
        parts = [work(mine)]
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(k) as ex:
            parts = list(ex.map(work, [mine[w::k] for w in range(k)]))
    gpu_ms, inner, nv = (sum(p[j] for p in parts) for j in range(3))
    wall = time.perf_counter() - wall0
    print(json.dumps({"mode": "batch", "rank": rank, "gpus": world, "meshes_total": args.meshes, "meshes_this_rank": len(mine), "workers": k, "wall_meshes_per_s_this_rank": len(mine) / (time.perf_counter() - wall0), "verts_per_mesh": nv // max(len(mine), 1),
                      "gpu_stage_ms_per_mesh": gpu_ms / max(len(mine), 1), "wall_s_per_mesh_incl_host_glue": wall / max(len(mine), 1),
                      "meshes_per_s_gpu_stages_this_rank": len(mine) / (gpu_ms * 1e-3), "inner_iterations_per_mesh": inner / max(len(mine), 1)}), flush=True)


if __name__ == "__main__":
    main()
