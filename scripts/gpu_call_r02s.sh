#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for th in 0 0.03 0.05; do
  MK_AMG_THETA_FIELD=$th timeout 900 python scripts/bench_sweep.py --levels 6 7 8 9 10 --torus > gpurun_out/r02s_sweep_theta$th.jsonl 2> gpurun_out/r02s_sweep_theta$th.err
  echo "theta_field $th"; python - <<PY
import json
for l in open('gpurun_out/r02s_sweep_theta$th.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,1) if isinstance(v,float) and v>1 else v) for k,v in d.items() if k in ('mesh','verts','step_ms','field_setup_ms','field_solve_ms','field_inner','igm_setup_ms','igm_solve_ms','igm_inner','igm_resolves','field_relres_max','field_status')})
PY
done
(timeout 600 python -m pytest tests/test_gpu_field.py -q --tb=short 2>&1 | tail -5)
