#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r02fc_tests.log
tail -4 gpurun_out/r02fc_tests.log
timeout 900 python bench.py > gpurun_out/r02fc_bench.json 2> gpurun_out/r02fc_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02fc_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'], d['roofline']['frac'])
print({k: d['roofline']['iteration'][k].get('frac') for k in d['roofline'].get('iteration', {})}, d.get('reference_sample'))
PY
timeout 900 python scripts/bench_sweep.py --levels 6 7 8 9 10 --torus > gpurun_out/r02fc_sweep.jsonl 2> gpurun_out/r02fc_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/r02fc_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print({k:(round(v,1) if isinstance(v,float) and v>1 else v) for k,v in d.items() if k in ('mesh','verts','step_ms','field_setup_ms','field_solve_ms','field_inner','igm_setup_ms','igm_solve_ms','igm_inner','igm_resolves','field_relres_max','field_status')})
PY
