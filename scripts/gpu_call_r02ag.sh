#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02ag_cycle.json 2> gpurun_out/r02ag_cycle.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02ag_cycle.json').read().strip().splitlines()[-1])
for k in ('field','igm'):
    v=d[k]; key=[x for x in v if x.startswith('cycle_us')][0]
    print(k, [round(x,1) for x in v[key]], 'spmv_dot_us', round(v['spmv_dot_us'],1), 'solve_ms', round(v['solve_ms'],1), 'inner', v['inner'], 'us/it', round(v['us_per_iteration'],1))
PY
