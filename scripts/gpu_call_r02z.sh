#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for rm in 0 500000; do
  MK_ROWMODE_BELOW=$rm timeout 300 python scripts/cycle_probe.py 9 > gpurun_out/r02z_cycle_rowmode_$rm.json 2> gpurun_out/r02z_cycle_rowmode_$rm.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02z_cycle_rowmode_$rm.json').read().strip().splitlines()[-1])
for k in ('field','igm'):
    v=d[k]; key=[x for x in v if x.startswith('cycle_us')][0]
    print('rowmode_below $rm', k, [round(x,1) for x in v[key]], 'solve_ms', round(v['solve_ms'],1), 'inner', v['inner'], 'us/it', round(v['us_per_iteration'],1))
PY
done
timeout 900 python scripts/bench_batch.py --meshes 64 --gpus-total 8 --workers 1 > gpurun_out/r02z_batch_w1.json 2> gpurun_out/r02z_batch_w1.err; cat gpurun_out/r02z_batch_w1.json
timeout 900 python scripts/bench_batch.py --meshes 64 --gpus-total 8 --workers 4 > gpurun_out/r02z_batch_w4.json 2> gpurun_out/r02z_batch_w4.err; cat gpurun_out/r02z_batch_w4.json; tail -3 gpurun_out/r02z_batch_w4.err
