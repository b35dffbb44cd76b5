#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python scripts/trace_iterations.py igm_round 9 > gpurun_out/r02y_trace_igm.log 2>&1
python scripts/trace_summary.py gpurun_out/trace_igm_round.csv 30
python - <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/trace_igm_round.csv')))
names=[r['name'] for r in rows]
idx=[i for i,n in enumerate(names) if 'k_spmv_sell<double, 2>' in n]
k=len(idx)//2
a,b=idx[k],idx[k+1]
t0=float(rows[a]['start_us'])
for r in rows[a:b]:
    print(f"{float(r['start_us'])-t0:8.1f} {float(r['dur_us']):7.2f}  {r['name'][:80]}")
print('iter span', float(rows[b]['start_us'])-t0)
PY
