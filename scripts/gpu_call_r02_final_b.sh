#!/bin/bash
# ncu evidence of round 2 (the .ncu-rep stays on the box: it is exported to CSV there, gpurun_out is limited to 64 MiB) + the bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02fb_bench.json 2> gpurun_out/r02fb_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02fb_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'], d['roofline']['frac'])
print({k: d['roofline']['iteration'][k].get('frac') for k in d['roofline'].get('iteration', {})})
PY
timeout 300 python scripts/profile_kernels.py prep /tmp/wl.npz 9 > gpurun_out/r02fb_prep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 600 --csv --log-file gpurun_out/launches_r02_field.csv python scripts/profile_kernels.py run /tmp/wl.npz field > gpurun_out/r02fb_ncu_field.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 600 --csv --log-file gpurun_out/launches_r02_igm.csv python scripts/profile_kernels.py run /tmp/wl.npz igm > gpurun_out/r02fb_ncu_igm.log 2>&1
wc -l gpurun_out/launches_r02_*.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_face|k_spmv_sell|k_update_xr_handover|k_cprec_finish|k_multi_dot|k_update_p_lo" --launch-skip 300 --launch-count 40 -f -o /tmp/ncu_r02_full python scripts/profile_kernels.py run /tmp/wl.npz field,igm > gpurun_out/r02fb_ncu_full.log 2>&1
ls -la /tmp/ncu_r02_full.ncu-rep
ncu -i /tmp/ncu_r02_full.ncu-rep --page raw --csv > gpurun_out/ncu_full_r02_raw.csv 2> gpurun_out/r02fb_ncu_export.err
ncu -i /tmp/ncu_r02_full.ncu-rep --page source --csv --kernel-name regex:k_spmv_face -c 1 > gpurun_out/ncu_source_spmv_face_r02.csv 2>> gpurun_out/r02fb_ncu_export.err
ls -la gpurun_out/ncu_full_r02_raw.csv gpurun_out/ncu_source_spmv_face_r02.csv
du -sh gpurun_out
