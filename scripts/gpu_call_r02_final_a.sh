#!/bin/bash
# Round-2 evidence call (one B200): GPU suite, smoke, default-flags bench line, reference arm, sanitizer passes, ncu launch lists + one
# `--set full` capture of the hot kernels. Nothing printed by a run under ncu / compute-sanitizer is a bench value.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r02fa_tests.log
tail -6 gpurun_out/r02fa_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fa_smoke.log 2>&1; tail -2 gpurun_out/r02fa_smoke.log
timeout 900 python bench.py > gpurun_out/r02fa_bench.json 2> gpurun_out/r02fa_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02fa_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'], d['roofline']['frac'])
print({k: d['roofline']['iteration'][k]['frac'] for k in d['roofline'].get('iteration', {})})
PY
timeout 900 python bench.py --impl reference > gpurun_out/r02fa_ref.json 2> gpurun_out/r02fa_ref.err; cut -c1-400 gpurun_out/r02fa_ref.json
# sanitizer: racecheck (shared-memory hazards of the ticket reductions / SELL / projector kernels) and memcheck on the small goldens
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_igm.py -q -x -k "golden and sphere_ico3" > gpurun_out/r02fa_racecheck_igm.log 2>&1; tail -4 gpurun_out/r02fa_racecheck_igm.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_field.py -q -x -k "golden and sphere_ico3 and 8" > gpurun_out/r02fa_racecheck_field.log 2>&1; tail -4 gpurun_out/r02fa_racecheck_field.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fa_memcheck_smoke.log 2>&1; tail -4 gpurun_out/r02fa_memcheck_smoke.log
# ncu
timeout 300 python scripts/profile_kernels.py prep /tmp/wl.npz 9 > gpurun_out/r02fa_prep.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/launches_r02_field.csv python scripts/profile_kernels.py run /tmp/wl.npz field > gpurun_out/r02fa_ncu_field.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 700 --csv --log-file gpurun_out/launches_r02_igm.csv python scripts/profile_kernels.py run /tmp/wl.npz igm > gpurun_out/r02fa_ncu_igm.log 2>&1
wc -l gpurun_out/launches_r02_*.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_spmv_face|k_spmv_sell|k_update_xr_handover|k_cprec_finish|k_multi_dot|k_update_p_lo" --launch-skip 300 --launch-count 48 -f -o gpurun_out/ncu_r02_full python scripts/profile_kernels.py run /tmp/wl.npz field,igm > gpurun_out/r02fa_ncu_full.log 2>&1
ls -la gpurun_out/ncu_r02_full.ncu-rep
