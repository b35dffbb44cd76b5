#!/bin/bash
# Round-end gpurun call: GPU tests, default-flags bench line, smoke(), per-solve iteration counts (typed / untyped aggregation) and the
# ncu launch lists of PCG iterations of both solves (cold-cache, serialised: shares, not absolutes).
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_next.log 2>&1; tail -3 gpurun_out/pytest_gpu_next.log
timeout 150 python bench.py > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err
python -c "
import json;d=json.load(open('gpurun_out/bench_next.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stages'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['igm_spmv']['frac'])"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()"
timeout 90 python scripts/profile_kernels.py prep /tmp/wl.npz 9 > gpurun_out/prep.log 2>&1
MK_VERBOSE=1 timeout 40 python scripts/profile_kernels.py run /tmp/wl.npz igm,round > gpurun_out/igm_verbose_default.log 2>&1
MK_VERBOSE=1 MK_IGM_CPREC=0 timeout 40 python scripts/profile_kernels.py run /tmp/wl.npz igm,round > gpurun_out/igm_verbose_projected.log 2>&1
grep -c "igm solve" gpurun_out/igm_verbose_default.log gpurun_out/igm_verbose_projected.log
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 320 --csv --log-file gpurun_out/launches_next_igm.csv python scripts/profile_kernels.py run /tmp/wl.npz igm > gpurun_out/ncu_igm.log 2>&1
timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 320 --csv --log-file gpurun_out/launches_next_field.csv python scripts/profile_kernels.py run /tmp/wl.npz field > gpurun_out/ncu_field.log 2>&1
wc -l gpurun_out/launches_next_*.csv
