#!/bin/bash
# round-2 GPU call 22: smoke + the default bench line + batch-64 launch list on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c22_smoke.log 2>&1; tail -1 gpurun_out/c22_smoke.log
timeout 900 python bench.py > gpurun_out/c22_bench.log 2> gpurun_out/c22_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads([l for l in open('gpurun_out/c22_bench.log') if l.startswith('{')][-1]);print('b256',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['vs_gpu_lib'], d['step_frac_of_ideal'])"
timeout 300 python scripts/profile_step.py 64 > gpurun_out/c22_step_b64.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c22_launches_b64.csv python scripts/profile_step.py 64 > gpurun_out/c22_ncu_b64.log 2>&1
python scripts/summarize_launches.py gpurun_out/c22_launches_b64.csv > gpurun_out/c22_launches_b64_summary.txt 2>&1; head -5 gpurun_out/c22_launches_b64_summary.txt
