#!/bin/bash
# round-2 GPU call 13: whole GPU test-suite (incl. deterministic + generic-config tests), smoke, the full bench line,
# cfg2 / cfg4 / WGAN-GP / decode lines, launch lists at batch 64 and 32
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c13_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c13_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_smoke.log 2>&1; tail -1 gpurun_out/c13_smoke.log
timeout 900 python bench.py > gpurun_out/c13_bench.log 2> gpurun_out/c13_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads([l for l in open('gpurun_out/c13_bench.log') if l.startswith('{')][-1]);print('b256',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_hbm']['frac'], d['roofline_hbm']['traffic'], d['vs_gpu_lib'], d['step_frac_of_ideal'])"
timeout 600 python bench.py --global-batch 64 --skip-cpu-baseline > gpurun_out/c13_bench_cfg2_b64.log 2> gpurun_out/c13_bench_cfg2.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c13_bench_cfg2_b64.log') if l.startswith('{')][-1]);print('b64',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['vs_gpu_lib'])"
timeout 600 python bench.py --workload cfg4 --skip-cpu-baseline --skip-lib-baseline --steps 10 > gpurun_out/c13_bench_cfg4_n1.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c13_bench_cfg4_n1.log') if l.startswith('{')][-1]);print('cfg4 n1',d['value'],d['ms_per_step'],d['launches_per_step'], d['step_model_tflops_per_gpu'])"
timeout 600 python bench.py --loss-mode wgan_gp --optimizer rmsprop --skip-cpu-baseline --skip-lib-baseline --steps 10 > gpurun_out/c13_bench_wgan_gp.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c13_bench_wgan_gp.log') if l.startswith('{')][-1]);print('wgan_gp',d['value'],d['ms_per_step'],d['launches_per_step'])"
timeout 600 python bench.py --workload decode > gpurun_out/c13_decode.log 2> gpurun_out/c13_decode.err; tail -c 600 gpurun_out/c13_decode.log
for B in 64 32; do
  timeout 300 python scripts/profile_step.py $B > gpurun_out/c13_step_b$B.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c13_launches_b$B.csv python scripts/profile_step.py $B > gpurun_out/c13_ncu_b$B.log 2>&1
  python scripts/summarize_launches.py gpurun_out/c13_launches_b$B.csv > gpurun_out/c13_launches_b${B}_summary.txt 2>&1; head -12 gpurun_out/c13_launches_b${B}_summary.txt
done
