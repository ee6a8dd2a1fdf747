#!/bin/bash
# round-2 GPU call 4: staged coalesced epilogue stores, reverse-order BN reductions, avgpool backward, new tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_round2 test_gpu_modules; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -x -q > gpurun_out/c4_$f.log 2>&1; echo "$f rc=$?" | tee -a gpurun_out/c4_$f.log
  grep -E "passed|failed|error" gpurun_out/c4_$f.log | tail -3
done
timeout 300 python scripts/sweep_conv.py > gpurun_out/c4_sweep_conv.txt 2>&1; cat gpurun_out/c4_sweep_conv.txt
timeout 300 python scripts/sweep_bn.py 128 > gpurun_out/c4_sweep_bn_128.txt 2>&1; cat gpurun_out/c4_sweep_bn_128.txt
run_bench() { # name, env..., batch
  local name=$1; shift; local b=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --global-batch $b --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c4_bench_$name.log 2>&1
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c4_bench_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'])"
}
run_bench b256 256 A=1
run_bench b64 64 A=1
run_bench b32 32 A=1
run_bench b256_norev 256 VG_BN_REVERSE=0
run_bench b64_norev 64 VG_BN_REVERSE=0
for b in 64; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c4_launches_b$b.csv python scripts/profile_step.py $b > gpurun_out/c4_ncu_b$b.log 2>&1
python scripts/summarize_launches.py gpurun_out/c4_launches_b$b.csv > gpurun_out/c4_launches_b${b}_summary.txt 2>&1
done
head -32 gpurun_out/c4_launches_b64_summary.txt
# ncu full capture of the degenerate wgrad strip kernel + one bn stream kernel for stall analysis
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_degenerate_strip -c 2 -o gpurun_out/c4_ncu_degenerate python scripts/profile_step.py 64 > gpurun_out/c4_ncu_deg.log 2>&1
ls -la gpurun_out/c4_ncu_degenerate.ncu-rep
