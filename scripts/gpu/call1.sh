#!/bin/bash
# round-2 GPU call 1: parity of the streaming BatchNorm kernels + contract fixes, first bench line with the
# stock-PyTorch baseline, BN kernel bandwidth sweep, launch list at batch 64.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
for f in test_gpu_round2 test_gpu_kernels test_gpu_modules; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -x -q -s > gpurun_out/c1_$f.log 2>&1; echo "$f rc=$?" | tee -a gpurun_out/c1_$f.log
  grep -E "passed|failed|error" gpurun_out/c1_$f.log | tail -3
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; tail -2 gpurun_out/c1_smoke.log
timeout 300 python scripts/sweep_bn.py 128 > gpurun_out/c1_sweep_bn_128.txt 2>&1
timeout 300 python scripts/sweep_bn.py 64 > gpurun_out/c1_sweep_bn_64.txt 2>&1
VG_BN_STREAM=0 timeout 300 python scripts/sweep_bn.py 128 > gpurun_out/c1_sweep_bn_128_old.txt 2>&1
cat gpurun_out/c1_sweep_bn_128.txt gpurun_out/c1_sweep_bn_128_old.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.log 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/c1_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 --global-batch 64 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c1_bench_b64.log 2> gpurun_out/c1_bench_b64.err
VG_BN_STREAM=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c1_bench_oldbn.log 2> gpurun_out/c1_bench_oldbn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c1_launches_b64.csv python scripts/profile_step.py 64 > gpurun_out/c1_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/c1_launches_b64.csv > gpurun_out/c1_launches_b64_summary.txt 2>&1
head -30 gpurun_out/c1_launches_b64_summary.txt
