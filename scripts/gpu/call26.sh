#!/bin/bash
# round-2 GPU call 26: TMA-store epilogue also for the scatter phases (VG_TC_TMA_STORE=2): parity, A/B against =1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VG_TC_TMA_STORE=2 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_full_size.py tests/test_gpu_round2.py tests/test_gpu_generic_configs.py tests/test_gpu_modules.py -q -x > gpurun_out/c26_tests_tma2.log 2>&1; echo "tests (TMA store 2) rc=$?"; tail -3 gpurun_out/c26_tests_tma2.log
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c26_$name.log 2> gpurun_out/c26_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c26_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c26_$name.err
}
BARGS="--global-batch 256"; run_bench b256_tma1 VG_TC_TMA_STORE=1; run_bench b256_tma2 VG_TC_TMA_STORE=2; run_bench b256_tma1_2 VG_TC_TMA_STORE=1; run_bench b256_tma2_2 VG_TC_TMA_STORE=2
VG_TC_TMA_STORE=2 timeout 300 python scripts/sweep_conv.py 64 > gpurun_out/c26_sweep_tma2.txt 2>&1; tail -1 gpurun_out/c26_sweep_tma2.txt
