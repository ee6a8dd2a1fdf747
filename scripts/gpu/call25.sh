#!/bin/bash
# round-2 GPU call 25: TMA-store epilogue (VG_TC_TMA_STORE=1): parity tests under it, A/B of the step and of the conv sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VG_TC_TMA_STORE=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_full_size.py tests/test_gpu_round2.py tests/test_gpu_generic_configs.py -q -x > gpurun_out/c25_tests_tma.log 2>&1; echo "tests (TMA store) rc=$?"; tail -3 gpurun_out/c25_tests_tma.log
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c25_$name.log 2> gpurun_out/c25_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c25_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c25_$name.err
}
BARGS="--global-batch 256"; run_bench b256_st VG_TC_TMA_STORE=0; run_bench b256_tma VG_TC_TMA_STORE=1; run_bench b256_st_2 VG_TC_TMA_STORE=0; run_bench b256_tma_2 VG_TC_TMA_STORE=1
BARGS="--global-batch 32"; run_bench b32_st VG_TC_TMA_STORE=0; run_bench b32_tma VG_TC_TMA_STORE=1
VG_TC_TMA_STORE=0 timeout 300 python scripts/sweep_conv.py 64 > gpurun_out/c25_sweep_st.txt 2>&1; tail -1 gpurun_out/c25_sweep_st.txt
VG_TC_TMA_STORE=1 timeout 300 python scripts/sweep_conv.py 64 > gpurun_out/c25_sweep_tma.txt 2>&1; tail -1 gpurun_out/c25_sweep_tma.txt
