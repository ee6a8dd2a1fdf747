#!/bin/bash
# round-2 GPU call 20: fused spectral-norm weight gradient (vg_conv_wgrad_sn): parity, whole suite, A/B of the step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "fused_spectral_norm or spectral" > gpurun_out/c20_sn_tests.log 2>&1; echo "sn tests rc=$?"; tail -3 gpurun_out/c20_sn_tests.log
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c20_$name.log 2> gpurun_out/c20_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c20_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c20_$name.err
}
BARGS="--global-batch 32";  run_bench b32_twocall VG_SN_FUSED_WGRAD=0;  run_bench b32_fused VG_SN_FUSED_WGRAD=1; run_bench b32_twocall_2 VG_SN_FUSED_WGRAD=0;  run_bench b32_fused_2 VG_SN_FUSED_WGRAD=1
BARGS="--global-batch 256"; run_bench b256_twocall VG_SN_FUSED_WGRAD=0; run_bench b256_fused VG_SN_FUSED_WGRAD=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c20_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c20_tests.log
