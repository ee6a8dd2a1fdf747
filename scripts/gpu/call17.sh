#!/bin/bash
# round-2 GPU call 17: programmatic dependent launch with a TAIL trigger (launch_dependents when a CTA has issued its
# last loads / MMAs) in the bulk-copy BatchNorm and tcgen05 kernels, A/B against the default build; a few parity tests under it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c17_$name.log 2> gpurun_out/c17_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c17_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c17_$name.err
}
L=$PWD/vae_gan_b200/lib/libvaegan_sm100_tailtrig.so
BARGS="--global-batch 32";  run_bench b32_default A=1;  run_bench b32_tail VG_LIB=$L; run_bench b32_default_2 A=1;  run_bench b32_tail_2 VG_LIB=$L
BARGS="--global-batch 256"; run_bench b256_default A=1; run_bench b256_tail VG_LIB=$L
BARGS="--global-batch 64"; run_bench b64_default A=1; run_bench b64_tail VG_LIB=$L
VG_LIB=$L timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_deterministic.py -x -q > gpurun_out/c17_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/c17_tests.log
