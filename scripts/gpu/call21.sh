#!/bin/bash
# round-2 GPU call 21: BatchNorm streaming-kernel launch parameters at the BASELINE batch (CTAs per SM, reverse walk)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c21_$name.log 2> gpurun_out/c21_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c21_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c21_$name.err
}
BARGS="--global-batch 256"
run_bench b256_default A=1
run_bench b256_ctas1 VG_BN_STREAM_CTAS=1
run_bench b256_ctas3 VG_BN_STREAM_CTAS=3
run_bench b256_norev VG_BN_REVERSE=0
run_bench b256_default_2 A=1
