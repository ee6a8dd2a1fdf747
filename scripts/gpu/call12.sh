#!/bin/bash
# round-2 GPU call 12: autotune the tile table for the BASELINE shapes, then A/B the step with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python scripts/autotune_baseline.py gpurun_out/c12_tile_table.json > gpurun_out/c12_autotune.log 2>&1; echo "autotune rc=$?"; grep -E "^==|entries" gpurun_out/c12_autotune.log
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c12_$name.log 2> gpurun_out/c12_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c12_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c12_$name.err
}
T=$PWD/gpurun_out/c12_tile_table.json
BARGS="--global-batch 32";  run_bench b32_heur VG_TILE_TABLE=0;  run_bench b32_table VG_TILE_TABLE=$T
BARGS="--global-batch 64";  run_bench b64_heur VG_TILE_TABLE=0;  run_bench b64_table VG_TILE_TABLE=$T
BARGS="--global-batch 128"; run_bench b128_heur VG_TILE_TABLE=0; run_bench b128_table VG_TILE_TABLE=$T
BARGS="--global-batch 256"; run_bench b256_heur VG_TILE_TABLE=0; run_bench b256_table VG_TILE_TABLE=$T
BARGS="--workload cfg4 --steps 10"; run_bench cfg4_heur VG_TILE_TABLE=0; run_bench cfg4_table VG_TILE_TABLE=$T
VG_TILE_TABLE=$T timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_kernels.py -x -q > gpurun_out/c12_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/c12_tests.log
