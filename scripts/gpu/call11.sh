#!/bin/bash
# round-2 GPU call 11: deterministic mode with per-split scratch slabs (tests + cost), ncu --set full of the low-efficiency kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deterministic.py -x -q > gpurun_out/c11_det_tests.log 2>&1; echo "det tests rc=$?"; tail -5 gpurun_out/c11_det_tests.log
run_bench() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline $BARGS > gpurun_out/c11_$name.log 2> gpurun_out/c11_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c11_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c11_$name.err
}
BARGS="--global-batch 256"
run_bench b256_det VG_DETERMINISTIC=1
BARGS="--global-batch 32"
run_bench b32_det VG_DETERMINISTIC=1
BARGS="--global-batch 64"
run_bench b64_det VG_DETERMINISTIC=1
timeout 600 python scripts/ncu_kernels.py > gpurun_out/c11_ncu_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/c11_ncu_plain.log
timeout 600 ncu --set full --clock-control none --profile-from-start off -o /tmp/c11_ncu python scripts/ncu_kernels.py > gpurun_out/c11_ncu.log 2>&1
ls -la /tmp/c11_ncu.ncu-rep; tail -2 gpurun_out/c11_ncu.log
ncu -i /tmp/c11_ncu.ncu-rep --page raw --csv > gpurun_out/c11_ncu_raw.csv 2>/dev/null; ls -la gpurun_out/c11_ncu_raw.csv
python scripts/ncu_traffic.py bn_act_bwd_apply=/tmp/c11_ncu.ncu-rep:bn_stream_kernel:64
cp profiles/r2_ncu_traffic.json gpurun_out/c11_ncu_traffic.json
