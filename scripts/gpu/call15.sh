#!/bin/bash
# round-2 GPU call 15 (8 GPUs): the N=8 bench line (and cfg4) with programmatic dependent launch and the tile table
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tr_run() { # name nproc extra-args...
  local name=$1; shift; local n=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu-baseline --skip-lib-baseline "$@" > gpurun_out/c15_$name.log 2> gpurun_out/c15_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c15_$name.log') if l.startswith('{')][-1]);print('$name',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c15_$name.err
}
tr_run n8 8
tr_run n8_cfg4 8 --workload cfg4 --steps 10
