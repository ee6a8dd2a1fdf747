#!/bin/bash
# round-2 GPU call 23 (4 GPUs): the N=4 bench line on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29644 bench.py --gpus 4 --steps 30 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c23_n4.log 2> gpurun_out/c23_n4.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c23_n4.log') if l.startswith('{')][-1]);print('n4',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c23_n4.err
