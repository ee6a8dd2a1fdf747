#!/bin/bash
# round-2 GPU call 14 (2 GPUs): multi-GPU tests, 2-rank equivalence check and an N=2 bench line with programmatic
# dependent launch and the tile table in place
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c14_gpus.txt
# tests/test_gpu_multi.py runs exactly scripts/ddp_check.py under torchrun; run it directly to keep its log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_check.py > gpurun_out/c14_ddp_check.log 2>&1; grep '^{' gpurun_out/c14_ddp_check.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ddp_check ok =', d['ok'])"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c14_n2.log 2> gpurun_out/c14_n2.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c14_n2.log') if l.startswith('{')][-1]);print('n2',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c14_n2.err
VG_DETERMINISTIC=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c14_n2_det.log 2> gpurun_out/c14_n2_det.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c14_n2_det.log') if l.startswith('{')][-1]);print('n2 det',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'])" || tail -5 gpurun_out/c14_n2_det.err
