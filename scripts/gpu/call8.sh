#!/bin/bash
# round-2 GPU call 8 (8 GPUs): 1/2/4/8 scaling on ONE box, cfg4 at 8 GPUs, A/B of the overlap and the peer SyncBN
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpus.txt
tr_run() { # name nproc extra-args...
  local name=$1; shift; local n=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 30 --warmup 5 --skip-cpu-baseline --skip-lib-baseline "$@" > gpurun_out/c8_$name.log 2> gpurun_out/c8_$name.err
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c8_$name.log') if l.startswith('{')][-1]);print('$name',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/c8_$name.err
}
tr_run n8 8
tr_run n8_cfg4 8 --workload cfg4 --steps 10
tr_run n4 4
tr_run n2 2
timeout 300 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c8_n1.log 2> gpurun_out/c8_n1.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c8_n1.log') if l.startswith('{')][-1]);print('n1',d['n_gpus'],d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'])"
VG_GRAD_BUCKETS=0 tr_run n8_nobuckets 8
VG_PEER_SYNCBN=0 tr_run n8_nccl_syncbn 8
VG_DIAG_NO_SYNCBN=1 tr_run n8_diag_nosyncbn 8
VG_DIAG_NO_GRAD_AR=1 tr_run n8_diag_noar 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_check.py > gpurun_out/c8_ddp_check.log 2>&1; grep '^{' gpurun_out/c8_ddp_check.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ddp_check ok =', d['ok'])"
