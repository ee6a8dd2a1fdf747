#!/bin/bash
# round-2 GPU call 3 (2 GPUs): data-parallel equivalence with bucketed overlapped all-reduce, N=2 bench; on GPU 0: epilogue-warp variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c3_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/c3_multi.log 2>&1; echo "multi rc=$?"; tail -5 gpurun_out/c3_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/ddp_check.py > gpurun_out/c3_ddp_check.log 2>&1; echo "ddp rc=$?"; tail -3 gpurun_out/c3_ddp_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c3_bench_n2.log 2> gpurun_out/c3_bench_n2.err; echo "bench n2 rc=$?"
VG_GRAD_BUCKETS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c3_bench_n2_nobuckets.log 2> gpurun_out/c3_bench_n2_nobuckets.err
python - <<'PY'
import json
for f in ["c3_bench_n2", "c3_bench_n2_nobuckets"]:
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.log") if l.startswith("{")][-1]); print(f, d["value"], d["ms_per_step"], d["launches_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/c3_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "stock_optimizer or input_pipeline" > gpurun_out/c3_round2.log 2>&1; tail -2 gpurun_out/c3_round2.log
for ew in 4 8 82; do
  VG_TC_EW64=$ew timeout 600 python bench.py --steps 20 --warmup 5 --global-batch 64 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c3_bench_b64_ew$ew.log 2>&1
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c3_bench_b64_ew$ew.log') if l.startswith('{')][-1]);print('b64 ew$ew',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['e2e_u8_pipeline']['value'])"
done
VG_TC_EW64=8 timeout 300 python scripts/sweep_conv.py > gpurun_out/c3_sweep_conv_ew8.txt 2>&1; grep -E "64->64|128->64|sum" gpurun_out/c3_sweep_conv_ew8.txt
VG_TC_EW64=82 timeout 300 python scripts/sweep_conv.py > gpurun_out/c3_sweep_conv_ew82.txt 2>&1; grep -E "64->64|128->64|sum" gpurun_out/c3_sweep_conv_ew82.txt
