#!/bin/bash
# round-2 GPU call 2: batched pack / spectral norm, sigma in the epilogue, RED.v4 split-K, N3 pipeline
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in test_gpu_round2 test_gpu_kernels test_gpu_modules; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -x -q -s > gpurun_out/c2_$f.log 2>&1; echo "$f rc=$?" | tee -a gpurun_out/c2_$f.log
  grep -E "passed|failed|error" gpurun_out/c2_$f.log | tail -3
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c2_smoke.log 2>&1; tail -2 gpurun_out/c2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c2_bench.log 2> gpurun_out/c2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ["gpurun_out/c2_bench.log"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["launches_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
for b in 64 32; do
  timeout 600 python bench.py --steps 20 --warmup 5 --global-batch $b --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c2_bench_b$b.log 2> gpurun_out/c2_bench_b$b.err
  python -c "import json;d=json.loads(open('gpurun_out/c2_bench_b$b.log').read().strip().splitlines()[-1]);print('b$b',d['value'],d['ms_per_step'],d['launches_per_step'])"
done
VG_PACK_CACHE=0 timeout 600 python bench.py --steps 20 --warmup 5 --global-batch 32 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c2_bench_b32_nocache.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/c2_bench_b32_nocache.log').read().strip().splitlines()[-1]);print('b32 nocache',d['value'],d['ms_per_step'],d['launches_per_step'])"
for b in 256 64; do
VG_TC_FUSE_STATS=0 timeout 600 python bench.py --steps 20 --warmup 5 --global-batch $b --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c2_bench_b${b}_nofuse.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/c2_bench_b${b}_nofuse.log').read().strip().splitlines()[-1]);print('b$b nofuse',d['value'],d['ms_per_step'],d['launches_per_step'])"
done
timeout 300 python scripts/sweep_conv.py > gpurun_out/c2_sweep_conv.txt 2>&1; tail -20 gpurun_out/c2_sweep_conv.txt
for b in 64 32; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c2_launches_b$b.csv python scripts/profile_step.py $b > gpurun_out/c2_ncu_b$b.log 2>&1
python scripts/summarize_launches.py gpurun_out/c2_launches_b$b.csv > gpurun_out/c2_launches_b${b}_summary.txt 2>&1
done
head -45 gpurun_out/c2_launches_b64_summary.txt
