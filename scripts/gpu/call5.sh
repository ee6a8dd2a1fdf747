#!/bin/bash
# round-2 GPU call 5: folded sampler, batched pack/SN tests, decode sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in test_gpu_round2 test_gpu_kernels test_gpu_modules; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/c5_$f.log 2>&1; echo "$f rc=$?" | tee -a gpurun_out/c5_$f.log
  grep -E "passed|failed|error" gpurun_out/c5_$f.log | tail -3
done
grep -E "^\[folded\]|^FAILED|^E  " gpurun_out/c5_test_gpu_round2.log | head -20
timeout 600 python bench.py --workload decode > gpurun_out/c5_decode.log 2> gpurun_out/c5_decode.err; tail -c 2500 gpurun_out/c5_decode.log; tail -3 gpurun_out/c5_decode.err
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c5_bench.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/c5_bench.log') if l.startswith('{')][-1]);print('b256',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['e2e_u8_pipeline']['value'])"
