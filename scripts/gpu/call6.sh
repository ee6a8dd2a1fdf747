#!/bin/bash
# round-2 GPU call 6: streaming degenerate wgrad, separate inference epilogue instantiations, hybrid folded sampler
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for f in test_gpu_kernels test_gpu_round2 test_gpu_modules; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -s > gpurun_out/c6_$f.log 2>&1; echo "$f rc=$?" | tee -a gpurun_out/c6_$f.log
  grep -E "passed|failed|error" gpurun_out/c6_$f.log | tail -3
done
grep -E "^\[folded\]|^FAILED|^E  " gpurun_out/c6_test_gpu_*.log | head -20
run_bench() { local name=$1; shift; local b=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --global-batch $b --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c6_bench_$name.log 2>&1
  python -c "import json;d=json.loads([l for l in open('gpurun_out/c6_bench_$name.log') if l.startswith('{')][-1]);print('$name',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['e2e_u8_pipeline']['value'])"
}
run_bench b256 256 A=1
run_bench b64 64 A=1
run_bench b32 32 A=1
run_bench b64_nodeg 64 VG_DEG_STREAM=0
timeout 600 python bench.py --workload decode > gpurun_out/c6_decode.log 2> gpurun_out/c6_decode.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/c6_decode.log') if l.startswith('{')][-1])
for r in d['sweep']: print(r)
print(d['encode_b256'], d['reconstruct_b256'], d['frac_of_tensor_peak'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c6_launches_b64.csv python scripts/profile_step.py 64 > gpurun_out/c6_ncu_b64.log 2>&1
python scripts/summarize_launches.py gpurun_out/c6_launches_b64.csv > gpurun_out/c6_launches_b64_summary.txt 2>&1
head -24 gpurun_out/c6_launches_b64_summary.txt
