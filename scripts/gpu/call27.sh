#!/bin/bash
# round-2 GPU call 27: final build (TMA-store epilogue default): whole GPU suite, smoke, the default bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c27_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c27_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c27_smoke.log 2>&1; tail -1 gpurun_out/c27_smoke.log
timeout 900 python bench.py > gpurun_out/c27_bench.log 2> gpurun_out/c27_bench.err; echo "bench rc=$?"
python -c "import json;d=json.loads([l for l in open('gpurun_out/c27_bench.log') if l.startswith('{')][-1]);print('b256',d['value'],d['ms_per_step'],d['launches_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['vs_gpu_lib'], d['step_frac_of_ideal'], d['clocks']['sm_mhz'])"
