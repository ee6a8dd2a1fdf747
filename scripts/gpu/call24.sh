#!/bin/bash
# round-2 GPU call 24: documentation lines - deterministic mode and the fp32 parity path at the BASELINE batch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
VG_DETERMINISTIC=1 timeout 400 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c24_det_b256.log 2> gpurun_out/c24_det_b256.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c24_det_b256.log') if l.startswith('{')][-1]);print('det b256',d['value'],d['ms_per_step'],d['launches_per_step'])" || tail -3 gpurun_out/c24_det_b256.err
timeout 600 python bench.py --fp32 --global-batch 64 --steps 5 --warmup 3 --skip-cpu-baseline --skip-lib-baseline > gpurun_out/c24_fp32_b64.log 2> gpurun_out/c24_fp32_b64.err
python -c "import json;d=json.loads([l for l in open('gpurun_out/c24_fp32_b64.log') if l.startswith('{')][-1]);print('fp32 b64',d['value'],d['ms_per_step'],d['launches_per_step'])" || tail -3 gpurun_out/c24_fp32_b64.err
