#!/bin/bash
# round-2 GPU call 19: full-size property tests (second version)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -s > gpurun_out/c19_full_size.log 2>&1; echo "full-size rc=$?"; tail -8 gpurun_out/c19_full_size.log; grep -n "^E  " gpurun_out/c19_full_size.log | head -12
