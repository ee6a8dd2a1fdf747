#!/bin/bash
# round-2 GPU call 29: the TMA-store bit-equality test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_round2.py -q -k tma_store > gpurun_out/c29_test.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c29_test.log
