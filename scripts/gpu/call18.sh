#!/bin/bash
# round-2 GPU call 18: full-size property tests, then the whole GPU suite on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_full_size.py -q -s > gpurun_out/c18_full_size.log 2>&1; echo "full-size rc=$?"; tail -12 gpurun_out/c18_full_size.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_full_size.py > gpurun_out/c18_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c18_tests.log
