#!/bin/bash
# round-2 GPU call 16: the whole GPU test-suite on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c16_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/c16_tests.log
