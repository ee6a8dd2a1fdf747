#!/bin/bash
# round-2 GPU call 28: batch-64 launch list of the last build (TMA-store epilogue)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python scripts/profile_step.py 64 > gpurun_out/c28_step_b64.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/c28_launches_b64.csv python scripts/profile_step.py 64 > gpurun_out/c28_ncu_b64.log 2>&1
python scripts/summarize_launches.py gpurun_out/c28_launches_b64.csv > gpurun_out/c28_launches_b64_summary.txt 2>&1; head -8 gpurun_out/c28_launches_b64_summary.txt
