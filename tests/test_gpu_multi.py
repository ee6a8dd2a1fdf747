"""N>1 on real GPUs (-m gpu; skipped unless >= 2 devices are visible): the data-parallel trainer
(SyncBN + gradient all-reduce over NCCL, batch split over ranks) reproduces the single-process
trainer on the global batch.  Runs scripts/ddp_check.py under torchrun."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_trainer_matches_global_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "scripts" / "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=str(ROOT))
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert line, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(line[-1])
    assert res["ok"], res
