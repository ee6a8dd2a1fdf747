"""Run the two kernels bench.py's `roofline` / `roofline_hbm` entries name, alone, a few times (for `ncu --set full`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench

dev = torch.device("cuda", 0)
peaks = bench.measured_peaks()
torch.cuda.profiler.start()
print(bench.roofline_probe(dev, 64, peaks))
print(bench.roofline_hbm_probe(dev, 64, peaks))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
