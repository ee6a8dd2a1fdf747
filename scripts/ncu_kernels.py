"""Run selected hot kernels once each inside a cudaProfiler range (for one `ncu --set full --profile-from-start off` pass):
wgrad / forward of the layers whose tensor-core efficiency is lowest, and the BatchNorm backward-apply streaming kernel.
usage: python scripts/ncu_kernels.py [batch]"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vae_gan_b200.functional as VF
from vae_gan_b200 import _lib
import bench

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
SHAPES = [("64->64 @96", 64, 64, 96, 3, 1, 1, False), ("128->128 @96", 128, 128, 96, 3, 1, 1, False),
          ("128->128 @48", 128, 128, 48, 3, 1, 1, False), ("256->256 @48", 256, 256, 48, 3, 1, 1, False),
          ("convT 128->64 @48", 128, 64, 48, 4, 2, 1, True)]
jobs = []
for name, cin, cout, h, k, st, pad, tr in SHAPES:
    geom = VF.ConvGeom(k, st, pad, tr)
    x = VF.as_act(torch.randn(B, cin, h, h, generator=g).to(dev), torch.bfloat16)
    w = (torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(dev)
    d, ho, wo = VF._conv_desc(x.shape, cout, geom, torch.bfloat16, torch.bfloat16)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev); pn = torch.empty_like(pk)
    s = _lib.stream_ptr()
    _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), None, pk.data_ptr(), pn.data_ptr(), s)
    y = VF.empty_act(B, cout, ho, wo, torch.bfloat16, dev)
    dy = VF.as_act(torch.randn(B, cout, ho, wo, generator=g).to(dev), torch.bfloat16)
    dw = torch.zeros_like(w)
    wsb = torch.empty(w.numel(), dtype=torch.float32, device=dev)
    keep = (x, w, pk, pn, y, dy, dw, wsb, d)
    jobs.append((name + " fwd", lambda d=d, x=x, pk=pk, pn=pn, y=y: _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), pk.data_ptr(), pn.data_ptr(), None, None, y.data_ptr(), None, s), keep))
    jobs.append((name + " wgrad", lambda d=d, x=x, dy=dy, dw=dw, wsb=wsb: _lib.call("vg_conv_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, wsb.data_ptr(), s), keep))
for _, fn, _k in jobs:
    fn(); fn()
torch.cuda.synchronize()
peaks = bench.measured_peaks()
torch.cuda.profiler.start()
for name, fn, _k in jobs:
    fn()
torch.cuda.synchronize()
print(bench.roofline_hbm_probe(dev, 64, peaks))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("order:", [j[0] for j in jobs])
