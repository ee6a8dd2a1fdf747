"""Convolution time vs reduction length at a fixed output size: 3x3, 96x96, batch 64."""
import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch
import vae_gan_b200.functional as VF
from vae_gan_b200 import _lib
dev = torch.device("cuda", 0)
B = 64
g = torch.Generator().manual_seed(0)
def timeit(fn, iters=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return sum(ts[1:-1]) / (len(ts) - 2)
for cin, cout, h in [(64, 128, 96), (128, 128, 96), (256, 128, 96), (512, 128, 96), (128, 256, 96), (256, 256, 96), (512, 256, 96), (64, 64, 96), (128, 64, 96), (256, 64, 96)]:
    geom = VF.ConvGeom(3, 1, 1, False)
    x = VF.as_act(torch.randn(B, cin, h, h, generator=g).to(dev), torch.bfloat16)
    w = (torch.randn((cout, cin, 3, 3), generator=g) / (cin * 9) ** 0.5).to(dev)
    d, ho, wo = VF._conv_desc(x.shape, cout, geom, torch.bfloat16, torch.bfloat16)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev); pn = torch.empty_like(pk)
    s = _lib.stream_ptr()
    _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), None, pk.data_ptr(), pn.data_ptr(), s)
    y = VF.empty_act(B, cout, ho, wo, torch.bfloat16, dev)
    t = timeit(lambda: _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), pk.data_ptr(), pn.data_ptr(), None, None, y.data_ptr(), None, s))
    fl = 2.0 * B * ho * wo * cin * cout * 9
    print(f"{cin:4d}->{cout:4d} @{h}: {t*1e3:8.1f} us  {fl/t/1e9:7.0f} TF/s   k-blocks/item {cin//64*9}")
    del x, y
