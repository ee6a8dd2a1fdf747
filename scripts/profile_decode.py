"""Decoder-only sampling (BASELINE config 5) bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
usage: python scripts/profile_decode.py [batch] [folded|module]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vae_gan_b200 as V
from vae_gan_b200.sampling import FoldedGenerator

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "folded"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with V.compute_dtype(torch.bfloat16), torch.no_grad():
    G, _ = V.build_vae_gan(feature_size=64, image_size=96)
    G = G.to(dev).eval()
    G.set_is_training(False)
    fn = FoldedGenerator(G).decode if which == "folded" else G.decode
    z = torch.randn(batch, 256, 24, 24, device=dev)
    for _ in range(2):
        fn(z)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = fn(z)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(which, batch, tuple(out.shape), float(out.float().mean()))
