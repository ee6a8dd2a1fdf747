"""One VAE-GAN training iteration bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
usage: python scripts/profile_step.py [batch] [--fp32]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vae_gan_b200 as V

batch = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with V.compute_dtype(torch.float32 if "--fp32" in sys.argv else torch.bfloat16):
    G, D = V.build_vae_gan(feature_size=64, image_size=96)
    G, D = G.to(dev).train(), D.to(dev).train()
    tr = V.VaeGanTrainer(G, D)
    x = torch.rand(batch, 1, 96, 96, generator=torch.Generator().manual_seed(1)).to(dev)
    for _ in range(2):
        tr.step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.step(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("losses", tr.read_losses())
