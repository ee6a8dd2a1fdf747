"""Autotune the tensor-core convolution tiles for the BASELINE configurations on this GPU and write
vae_gan_b200/tile_table.json (loaded at import).  usage: python scripts/autotune_baseline.py [out.json] [--quick]
Per-GPU batches covered: 256 / 128 / 64 / 32 at 96x96, feature size 64 (BASELINE configs 1-3 on 1 / 2 / 4 / 8 GPUs) and 16 at
256x256 with the widths doubled (config 4)."""
import os
import sys
from pathlib import Path
os.environ["VG_TILE_TABLE"] = "0"           # measure against the heuristics, not against an older table
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vae_gan_b200 as V
from vae_gan_b200 import tune

out = next((a for a in sys.argv[1:] if not a.startswith("--")), str(tune.DEFAULT_TABLE))
quick = "--quick" in sys.argv
dev = torch.device("cuda", 0)
table = {}
configs = [(64, 96, b) for b in ((32, 256) if quick else (256, 128, 64, 32))] + ([] if quick else [(128, 256, 16)])
with V.compute_dtype(torch.bfloat16):
    for fs, S, B in configs:
        torch.manual_seed(0)
        G, D = V.build_vae_gan(feature_size=fs, image_size=S)
        G, D = G.to(dev).train(), D.to(dev).train()
        tr = V.VaeGanTrainer(G, D)
        x = torch.rand(B, 1, S, S, device=dev)
        tr.step(x)
        keys = tune.record(lambda: tr.step(x))
        del tr, G, D
        torch.cuda.empty_cache()
        print(f"== feature_size {fs}, {S}x{S}, batch {B}: {len(keys)} distinct (shape, direction) keys", flush=True)
        t = tune.autotune(keys=keys, verbose=True)
        table.update(t)
        tune.clear()
        gain = sum(v["heuristic_ms"] - v["ms"] for v in t.values())
        print(f"   {len(t)} entries, {gain * 1e3:.0f} us per pass over the distinct shapes", flush=True)
tune.save(table, out, meta={"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "min_gain": 0.03})
print("wrote", out, len(table), "entries")
