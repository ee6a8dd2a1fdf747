"""Isolated check of the NVLink peer all-reduce (run under torchrun, >= 2 GPUs)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.distributed as dist
from vae_gan_b200.dist import PeerExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
px = PeerExchange(dist.group.WORLD, dev, n_slots=16)
print(f"rank {rank}: peer ptrs {[hex(px.desc.peer_data[r] or 0) for r in range(world)]}", flush=True)
epoch = torch.zeros(1, dtype=torch.int64, device=dev)
ok = True
for step in range(1, 40):
    epoch += 1
    px.reset()
    for k in range(5):
        n = [2, 128, 1024, 2048, 256][k]
        v = torch.arange(n, dtype=torch.float64, device=dev) * (rank + 1) + step + k
        want = sum(torch.arange(n, dtype=torch.float64) * (r + 1) + step + k for r in range(world))
        px.allreduce_(v, epoch)
        torch.cuda.synchronize()
        if not torch.equal(v.cpu(), want):
            ok = False
            print(f"rank {rank} step {step} k {k}: MISMATCH max err {float((v.cpu()-want).abs().max())}", flush=True)
print(f"rank {rank}: {'OK' if ok else 'FAILED'}", flush=True)
dist.barrier()
os._exit(0 if ok else 1)
