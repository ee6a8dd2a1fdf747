"""Data-parallel equivalence check (run under torchrun, one rank per GPU):
the N-rank trainer (SyncBN statistics + gradient all-reduce, batch split over ranks) must reproduce
the single-process trainer on the GLOBAL batch.  Prints one JSON line on rank 0
and exits non-zero on mismatch."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import vae_gan_b200 as V


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    results = {}
    ok = True
    # (compute dtype, loss mode, optimizer): north_star's BCE + Adam in both precisions, and the notebook's own
    # WGAN-GP + RMSprop + clamp (double backward with SyncBN sums exchanged in the second-order pass too)
    for cdt, loss_mode, opt in ((torch.float32, "bce", "adam"), (torch.bfloat16, "bce", "adam"), (torch.float32, "wgan_gp", "rmsprop")):
        # 16 samples per rank: at 4 per rank single LeakyReLU kink flips in the 4-Linear head moved whole rows of D's bf16
        # gradient by O(1) (rel-L2 7.8e-2 between partitions in round 1); they average out with the batch
        B, S, fs, steps = 16 * world, 32, 64, 2
        gen = torch.Generator().manual_seed(5)
        xs = [torch.rand(B, 1, S, S, generator=gen).to(dev) for _ in range(steps)]

        def make(pg):
            torch.manual_seed(0)
            G, D = V.build_vae_gan(feature_size=fs, image_size=S)
            G, D = G.to(dev).train(), D.to(dev).train()
            V.rng.seed = 77
            V.rng.step_tensor(dev).zero_()
            V.config.process_group = None
            V.config.sample_offset = 0
            return V.VaeGanTrainer(G, D, process_group=pg, loss_mode=loss_mode, optimizer=opt)

        with V.compute_dtype(cdt):
            single = make(None)
            ref = []
            g_ref_g = g_ref_d = None
            for x in xs:
                single.step(x)
                ref.append(single.read_losses())
                if g_ref_g is None:          # gradients of the FIRST step: computed before any divergence can build up
                    g_ref_g, g_ref_d = single.fg.g.clone(), single.fd.g.clone()
            p_ref_g, p_ref_d = single.fg.p.clone(), single.fd.p.clone()
            dp = make(dist.group.WORLD)
            got = []
            lb = B // world
            g_dp_g = g_dp_d = None
            for x in xs:
                dp.step(x[rank * lb:(rank + 1) * lb].contiguous())
                got.append(dp.read_losses())
                if g_dp_g is None:
                    g_dp_g, g_dp_d = dp.fg.g.clone(), dp.fd.g.clone()
            V.config.process_group = None
            # step-1 quantities computed BEFORE any optimizer update must agree tightly; everything
            # after an update inherits Adam's lr*sign(g) amplification of summation-order noise
            pre = ("d_loss", "real_loss", "fake_loss", "recon", "kl", "gp")
            worst = worst_pre = 0.0
            for i, (a, b) in enumerate(zip(got, ref)):
                for k in b:
                    e = abs(a[k] - b[k]) / max(1.0, abs(b[k]))
                    if i == 0 and k in pre:
                        worst_pre = max(worst_pre, e)
                    else:
                        worst = max(worst, e)
            lr = 3e-4
            frac_g = float(((dp.fg.p - p_ref_g).abs() > 0.5 * lr).float().mean())
            frac_d = float(((dp.fd.p - p_ref_d).abs() > 0.5 * lr).float().mean())
            # first-step gradients (after the all-reduce every rank holds the global gradient); D's is taken before
            # any update, G's after the first D update
            gl2_g = float((g_dp_g - g_ref_g).norm() / g_ref_g.norm().clamp_min(1e-30))
            gl2_d = float((g_dp_d - g_ref_d).norm() / g_ref_d.norm().clamp_min(1e-30))
            # all ranks must hold identical parameters
            chk = torch.stack([dp.fg.p.double().sum(), dp.fd.p.double().sum()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(((hi - lo).abs() <= 1e-6 * hi.abs().clamp_min(1)).all())
        results[f"{cdt}/{loss_mode}/{opt}"] = dict(peer_syncbn=dp.peer is not None, grad_buckets=(len(dp.buckets_d.ranges) if dp.buckets_d is not None else 0), worst_pre_update_loss_rel=worst_pre, worst_loss_rel=worst, frac_params_off_g=frac_g, frac_params_off_d=frac_d, grad_rel_l2_g=gl2_g, grad_rel_l2_d=gl2_d,
                                                       replicas_identical=same)
        lim = 2e-3 if cdt == torch.float32 else 0.08
        tol_pre = 2e-5 if cdt == torch.float32 else 2e-2
        tol_post = 5e-3 if cdt == torch.float32 else 5e-2
        # Stated bounds on the first-step gradient rel-L2 between the 2-rank and the 1-rank run.  fp32: 2e-3.  bf16: G 5e-3;
        # D 1e-1 - D's gradient at random initialisation amplifies bf16 rounding noise to ~1e-1 against the exact gradient
        # (tests/test_gpu_round2.py::test_train_step_bs64_bf16_flat_tolerance: ours and torch's bf16 autocast both), and a
        # different partition de-correlates the roundings (different fp32 summation order -> a few bf16 ties break the
        # other way -> saturates at the bf16 noise floor within a few layers), so two bf16 runs differ by that much
        gtol = 2e-3 if cdt == torch.float32 else 1e-1
        gtol_g = 2e-3 if cdt == torch.float32 else 5e-3
        if opt == "rmsprop":
            # RMSprop's first steps move EVERY element by ~10*lr*sign(g) (v = 0.01 g^2) and the clamp keeps all of D
            # within +-0.01, so elements whose gradient is summation-order noise flip freely in any implementation
            # and later steps diverge chaotically; judge the first-step gradients and D's parameters instead
            params_ok = gl2_g <= gtol_g and gl2_d <= gtol and frac_d <= lim
        else:
            params_ok = frac_g <= lim and frac_d <= lim and gl2_d <= gtol and gl2_g <= gtol_g
        ok = ok and worst_pre <= tol_pre and worst <= tol_post and params_ok and same
    if rank == 0:
        print(json.dumps(dict(world=world, ok=ok, **results)))
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
