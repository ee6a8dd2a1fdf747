"""world_size-2 `gloo` test (CPU): the data-parallel contract the product implements.

Each rank runs the ORACLE on its half of the batch with (a) SyncBN - per-channel sum / sum-of-squares
all-reduced, exactly what `functional.bn_batch_stats` / `_bn_backward` exchange - (b) the mean-type
losses normalised by the GLOBAL element count while KL stays a sum (`VgLossDesc.n_pix_global`), and
(c) an all-reduce(SUM) of the gradients.  The result must equal the single-process global-batch
gradients, and the host-side partition helpers must give partition-invariant dropout masks."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _sync_batch_norm(x, P, pre, training):
    from oracle import vaegan_oracle as O
    import torch.distributed.nn.functional as dfn
    assert training
    c = x.shape[1]
    n_local = x.numel() // c
    s = torch.stack([x.sum((0, 2, 3)), (x * x).sum((0, 2, 3))])
    s = dfn.all_reduce(s, op=dist.ReduceOp.SUM)            # differentiable all-reduce
    n = n_local * dist.get_world_size()
    mean = s[0] / n
    var = s[1] / n - mean * mean
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + O.BN_EPS)
    return xh * P[pre + ".weight"][None, :, None, None] + P[pre + ".bias"][None, :, None, None]


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import vaegan_oracle as O
    from tests.gpu_util import generator_masks
    B, S, fs = 4, 16, 8
    spec = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    P0 = O.make_generator_params(spec, seed=3, dtype=torch.float64)
    gen = torch.Generator().manual_seed(1)
    for k in P0:                                            # non-trivial BN affine parameters
        if k.endswith((".weight", ".bias")) and P0[k].dim() == 1:
            P0[k] = P0[k] + 0.1 * torch.randn(P0[k].shape, generator=gen, dtype=torch.float64)
    x = torch.rand(B, 1, S, S, generator=gen, dtype=torch.float64)
    eps = torch.randn(B, spec.feature_depth, S // 4, S // 4, generator=gen, dtype=torch.float64)
    lb = B // world
    # partition-invariant masks: rank r draws the slice of the global mask that belongs to its samples
    g_masks, _ = generator_masks(spec, B, S, seed=9)
    l_masks, _ = generator_masks(spec, lb, S, seed=9, sample_offset=rank * lb)
    for k in g_masks:
        assert torch.equal(l_masks[k], g_masks[k][rank * lb:(rank + 1) * lb]), k

    def loss_fn(P, xb, eb, masks, n_pix_global):
        y, mu, lv = O.generator_forward(xb, P, spec, True, True, eb, masks)
        d = y - xb
        recon = (d.abs().sum() + (d * d).sum()) / n_pix_global      # means over the GLOBAL batch
        return 10 * recon + 0.1 * O.kl_divergence(mu, lv)            # KL is a sum (README.md:824)

    # single process, global batch
    Pg = O.clone_params(P0, requires_grad=True)
    keys = O.trainable_keys(Pg)
    want = torch.autograd.grad(loss_fn(Pg, x, eps, g_masks, x.numel()), [Pg[k] for k in keys])
    # data parallel
    orig = O.batch_norm
    O.batch_norm = _sync_batch_norm
    try:
        Pl = O.clone_params(P0, requires_grad=True)
        sl = slice(rank * lb, (rank + 1) * lb)
        got = torch.autograd.grad(loss_fn(Pl, x[sl], eps[sl], l_masks, x.numel()), [Pl[k] for k in keys])
    finally:
        O.batch_norm = orig
    worst = 0.0
    for k, g, w in zip(keys, got, want):
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        denom = float(w.abs().max())
        if denom > 1e-12:
            worst = max(worst, float((g - w).abs().max()) / denom)
    out_q.put((rank, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_contract_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst in res:
        assert worst < 1e-9, f"rank {rank}: summed DP gradients differ from the global-batch gradients by {worst:.2e}"


def _worker_gp(rank, world, port, out_q):
    """Gradient penalty under data parallelism (vae_gan_b200/train.py, loss_mode="wgan_gp"): every rank differentiates
    the sum of ITS critic outputs with respect to ITS interpolates, the SyncBN all-reduces (forward sums, backward sums
    and - in the second-order pass - their derivatives) supply the cross-rank terms, the local mean penalty is scaled
    by 1/world, and the parameter gradients are summed."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import vaegan_oracle as O
    from tests.gpu_util import discriminator_masks
    B, S, fs = 4, 16, 8
    spec = O.DiscriminatorSpec(1, fs, (1, 1, 1), (1, 2, 2), (2 * fs, 4 * fs, 8 * fs), input_size=S)
    P0 = O.make_discriminator_params(spec, seed=21, dtype=torch.float64)
    gen = torch.Generator().manual_seed(2)
    for k in P0:
        if k.endswith((".weight", ".bias")) and P0[k].dim() == 1 and "bn" in k:
            P0[k] = P0[k] + 0.1 * torch.randn(P0[k].shape, generator=gen, dtype=torch.float64)
    real = torch.rand(B, 1, S, S, generator=gen, dtype=torch.float64)
    fake = torch.rand(B, 1, S, S, generator=gen, dtype=torch.float64)
    alpha = torch.rand(B, 1, 1, 1, generator=gen, dtype=torch.float64)
    lb = B // world
    g_masks, _ = discriminator_masks(spec, B, 13, 0)
    l_masks, _ = discriminator_masks(spec, lb, 13, 0, sample_offset=rank * lb)
    # single process, global batch
    Pg = O.clone_params(P0, requires_grad=True)
    keys = O.trainable_keys(Pg)
    gp_g = O.gradient_penalty(Pg, spec, real, fake, alpha, g_masks)
    want = torch.autograd.grad(gp_g, [Pg[k] for k in keys], allow_unused=True)
    # data parallel
    orig = O.batch_norm
    O.batch_norm = _sync_batch_norm
    try:
        Pl = O.clone_params(P0, requires_grad=True)
        sl = slice(rank * lb, (rank + 1) * lb)
        gp_l = O.gradient_penalty(Pl, spec, real[sl], fake[sl], alpha[sl], l_masks) / world
        got = torch.autograd.grad(gp_l, [Pl[k] for k in keys], allow_unused=True)
    finally:
        O.batch_norm = orig
    tot = gp_l.detach().clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    worst = abs(float(tot) - float(gp_g)) / max(abs(float(gp_g)), 1e-12)
    for k, g, w in zip(keys, got, want):
        if w is None:
            continue
        g = (g if g is not None else torch.zeros_like(w)).clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        denom = float(w.abs().max())
        if denom > 1e-12:
            worst = max(worst, float((g - w).abs().max()) / denom)
    out_q.put((rank, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_penalty_data_parallel_contract_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker_gp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst in res:
        assert worst < 1e-8, f"rank {rank}: DP gradient penalty / its summed gradients differ from the global batch by {worst:.2e}"


# ----------------------------------------------------------------------------------------------------------------------
# bucketed, backward-overlapped gradient all-reduce (train.GradBuckets): host logic under gloo, world_size 2
# ----------------------------------------------------------------------------------------------------------------------
def _bucket_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn
    from vae_gan_b200.train import FlatParams, GradBuckets
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(300, 400), nn.Linear(400, 50), nn.Linear(50, 2000), nn.Linear(2000, 3))
    flat = FlatParams(net)
    # 0.25 MB buckets over ~0.9 MB of parameters: several buckets, the 400x300 weight alone exceeds one
    gb = GradBuckets(flat, dist.group.WORLD, None, bucket_mb=0.25)
    covered = sorted(gb.ranges)
    ok = covered[0][0] == 0 and covered[-1][1] == flat.total and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    ok = ok and all(gb.ranges[i][0] >= gb.ranges[i + 1][1] for i in range(len(gb.ranges) - 1))     # launch order = reverse parameter order
    params = list(net.parameters())
    results = []
    for uses in (1, 2):                          # 2 = the discriminator step: D(real) and D(fake) both contribute
        gb.begin()
        for p in params:
            for _ in range(uses):
                gb.use(p)
        g_local = torch.arange(flat.total, dtype=torch.float32) * (rank + 1) * 1e-3
        flat.g.copy_(g_local)
        early = 0
        for u in range(uses):                    # backward: last layer first; the second pass completes the counts
            for p in reversed(params):
                before = len(gb.order)
                gb.done(p)
                early += len(gb.order) - before
        sent_before_flush = len(gb.order)
        gb.flush()
        want = torch.arange(flat.total, dtype=torch.float32) * 1e-3 * sum(r + 1 for r in range(world))
        results.append((bool(torch.allclose(flat.g, want, rtol=1e-6, atol=0)), sent_before_flush, len(gb.ranges), list(gb.order)))
    if rank == 0:
        out_q.put((ok, results))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_buckets_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 500) + 600
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, results = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok, "bucket ranges must tile the flat buffer in reverse parameter order"
    for summed, sent_early, n_buckets, order in results:
        assert summed, "all-reduced gradient != sum over ranks"
        assert n_buckets >= 3
        assert sent_early == n_buckets, "every bucket should have been launched from inside the backward"
        assert order == list(range(n_buckets)), f"buckets launched out of backward order: {order}"
