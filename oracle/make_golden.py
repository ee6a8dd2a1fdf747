"""Generate tests/golden/*.pt by running the REAL reference notebook classes.

Run inside the build container only (needs /root/reference):
    python oracle/make_golden.py

The weights are NOT stored: they come from `vaegan_oracle.make_*_params(seed)` and are
loaded into the reference modules with load_state_dict (which also proves state-dict key
compatibility).  Stored: inputs, injected randomness (dropout keep-masks, reparam noise,
GP alpha), and the reference's outputs / gradients / post-step parameters (large tensors
as summaries).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import vaegan_oracle as O          # noqa: E402
from oracle.load_reference import load_reference_namespace  # noqa: E402

GOLDEN = Path(__file__).resolve().parent.parent / "tests" / "golden"


def summarize(t: torch.Tensor):
    t = t.detach().to(torch.float32).cpu()
    if t.numel() <= 20000:
        return t.clone()
    flat = t.flatten()
    stride = max(1, flat.numel() // 4096)
    return dict(summary=True, shape=tuple(t.shape), sum=float(flat.double().sum()),
                asum=float(flat.double().abs().sum()), stride=stride,
                sample=flat[::stride][:4096].clone())


class MaskSeq(nn.Module):
    """Stands in for nn.Dropout/nn.Dropout2d: the i-th train-mode call multiplies by the
    i-th supplied keep-mask / (1-p)."""

    def __init__(self, masks, p=0.5):
        super().__init__()
        self.masks, self.p, self.calls = list(masks), p, 0

    def forward(self, x):
        if not self.training:
            return x
        m = self.masks[self.calls % len(self.masks)]
        self.calls += 1
        return x * (m.to(x.dtype) / (1.0 - self.p))


@contextlib.contextmanager
def patched_randn_like(seq):
    orig = torch.randn_like
    it = iter(seq)
    torch.randn_like = lambda t, *a, **k: next(it).to(t.dtype)
    try:
        yield
    finally:
        torch.randn_like = orig


def build_ref(ns, spec_g, spec_d, Pg, Pd):
    G = ns["UnsupervisedGeneratorNetwork"](
        encoder=ns["Encoder"](spec_g.in_channels, spec_g.depth, spec_g.length, spec_g.feature_size),
        decoder=ns["Decoder"](spec_g.feature_depth, spec_g.depth, spec_g.length,
                              spec_g.reconstruction_channels),
        code_processor=ns["SpatialVAECodeProcessor"](spec_g.feature_depth, True), is_vae=True)
    D = ns["Discriminator"](ns["ResBlockDiscriminator"], spec_d.num_stride_conv1,
                            spec_d.num_features_conv1, list(spec_d.num_blocks),
                            list(spec_d.num_strides_res), list(spec_d.num_features_res))
    D.linear_len = spec_d.linear_len                 # SURVEY.md D3 adapter (ref hard-codes 256)
    D.linear_1 = nn.Linear(spec_d.linear_len, 1024)
    G.load_state_dict(Pg, strict=True)
    D.load_state_dict(Pd, strict=True)
    return G, D


def g_blocks(G):
    return {"encoder.encoder." + n: m for n, m in G.encoder.encoder.named_children()} | \
           {"decoder.decoder." + n: m for n, m in G.decoder.decoder.named_children()}


def d_blocks(D):
    out = {}
    for i, layer in enumerate(D.res_layers):
        for j, blk in enumerate(layer):
            out[f"res_layers.{i}.{j}"] = blk
    return out


PHILOX_SEED = 0x5EED5EED     # = vae_gan_b200.functional.rng default seed


def philox_mask_nchw(shape, offset, p=0.5, seed=PHILOX_SEED):
    """Keep-mask a vae_gan_b200 dropout site (Philox stream `offset`) draws for an activation of
    logical shape (N,C,H,W): the kernels index elements in NHWC order."""
    n, c, h, w = shape
    m = O.philox_keep_mask(n * h * w * c, seed, offset, p).reshape(n, h, w, c)
    return torch.from_numpy(m).permute(0, 3, 1, 2).contiguous()


def philox_keep2d(n, c, offset, p=0.5, seed=PHILOX_SEED):
    sc = O.philox_keep_scale2d(n, c, seed, offset, p)
    return torch.from_numpy((sc > 0).astype(np.uint8)).reshape(n, c, 1, 1)


def rand_masks_g(spec_g, B, S, gen, first_site=0):
    """Masks for the generator's dropout sites; site ids follow forward order (0,1,2,...)."""
    masks = {}
    h = S
    for i, (pre, cin, cout, mode) in enumerate(spec_g.encoder_blocks() + spec_g.decoder_blocks()):
        masks[pre] = philox_mask_nchw((B, cin, h, h), first_site + i)
        h = h // 2 if mode == "downsample" else (h * 2 if mode == "upsample" else h)
    return masks


def rand_masks_d(spec_d, B, gen, first_site=0):
    return {pre: philox_keep2d(B, cout, first_site + i)
            for i, (pre, cin, cout, st) in enumerate(spec_d.res_blocks())}


def main():
    ns = load_reference_namespace()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(4)
    gen = torch.Generator().manual_seed(20240611)
    B, S, fs = 2, 16, 8
    spec_g = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    spec_d = O.DiscriminatorSpec(1, fs, (1, 2, 1), (1, 2, 2), (2 * fs, 4 * fs, 8 * fs), input_size=S)

    # ---------------- 1. generator fwd/bwd ----------------
    Pg = O.make_generator_params(spec_g, seed=11)
    Pd = O.make_discriminator_params(spec_d, seed=12)
    # make BN affine params non-trivial so their gradients are exercised
    for P in (Pg, Pd):
        for k in P:
            if k.endswith(("bn1.weight", "bn2.weight", "shortcut.1.weight")):
                P[k] = 1.0 + 0.2 * torch.randn(P[k].shape, generator=gen)
            elif k.endswith(("bn1.bias", "bn2.bias", "shortcut.1.bias")) or k.endswith(".bias"):
                P[k] = 0.1 * torch.randn(P[k].shape, generator=gen)
    G, D = build_ref(ns, spec_g, spec_d, Pg, Pd)
    x = torch.rand(B, 1, S, S, generator=gen)
    eps = torch.randn(B, spec_g.feature_depth, S // 4, S // 4, generator=gen)
    gm = rand_masks_g(spec_g, B, S, gen)
    for pre, blk in g_blocks(G).items():
        blk.dropout = MaskSeq([gm[pre]])
    G.train()
    with patched_randn_like([eps]):
        y, mu, lv = G(x)
    loss = 10 * (nn.L1Loss()(y, x) + nn.MSELoss()(y, x)) + \
        0.1 * (-0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp()))
    loss.backward()
    sd_after = G.state_dict()
    torch.save(dict(
        spec=dict(depth=2, length=1, feature_size=fs), seed_g=11, B=B, S=S, philox_seed=PHILOX_SEED,
        bn_perturb_seed=20240611, x=x, eps=eps, masks=gm,
        params={k: v.clone() for k, v in Pg.items() if "bn" in k or "shortcut.1" in k or k.endswith(".bias")},
        y=y.detach(), mu=mu.detach(), log_var=lv.detach(), loss=loss.detach(),
        grads={k: summarize(p.grad) for k, p in G.named_parameters()},
        buffers_after={k: v.clone() for k, v in sd_after.items() if O.is_buffer_key(k)},
    ), GOLDEN / "generator_fwd_bwd.pt")

    # eval-mode forward + decode (config 5 path)
    G.eval()
    G.set_is_training(False)
    with torch.no_grad():
        ye, mue, lve = G(x)
        dec = G.decode(eps)
    torch.save(dict(x=x, z=eps, y=ye, mu=mue, log_var=lve, decoded=dec,
                    note="uses the state (running stats) left by generator_fwd_bwd.pt"),
               GOLDEN / "generator_eval.pt")

    # ---------------- 2. discriminator fwd/bwd ----------------
    dm = rand_masks_d(spec_d, B, gen)
    for pre, blk in d_blocks(D).items():
        blk.dropout = MaskSeq([dm[pre]])
    D.train()
    xin = x.clone().requires_grad_(True)
    logits = D(xin)
    (logits * torch.tensor([[1.0], [-0.5]])).sum().backward()
    sd_after = D.state_dict()
    torch.save(dict(
        spec=dict(num_stride_conv1=1, num_features_conv1=fs, num_blocks=(1, 2, 1),
                  num_strides_res=(1, 2, 2), num_features_res=(2 * fs, 4 * fs, 8 * fs), input_size=S),
        seed_d=12, x=x, masks=dm, logit_weights=torch.tensor([[1.0], [-0.5]]),
        params={k: v.clone() for k, v in Pd.items() if "bn" in k or "shortcut.1" in k or k.endswith(".bias")},
        logits=logits.detach(), dx=xin.grad.clone(),
        grads={k: summarize(p.grad) for k, p in D.named_parameters()},
        buffers_after={k: v.clone() for k, v in sd_after.items() if O.is_buffer_key(k)},
    ), GOLDEN / "discriminator_fwd_bwd.pt")

    # ---------------- 3. single blocks, every mode ----------------
    cases = {}
    for mode in ("level", "downsample", "upsample"):
        for res_mode in ("pre-activation", "standard"):
            torch.manual_seed(5)
            blk = ns["ResBlockVAE"](6, 10, mode=mode, res_mode=res_mode)
            for m in blk.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.weight.data.uniform_(0.5, 1.5)
                    m.bias.data.uniform_(-0.3, 0.3)
            sd0 = {k: v.clone() for k, v in blk.state_dict().items()}
            xi = torch.randn(3, 6, 8, 8, generator=gen).requires_grad_(True)
            c_mask = 6 if res_mode == "pre-activation" else 10
            hm = 8 if (res_mode == "pre-activation" or mode == "level") else (4 if mode == "downsample" else 16)
            keep = philox_mask_nchw((3, c_mask, hm, hm), 0)
            blk.dropout = MaskSeq([keep])
            blk.train()
            out = blk(xi)
            gy = torch.randn(out.shape, generator=gen)
            out.backward(gy)
            cases[f"vae/{mode}/{res_mode}"] = dict(
                state=sd0, x=xi.detach().clone(), keep=keep, gy=gy, out=out.detach(), dx=xi.grad.clone(),
                grads={k: p.grad.clone() for k, p in blk.named_parameters()},
                state_after={k: v.clone() for k, v in blk.state_dict().items()})
    for stride, cin, cout in ((1, 6, 10), (2, 6, 10), (1, 10, 10)):
        for res_mode in ("pre-activation", "standard"):
            torch.manual_seed(7)
            blk = ns["ResBlockDiscriminator"](cin, cout, res_stride=stride, res_mode=res_mode)
            for m in blk.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.weight.data.uniform_(0.5, 1.5)
                    m.bias.data.uniform_(-0.3, 0.3)
            sd0 = {k: v.clone() for k, v in blk.state_dict().items()}
            xi = torch.randn(3, cin, 8, 8, generator=gen).requires_grad_(True)
            keep = philox_keep2d(3, cout, 0)
            blk.dropout = MaskSeq([keep])
            blk.train()
            out = blk(xi)
            gy = torch.randn(out.shape, generator=gen)
            out.backward(gy)
            cases[f"disc/s{stride}_{cin}_{cout}/{res_mode}"] = dict(
                state=sd0, x=xi.detach().clone(), keep=keep, gy=gy, out=out.detach(), dx=xi.grad.clone(),
                grads={k: p.grad.clone() for k, p in blk.named_parameters()},
                state_after={k: v.clone() for k, v in blk.state_dict().items()})
    torch.save(cases, GOLDEN / "blocks.pt")

    # ---------------- 4. the reference's own training loop, 2 iterations ----------------
    Pg = O.make_generator_params(spec_g, seed=21)
    Pd = O.make_discriminator_params(spec_d, seed=22)
    G, D = build_ref(ns, spec_g, spec_d, Pg, Pd)
    n_iter = 2
    xs = [torch.rand(B, 1, S, S, generator=gen, dtype=torch.float64) for _ in range(n_iter)]
    epss = [torch.randn(B, spec_g.feature_depth, S // 4, S // 4, generator=gen) for _ in range(n_iter)]
    gms = [rand_masks_g(spec_g, B, S, gen) for _ in range(n_iter)]
    # per iteration D is called 4x: real, fake, GP-interpolates, generator step
    dms = [[rand_masks_d(spec_d, B, gen) for _ in range(4)] for _ in range(n_iter)]
    for pre, blk in g_blocks(G).items():
        blk.dropout = MaskSeq([gms[i][pre] for i in range(n_iter)])
    for pre, blk in d_blocks(D).items():
        blk.dropout = MaskSeq([dms[i][c][pre] for i in range(n_iter) for c in range(4)])
    np.random.seed(99)
    alphas = [torch.tensor(np.random.random((B, 1, 1, 1)), dtype=torch.float32) for _ in range(n_iter)]
    np.random.seed(99)
    optG = torch.optim.RMSprop(G.parameters(), lr=3e-4, weight_decay=1e-5)
    optD = torch.optim.RMSprop(D.parameters(), lr=3e-4, weight_decay=1e-5)
    buf = io.StringIO()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td, patched_randn_like(epss), contextlib.redirect_stdout(buf):
        os.chdir(td)
        try:
            ns["train_network_wgan"](
                n_epochs=1, dataloader=xs, vae_generator=G, discriminator=D, optimizer_G=optG,
                optimizer_D=optD, reconstruction_loss_funs=[nn.L1Loss(), nn.MSELoss()],
                Tensor=torch.FloatTensor, sample_interval=10 ** 9,
                gan_inference_folder=Path(td) / "inf", adversarial_loss_weight=1,
                reconstruction_loss_weight=10, kl_weight=0.1, use_neptune=False, n_critics=1)
        finally:
            os.chdir(cwd)
    torch.save(dict(
        spec_g=dict(depth=2, length=1, feature_size=fs), seed_g=21, seed_d=22,
        spec_d=dict(num_stride_conv1=1, num_features_conv1=fs, num_blocks=(1, 2, 1),
                    num_strides_res=(1, 2, 2), num_features_res=(2 * fs, 4 * fs, 8 * fs), input_size=S),
        xs=xs, epss=epss, g_masks=gms, d_masks=dms, alphas=alphas, log=buf.getvalue(),
        g_after={k: summarize(v) for k, v in G.state_dict().items()},
        d_after={k: summarize(v) for k, v in D.state_dict().items()},
    ), GOLDEN / "train_wgan_gp_2iters.pt")
    print(buf.getvalue())

    # ---------------- 5. structural known-answers ----------------
    from oracle.load_reference import build_reference_models
    Gf, Df, _ = build_reference_models(ns, image_size=96)
    torch.save(dict(
        g_keys=[(k, tuple(v.shape)) for k, v in Gf.state_dict().items()],
        d_keys=[(k, tuple(v.shape)) for k, v in Df.state_dict().items()],
        g_params=sum(p.numel() for p in Gf.parameters()),
        d_params=sum(p.numel() for p in Df.parameters()),
    ), GOLDEN / "structure_96.pt")
    for f in sorted(GOLDEN.glob("*.pt")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
