"""Module- and step-level parity (-m gpu), through the nn.Module surface -> autograd.Functions ->
C ABI.  Checked against (a) golden vectors produced by the REAL reference notebook classes
(tests/golden, fp32 path, tolerance 1e-5 on activations) and (b) the oracle on seeded inputs at the
real sizes (bf16 tensor-core path, tolerance 2e-2)."""
import re

import pytest
import torch

from oracle import vaegan_oracle as O
from tests.gpu_util import (assert_close, dev, discriminator_masks, generator_masks, load_params_into, nchw,
                            philox_keep2d, philox_mask_nchw, relmax)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    import vae_gan_b200  # noqa: F401
    yield


def V():
    import vae_gan_b200 as v
    return v


def _check_summary(got, want, tol, name):
    if isinstance(want, dict) and want.get("summary"):
        flat = got.detach().float().flatten().cpu()
        return assert_close(flat[:: want["stride"]][:4096], want["sample"], tol, name)
    return assert_close(got, want, tol, name)


# ------------------------------------------------------------------------------------------------
# (a) golden vectors from the executed reference, fp32 path
# ------------------------------------------------------------------------------------------------
def test_blocks_against_reference_golden(golden_dir):
    v = V()
    cases = torch.load(golden_dir / "blocks.pt")
    worst = {}
    with v.compute_dtype(torch.float32):
        for name, c in cases.items():
            kind, cfg, res_mode = name.split("/")
            if kind == "vae":
                blk = v.ResBlockVAE(6, 10, mode=cfg, res_mode=res_mode)
            else:
                m = re.match(r"s(\d)_(\d+)_(\d+)", cfg)
                st, cin, cout = int(m.group(1)), int(m.group(2)), int(m.group(3))
                blk = v.ResBlockDiscriminator(cin, cout, res_stride=st, res_mode=res_mode)
            blk.load_state_dict(c["state"], strict=True)
            blk = blk.to(dev()).train()
            v.rng.seed = 0x5EED5EED
            v.rng.reset_sites()
            x = c["x"].to(dev()).requires_grad_(True)
            out = blk(x)
            assert out.dtype == torch.float32 and out.shape == c["out"].shape
            out.backward(c["gy"].to(dev()))
            e = [assert_close(out, c["out"], 2e-5, name + " out"),
                 assert_close(x.grad, c["dx"], 1e-4, name + " dx")]
            for k, p in blk.named_parameters():
                e.append(assert_close(p.grad, c["grads"][k], 2e-4, name + " grad " + k))
            sd = blk.state_dict()
            for k, val in c["state_after"].items():
                if O.is_buffer_key(k):
                    assert_close(sd[k].float(), val.float(), 5e-5, name + " buffer " + k)
            worst[name] = max(e)
    print("worst max-normalised errors:", {k: f"{e:.1e}" for k, e in worst.items()})


def test_generator_against_reference_golden(golden_dir):
    v = V()
    g = torch.load(golden_dir / "generator_fwd_bwd.pt")
    spec = O.GeneratorSpec(**g["spec"])
    P = O.make_generator_params(spec, seed=g["seed_g"])
    P.update({k: t.clone() for k, t in g["params"].items()})
    with v.compute_dtype(torch.float32):
        G = v.UnsupervisedGeneratorNetwork(
            encoder=v.Encoder(1, spec.depth, spec.length, spec.feature_size),
            decoder=v.Decoder(spec.feature_depth, spec.depth, spec.length, 1),
            code_processor=v.SpatialVAECodeProcessor(spec.feature_depth, True), is_vae=True)
        load_params_into(G, P)
        G = G.to(dev()).train()
        G.code_processor.eps_override = g["eps"]
        v.rng.seed = g["philox_seed"]
        v.rng.reset_sites()
        x = g["x"].to(dev())
        y, mu, lv = G(x)
        assert_close(y, g["y"], 2e-5, "y")
        assert_close(mu, g["mu"], 2e-5, "mu")
        assert_close(lv, g["log_var"], 2e-5, "log_var")
        loss = 10 * O.reconstruction_loss(y, x) + 0.1 * O.kl_divergence(mu, lv)     # user-side torch ops: drop-in use
        assert abs(float(loss) - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
        loss.backward()
        for k, p in G.named_parameters():
            _check_summary(p.grad, g["grads"][k], 3e-4, "grad " + k)
        sd = G.state_dict()
        for k, val in g["buffers_after"].items():
            assert_close(sd[k].float(), val.float(), 5e-5, "buffer " + k)
        # eval forward and decode (BASELINE config 5 path) continue from that state
        e = torch.load(golden_dir / "generator_eval.pt")
        G.eval()
        G.set_is_training(False)
        with torch.no_grad():
            ye, mue, lve = G(e["x"].to(dev()))
            dec = G.decode(e["z"].to(dev()))
        assert_close(ye, e["y"], 2e-5, "eval y")
        assert_close(mue, e["mu"], 2e-5, "eval mu")
        assert_close(dec, e["decoded"], 2e-5, "decode")


def test_discriminator_against_reference_golden(golden_dir):
    v = V()
    g = torch.load(golden_dir / "discriminator_fwd_bwd.pt")
    sp = dict(g["spec"])
    spec = O.DiscriminatorSpec(**sp)
    P = O.make_discriminator_params(spec, seed=g["seed_d"])
    P.update({k: t.clone() for k, t in g["params"].items()})
    with v.compute_dtype(torch.float32):
        D = v.Discriminator(v.ResBlockDiscriminator, sp["num_stride_conv1"], sp["num_features_conv1"],
                            list(sp["num_blocks"]), list(sp["num_strides_res"]), list(sp["num_features_res"]),
                            input_size=sp["input_size"])
        load_params_into(D, P)
        D = D.to(dev()).train()
        v.rng.seed = 0x5EED5EED
        v.rng.reset_sites()
        x = g["x"].to(dev()).requires_grad_(True)
        logits = D(x)
        assert_close(logits, g["logits"], 3e-5, "logits")
        (logits * g["logit_weights"].to(dev())).sum().backward()
        assert_close(x.grad, g["dx"], 3e-4, "dx")
        for k, p in D.named_parameters():
            _check_summary(p.grad, g["grads"][k], 5e-4, "grad " + k)
        sd = D.state_dict()
        for k, val in g["buffers_after"].items():
            assert_close(sd[k].float(), val.float(), 5e-5, "buffer " + k)


# ------------------------------------------------------------------------------------------------
# (b) the real sizes on the bf16 tensor-core path vs the oracle
# ------------------------------------------------------------------------------------------------
def _oracle_generator(P, spec, x, eps, masks):
    Pr = O.clone_params(P, requires_grad=True)
    y, mu, lv = O.generator_forward(x, Pr, spec, True, True, eps, masks)
    return Pr, y, mu, lv


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float32, 5e-5)])
def test_generator_full_size_vs_oracle(dtype, tol):
    v = V()
    B, S = 4, 96
    spec = O.GeneratorSpec(depth=2, length=1, feature_size=64)
    P = O.make_generator_params(spec, seed=3)
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(B, 1, S, S, generator=gen)
    eps = torch.randn(B, 256, S // 4, S // 4, generator=gen)
    with v.compute_dtype(dtype):
        G, _ = v.build_vae_gan(image_size=S)
        load_params_into(G, P)
        G = G.to(dev()).train()
        G.code_processor.eps_override = eps
        v.rng.seed = 2024
        v.rng.reset_sites()
        y, mu, lv = G(x.to(dev()))
        loss = 10 * O.reconstruction_loss(y, x.to(dev())) + 0.1 * O.kl_divergence(mu, lv)
        loss.backward()
    masks, _ = generator_masks(spec, B, S, 2024)
    Pr, yr, mur, lvr = _oracle_generator(P, spec, x, eps, masks)
    lossr = 10 * O.reconstruction_loss(yr, x) + 0.1 * O.kl_divergence(mur, lvr)
    keys = O.trainable_keys(Pr)
    grads = dict(zip(keys, torch.autograd.grad(lossr, [Pr[k] for k in keys])))
    errs = {"y": assert_close(y, yr, tol, "y"), "mu": assert_close(mu, mur, tol, "mu"),
            "lv": assert_close(lv, lvr, tol, "log_var")}
    assert abs(float(loss) - float(lossr)) <= tol * abs(float(lossr))
    worst = 0.0
    for k, p in G.named_parameters():
        worst = max(worst, assert_close(p.grad, grads[k], tol * 2.5, "grad " + k))
    print(f"[{dtype}] activations {errs}, worst param-grad error {worst:.2e}, loss {float(loss):.4f} vs {float(lossr):.4f}")


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float32, 5e-5)])
def test_discriminator_full_size_vs_oracle(dtype, tol):
    v = V()
    B, S = 4, 96
    spec = O.DiscriminatorSpec(input_size=S)
    P = O.make_discriminator_params(spec, seed=4)
    gen = torch.Generator().manual_seed(99)
    x = torch.rand(B, 1, S, S, generator=gen)
    with v.compute_dtype(dtype):
        _, D = v.build_vae_gan(image_size=S)
        load_params_into(D, P)
        D = D.to(dev()).train()
        v.rng.seed = 7
        v.rng.reset_sites()
        xi = x.to(dev()).requires_grad_(True)
        logits = D(xi)
        wts = torch.tensor([[1.0], [-0.5], [0.25], [2.0]], device=dev())
        (logits * wts).sum().backward()
    masks, _ = discriminator_masks(spec, B, 7, 0)
    Pr = O.clone_params(P, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    lr = O.discriminator_forward(xr, Pr, spec, True, masks)
    keys = O.trainable_keys(Pr)
    grads = torch.autograd.grad((lr * wts.cpu()).sum(), [xr] + [Pr[k] for k in keys])
    e_log = assert_close(logits, lr, tol, "logits")
    e_dx = assert_close(xi.grad, grads[0], tol * 2.5, "dx")
    gd = dict(zip(keys, grads[1:]))
    worst = 0.0
    for k, p in D.named_parameters():
        worst = max(worst, assert_close(p.grad, gd[k], tol * 2.5, "grad " + k))
    sd = D.state_dict()
    for k in Pr:
        if O.is_buffer_key(k) and not k.endswith("num_batches_tracked"):
            assert_close(sd[k].float(), Pr[k].float(), max(tol, 1e-4), "buffer " + k)
    print(f"[{dtype}] logits {e_log:.2e} dx {e_dx:.2e} worst param-grad {worst:.2e}")


# ------------------------------------------------------------------------------------------------
# the training iteration
# ------------------------------------------------------------------------------------------------
def _run_trainer_vs_oracle(dtype, loss_mode, opt, B, S, fs, steps, tol_loss, tol_param, use_graph=False):
    v = V()
    spec_g = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    spec_d = O.DiscriminatorSpec(1, fs, (1, 1, 1), (1, 2, 2), (2 * fs, 4 * fs, 8 * fs), input_size=S)
    Pg = O.make_generator_params(spec_g, seed=5)
    Pd = O.make_discriminator_params(spec_d, seed=6)
    gen = torch.Generator().manual_seed(31)
    xs = [torch.rand(B, 1, S, S, generator=gen) for _ in range(steps)]
    epss = [torch.randn(B, spec_g.feature_depth, S // 4, S // 4, generator=gen) for _ in range(steps)]
    seed = 4242
    with v.compute_dtype(dtype):
        G, D = v.build_vae_gan(feature_size=fs, image_size=S)
        load_params_into(G, Pg)
        load_params_into(D, Pd)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        v.rng.seed = seed
        v.rng.step_tensor(dev()).zero_()
        tr = v.VaeGanTrainer(G, D, loss_mode=loss_mode, optimizer=opt, lr=3e-4)
        og = O.OptState(kind=opt, lr=3e-4, weight_decay=1e-5 if opt == "rmsprop" else 0.0)
        od = O.OptState(kind=opt, lr=3e-4, weight_decay=1e-5 if opt == "rmsprop" else 0.0)
        Pg_r, Pd_r = O.clone_params(Pg), O.clone_params(Pd)
        report = []
        for i in range(steps):
            G.code_processor.eps_override = epss[i]
            losses = tr.step(xs[i].to(dev()))
            got = tr.read_losses()
            step = i + 1
            gm, site = generator_masks(spec_g, B, S, seed, 0, step)
            dm_real, site = discriminator_masks(spec_d, B, seed, site, step)
            dm_fake, site = discriminator_masks(spec_d, B, seed, site, step)
            dm_gen, site = discriminator_masks(spec_d, B, seed, site, step)
            want = O.train_step(Pg_r, Pd_r, og, od, xs[i], spec_g, spec_d, eps_noise=epss[i], g_masks=gm,
                                d_masks_real=dm_real, d_masks_fake=dm_fake, d_masks_gen=dm_gen, loss_mode=loss_mode)
            for k in ("d_loss", "g_loss", "recon", "kl", "adv"):
                w = float(want[k])
                assert abs(got[k] - w) <= tol_loss * max(abs(w), 1e-3), (i, k, got[k], w)
            report.append({k: (round(got[k], 5), round(float(want[k]), 5)) for k in ("d_loss", "g_loss", "kl")})
            assert_close(tr.last["gen"], want["gen"], tol_loss * 2, f"step {i} gen")
        print(f"[{dtype} {loss_mode}/{opt}] (ours, oracle):", report)
        for k, p in G.named_parameters():
            assert_close(p.data, Pg_r[k], tol_param, "G param " + k)
        for k, p in D.named_parameters():
            assert_close(p.data, Pd_r[k], tol_param, "D param " + k)
    return tr


def test_train_step_fp32_bce_adam_vs_oracle():
    _run_trainer_vs_oracle(torch.float32, "bce", "adam", B=2, S=32, fs=8, steps=2, tol_loss=2e-4, tol_param=2e-3)


def test_train_step_fp32_wgan_rmsprop_vs_oracle():
    """The reference's own critic loss + clamp + RMSprop (without the gradient penalty)."""
    _run_trainer_vs_oracle(torch.float32, "wgan", "rmsprop", B=2, S=32, fs=8, steps=2, tol_loss=2e-4, tol_param=5e-3)


def test_train_step_bf16_full_size_vs_oracle():
    _run_trainer_vs_oracle(torch.bfloat16, "bce", "adam", B=4, S=96, fs=64, steps=1, tol_loss=2e-2, tol_param=5e-2)


def test_cuda_graph_replay_matches_eager():
    """Whole-iteration CUDA graph: replay i must equal eager step i (same Philox step counter)."""
    v = V()
    B, S, fs = 2, 32, 64
    gen = torch.Generator().manual_seed(8)
    xs = [torch.rand(B, 1, S, S, generator=gen).to(dev()) for _ in range(6)]

    def make():
        torch.manual_seed(0)
        G, D = v.build_vae_gan(feature_size=fs, image_size=S)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        v.rng.seed = 11
        v.rng.step_tensor(dev()).zero_()
        return v.VaeGanTrainer(G, D)

    with v.compute_dtype(torch.bfloat16):
        eager_tr = make()
        for _ in range(3):
            eager_tr.step(xs[0])
        graph_tr = make()
        graph_tr.capture(xs[0], warmup=3)          # 3 eager warm-up steps on xs[0], then capture
        for x in xs[1:4]:
            eager_tr.step(x)
            ref = eager_tr.read_losses()
            graph_tr.step(x)
            got = graph_tr.read_losses()
            for k in ref:
                assert abs(got[k] - ref[k]) <= 2e-3 * max(1.0, abs(ref[k])), (k, got[k], ref[k])
        assert all(abs(val) < float("inf") for val in got.values())
        # replay draws fresh dropout masks: the device step counter advanced once per step
        assert int(v.rng.step_tensor(dev())) == 6


def test_known_answer_clamped_critic():
    """Behavioural known-answer from the notebook's log (README.md:971-972): in WGAN mode every D
    parameter is clamped to +-0.01 after the first step (README.md:805-806)."""
    v = V()
    with v.compute_dtype(torch.float32):
        torch.manual_seed(0)
        G, D = v.build_vae_gan(feature_size=8, image_size=32)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        tr = v.VaeGanTrainer(G, D, loss_mode="wgan", optimizer="rmsprop")
        x = torch.rand(2, 1, 32, 32, generator=torch.Generator().manual_seed(0)).to(dev())
        tr.step(x)
        assert max(float(p.abs().max()) for p in D.parameters()) <= 0.01 + 1e-8
        tr.step(x)
        l = tr.read_losses()
        assert abs(l["real_loss"]) < 0.05 and abs(l["fake_loss"]) < 0.05      # tiny logits after the clamp
