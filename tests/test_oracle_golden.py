"""Pin oracle/vaegan_oracle.py against golden vectors produced by the REAL reference
(oracle/make_golden.py) -- CPU only."""
import re

import numpy as np
import pytest
import torch

from oracle import vaegan_oracle as O

TOL = dict(rtol=2e-5, atol=2e-6)


def check(got, want, name="", rtol=2e-5, atol=2e-6):
    if isinstance(want, dict) and want.get("summary"):
        flat = got.detach().float().flatten()
        smp = flat[:: want["stride"]][:4096]
        torch.testing.assert_close(smp, want["sample"], rtol=rtol, atol=atol, msg=lambda m: f"{name}: {m}")
        assert abs(float(flat.double().abs().sum()) - want["asum"]) <= 1e-4 * max(1.0, want["asum"]), name
        return
    torch.testing.assert_close(got.detach().float(), want.float(), rtol=rtol, atol=atol,
                               msg=lambda m: f"{name}: {m}")


def _perturbed(P, stored):
    P = dict(P)
    for k, v in stored.items():
        P[k] = v.clone()
    return P


def test_generator_fwd_bwd(golden_dir):
    g = torch.load(golden_dir / "generator_fwd_bwd.pt")
    spec = O.GeneratorSpec(**g["spec"])
    P = _perturbed(O.make_generator_params(spec, seed=g["seed_g"]), g["params"])
    P = O.clone_params(P, requires_grad=True)
    y, mu, lv = O.generator_forward(g["x"], P, spec, True, True, g["eps"], g["masks"])
    check(y, g["y"], "y")
    check(mu, g["mu"], "mu")
    check(lv, g["log_var"], "log_var")
    loss = 10 * O.reconstruction_loss(y, g["x"]) + 0.1 * O.kl_divergence(mu, lv)
    check(loss, g["loss"], "loss", rtol=1e-5)
    keys = O.trainable_keys(P)
    grads = torch.autograd.grad(loss, [P[k] for k in keys])
    for k, gr in zip(keys, grads):
        check(gr, g["grads"][k], k, rtol=1e-4, atol=1e-4 * float(gr.abs().max()) + 1e-7)
    for k, v in g["buffers_after"].items():
        check(P[k].float(), v.float(), k)

    # eval-mode forward and decode() continue from that state
    e = torch.load(golden_dir / "generator_eval.pt")
    Pn = O.clone_params(P)
    with torch.no_grad():
        ye, mue, lve = O.generator_forward(e["x"], Pn, spec, False, False)
        dec = O.decoder_forward(e["z"], Pn, spec, False)
    check(ye, e["y"], "eval y")
    check(mue, e["mu"], "eval mu")
    check(dec, e["decoded"], "decode")


def test_discriminator_fwd_bwd(golden_dir):
    g = torch.load(golden_dir / "discriminator_fwd_bwd.pt")
    spec = O.DiscriminatorSpec(**g["spec"])
    P = _perturbed(O.make_discriminator_params(spec, seed=g["seed_d"]), g["params"])
    P = O.clone_params(P, requires_grad=True)
    x = g["x"].clone().requires_grad_(True)
    logits = O.discriminator_forward(x, P, spec, True, g["masks"])
    check(logits, g["logits"], "logits")
    keys = O.trainable_keys(P)
    grads = torch.autograd.grad((logits * g["logit_weights"]).sum(), [x] + [P[k] for k in keys])
    check(grads[0], g["dx"], "dx", rtol=1e-4, atol=1e-6)
    for k, gr in zip(keys, grads[1:]):
        check(gr, g["grads"][k], k, rtol=1e-4, atol=1e-4 * float(gr.abs().max()) + 1e-7)
    for k, v in g["buffers_after"].items():
        check(P[k].float(), v.float(), k)


def test_blocks_every_mode(golden_dir):
    cases = torch.load(golden_dir / "blocks.pt")
    assert len(cases) == 12
    for name, c in cases.items():
        kind, cfg, res_mode = name.split("/")
        P = {"b." + k: v.clone() for k, v in c["state"].items()}
        P = O.clone_params(P, requires_grad=True)
        x = c["x"].clone().requires_grad_(True)
        if kind == "vae":
            out = O.resblock_vae(x, P, "b", cfg, res_mode, True, c["keep"])
        else:
            m = re.match(r"s(\d)_(\d+)_(\d+)", cfg)
            st, cin, cout = int(m.group(1)), int(m.group(2)), int(m.group(3))
            out = O.resblock_discriminator(x, P, "b", cin, cout, st, res_mode, True, c["keep"])
        check(out, c["out"], name + " out", rtol=1e-4, atol=1e-5)
        keys = O.trainable_keys(P)
        grads = torch.autograd.grad(out, [x] + [P[k] for k in keys], c["gy"])
        check(grads[0], c["dx"], name + " dx", rtol=1e-4, atol=1e-5)
        for k, gr in zip(keys, grads[1:]):
            check(gr, c["grads"][k[2:]], name + " " + k, rtol=1e-4, atol=1e-4 * float(gr.abs().max()) + 1e-6)
        for k, v in c["state_after"].items():
            if O.is_buffer_key(k):
                check(P["b." + k].float(), v.float(), name + " " + k, rtol=1e-4, atol=1e-6)


def test_reference_training_loop_two_iterations(golden_dir):
    """The reference's own train_network_wgan (WGAN-GP + clamp + RMSprop) vs oracle.train_step."""
    g = torch.load(golden_dir / "train_wgan_gp_2iters.pt")
    spec_g = O.GeneratorSpec(**g["spec_g"])
    spec_d = O.DiscriminatorSpec(**g["spec_d"])
    Pg = O.make_generator_params(spec_g, seed=g["seed_g"])
    Pd = O.make_discriminator_params(spec_d, seed=g["seed_d"])
    og = O.OptState(kind="rmsprop", lr=3e-4, weight_decay=1e-5)
    od = O.OptState(kind="rmsprop", lr=3e-4, weight_decay=1e-5)
    logged = re.findall(r"\[D loss: ([-\d.e+]+)\] \[G loss: ([-\d.e+]+)\] \[Recon loss: ([-\d.e+]+)\] \[KL: ([-\d.e+]+)\]",
                        g["log"])
    assert len(logged) == 2
    for i in range(2):
        dm = g["d_masks"][i]
        out = O.train_step(Pg, Pd, og, od, g["xs"][i].float(), spec_g, spec_d,
                           eps_noise=g["epss"][i], g_masks=g["g_masks"][i], d_masks_real=dm[0],
                           d_masks_fake=dm[1], d_masks_gp=dm[2], d_masks_gen=dm[3],
                           loss_mode="wgan_gp", gp_alpha=g["alphas"][i])
        d_l, g_l, r_l, k_l = (float(v) for v in logged[i])
        assert abs(float(out["d_loss"]) - d_l) <= 2e-3 + 1e-3 * abs(d_l)
        assert abs(float(out["g_loss"]) - g_l) <= 2e-3 + 1e-4 * abs(g_l)
        assert abs(float(out["recon"]) - r_l) <= 2e-3
        assert abs(float(out["kl"]) - k_l) <= 2e-3 + 1e-4 * abs(k_l)
    # known-answer from the notebook's own log (README.md:971-972): after the first clamp
    # the critic loss sits at ~lambda_gp * 1
    assert abs(float(out["d_loss"]) - 10.0) < 0.05
    for k, v in g["g_after"].items():
        check(Pg[k].float(), v if not isinstance(v, torch.Tensor) else v.float(), "G." + k, rtol=2e-4, atol=1e-5)
    for k, v in g["d_after"].items():
        check(Pd[k].float(), v if not isinstance(v, torch.Tensor) else v.float(), "D." + k, rtol=2e-4, atol=1e-5)


def test_structure_matches_reference_state_dict(golden_dir):
    s = torch.load(golden_dir / "structure_96.pt")
    Pg = O.make_generator_params(O.GeneratorSpec())
    Pd = O.make_discriminator_params(O.DiscriminatorSpec(input_size=96))
    assert [(k, tuple(v.shape)) for k, v in Pg.items()] and \
        sorted((k, tuple(v.shape)) for k, v in Pg.items()) == sorted(s["g_keys"])
    assert sorted((k, tuple(v.shape)) for k, v in Pd.items()) == sorted(s["d_keys"])
    assert sum(Pg[k].numel() for k in O.trainable_keys(Pg)) == s["g_params"] == 4192783
    assert sum(Pd[k].numel() for k in O.trainable_keys(Pd)) == s["d_params"] == 24353857


def test_optimizers_match_torch_optim():
    torch.manual_seed(0)
    for kind in ("adam", "rmsprop"):
        p0 = torch.randn(257)
        P = {"w": p0.clone()}
        ref = p0.clone().requires_grad_(True)
        opt = (torch.optim.Adam([ref], lr=3e-4) if kind == "adam"
               else torch.optim.RMSprop([ref], lr=3e-4, weight_decay=1e-5))
        st = O.OptState(kind=kind, lr=3e-4, weight_decay=0.0 if kind == "adam" else 1e-5)
        for _ in range(5):
            g = torch.randn(257)
            ref.grad = g.clone()
            opt.step()
            O.optimizer_step(P, {"w": g}, st)
        torch.testing.assert_close(P["w"], ref.detach(), rtol=1e-6, atol=1e-7)


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    out = O.philox4x32_10([np.uint32(0)] * 4, (np.uint32(0), np.uint32(0)))
    assert [int(v) for v in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    ff = np.uint32(0xFFFFFFFF)
    out = O.philox4x32_10([ff] * 4, (ff, ff))
    assert [int(v) for v in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    out = O.philox4x32_10([np.uint32(0x243F6A88), np.uint32(0x85A308D3), np.uint32(0x13198A2E), np.uint32(0x03707344)],
                          (np.uint32(0xA4093822), np.uint32(0x299F31D0)))
    assert [int(v) for v in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    m = O.philox_keep_mask(100000, seed=2024, offset=3, p=0.5)
    assert 0.49 < m.mean() < 0.51
