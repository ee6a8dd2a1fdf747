"""Full-size checks (-m gpu): BASELINE.json's own sizes - global batch 256, 1x96x96, feature size 64 - are far beyond what
the fp64 CPU oracle can run in a test, so the kernels are checked there through size-independent properties:

* adjointness: <conv(x), dy> = <x, dgrad(dy)> = <W, wgrad(x, dy)> ties the three tensor-core GEMMs of a layer together with
  no reference at all (every indexing / tiling / padding / stride-phase bug breaks it), on the step's hottest layer shapes
  at batch 256;
* BatchNorm + LeakyReLU forward / backward on full-size tensors against plain torch fp32 math on the same GPU, elementwise
  at bf16 rounding, with the parameter gradients and torch's running-statistics rule;
* batch-mate independence: in eval mode a sample's reconstruction / logit does not depend on the other 255 samples
  (batch 256 vs slices of 64) - the pixel tiles of the convolutions span images (TN > 1 boxes), so this is a real check;
* one full-size training iteration: finite losses, the BCE discriminator loss of an untrained network of order 2 ln 2, every
  parameter of both networks moved by at most lr (first Adam step), BatchNorm running statistics updated.
"""
import math

import pytest
import torch

from tests.gpu_util import dev

pytestmark = pytest.mark.gpu
B, S, FS = 256, 96, 64


def V():
    import vae_gan_b200 as v
    return v


def VF():
    import vae_gan_b200.functional as vf
    return vf


def _dot(a, b):
    return float(torch.dot(a.reshape(-1).double(), b.reshape(-1).double())) if a.numel() < (1 << 26) else \
        sum(float(torch.dot(x.reshape(-1).double(), y.reshape(-1).double())) for x, y in zip(a.chunk(16), b.chunk(16)))


# (c_in, c_out, h, k, stride, pad, transposed): D res conv2 @96 (the roofline layer), D downsample conv1, the 1x1 stride-2 shortcut,
# G upsample ConvTranspose2d 4x4, D 512 -> 512 @24, G 64 -> 64 @96
LAYERS = [(128, 128, 96, 3, 1, 1, False), (128, 256, 96, 3, 2, 1, False), (128, 256, 96, 1, 2, 0, False),
          (256, 128, 24, 4, 2, 1, True), (512, 512, 24, 3, 1, 1, False), (64, 64, 96, 3, 1, 1, False)]


@pytest.mark.parametrize("cin,cout,h,k,stride,pad,tr", LAYERS)
def test_conv_adjointness_at_batch_256(cin, cout, h, k, stride, pad, tr):
    vf = VF()
    g = torch.Generator(device=dev()).manual_seed(cin * 7 + cout + h)
    x = vf.as_act(torch.randn(B, cin, h, h, generator=g, device=dev()), torch.bfloat16).requires_grad_(True)
    wshape = (cin, cout, k, k) if tr else (cout, cin, k, k)
    w = (torch.randn(wshape, generator=g, device=dev()) / math.sqrt(cin * k * k)).to(torch.bfloat16).float().requires_grad_(True)
    y = vf.conv(x, w, None, geom=vf.ConvGeom(k, stride, pad, tr))
    dy = vf.as_act(torch.randn(y.shape, generator=g, device=dev()), torch.bfloat16)
    y.backward(dy)
    torch.cuda.synchronize()
    s_fwd, s_dgrad, s_wgrad = _dot(y.detach(), dy), _dot(x.detach(), x.grad), _dot(w.detach(), w.grad)
    scale = math.sqrt(float(y.numel()))          # the dot products are sums of y.numel() unit-variance terms
    # y and dx are rounded to bf16 (uniform error, std 0.58 * 2^-9 relative per element, random sign): the differences
    # have a standard deviation of ~1.5e-3 * scale; 1e-2 is > 6 sigma, while e.g. 1 % of the outputs missing moves a sum by
    # 0.1 * scale
    assert abs(s_fwd - s_dgrad) <= 1e-2 * scale, (s_fwd, s_dgrad, scale)
    assert abs(s_fwd - s_wgrad) <= 1e-2 * scale, (s_fwd, s_wgrad, scale)
    assert abs(s_fwd) < 8 * scale and math.isfinite(s_fwd)


@pytest.mark.parametrize("c,h", [(128, 96), (512, 24), (1, 96)])
def test_batchnorm_against_torch_fp32_at_batch_256(c, h):
    """Training-mode BatchNorm + LeakyReLU forward / backward on a full-size tensor (2.4 M values per channel) against plain
    torch fp32 math on the same GPU, elementwise at bf16 rounding; statistics identities of the output; running statistics.
    (The textbook identities sum(dx) = 0 and sum(dx * x_hat) = 0 cannot be tested on the bf16 result itself: dx = a * g + c0 +
    c1 * (x - mean) with |c0|, |c1 (x - mean)| ~ 1e-3 far below the bf16 resolution of a * g, and a * g takes few distinct
    mantissas for a bf16 g, so the rounding of dx is not unbiased with respect to those terms - measured sum(dx) up to 10 % of
    a * sum(g); the fp32 reference below has the same property once rounded.)"""
    import torch.nn.functional as F
    vf = VF()
    g = torch.Generator(device=dev()).manual_seed(c + h)
    x = vf.as_act(torch.randn(B, c, h, h, generator=g, device=dev()) * 1.7 + 0.3, torch.bfloat16).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(dev()).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * torch.randn(c, generator=g, device=dev()))
        bn.bias.copy_(0.2 * torch.randn(c, generator=g, device=dev()))
    y = vf.bn_act(x, bn, slope=0.2, training=True)
    dy = vf.as_act(torch.randn(y.shape, generator=g, device=dev()), torch.bfloat16)
    y.backward(dy)
    # reference: torch fp32 on the same (bf16-rounded) inputs
    xr = x.detach().float().contiguous().requires_grad_(True)
    gam, bet = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    yr = F.leaky_relu(F.batch_norm(xr, None, None, gam, bet, True, 0.1, 1e-5), 0.2)
    yr.backward(dy.float())
    ulp = 2.0 ** -8                                          # bf16 round-to-nearest: relative error <= 2^-9; 2x margin
    yerr = (y.detach().float() - yr.detach()).abs() - ulp * yr.detach().abs()
    assert float(yerr.max()) <= 1e-5, float(yerr.max())
    # dx: elements whose pre-activation is within fp32 rounding of the LeakyReLU kink (|gamma x_hat + beta| < ~1e-7; a few
    # tens of 3e8) may take the other slope in one of the two implementations
    dxerr = (x.grad.float() - xr.grad).abs() - ulp * xr.grad.abs()
    n_bad = int((dxerr > 1e-5).sum())
    assert n_bad <= max(2, int(1e-6 * dxerr.numel())), (n_bad, float(dxerr.max()))
    # parameter gradients: sums of 1.5e5 .. 2.4e6 terms; a kink flip (above) moves one by up to 0.8 * |dy * x_hat| ~ 1
    assert float((bn.weight.grad - gam.grad).abs().max()) <= 2e-4 * float(gam.grad.abs().max()) + 2.0
    assert float((bn.bias.grad - bet.grad).abs().max()) <= 2e-4 * float(bet.grad.abs().max()) + 2.0
    xf = x.detach().float()
    assert float((bn.running_mean - 0.1 * xf.mean((0, 2, 3))).abs().max()) <= 1e-4
    assert float((bn.running_var - (0.9 + 0.1 * xf.var((0, 2, 3), unbiased=True))).abs().max()) <= 1e-3


def test_eval_outputs_do_not_depend_on_batch_mates_at_batch_256():
    v = V()
    with v.compute_dtype(torch.bfloat16), torch.no_grad():
        torch.manual_seed(0)
        G, D = v.build_vae_gan(feature_size=FS, image_size=S)
        G, D = G.to(dev()).eval(), D.to(dev()).eval()
        G.set_is_training(False)
        x = torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(1)).to(dev())
        full_gen = G(x)[0]
        full_logit = D(x)
        for lo in (0, 64, 192):
            part_gen = G(x[lo:lo + 64])[0]
            part_logit = D(x[lo:lo + 64])
            # convolutions, eval-mode BatchNorm and the activations are per-sample, and every kernel form reduces an output in
            # the same order: identical bits whatever the batch (bound: bf16 noise, should a tile choice ever change that)
            gerr = float((part_gen - full_gen[lo:lo + 64]).abs().max())
            assert gerr <= 2e-2 * max(1.0, float(full_gen.abs().max())), f"reconstruction of samples {lo}.. depends on the batch: {gerr}"
            if gerr != 0.0:
                print(f"note: batch 64 vs 256 reconstructions differ by {gerr} (not bit-identical)")
            # the Linear head runs split-K with a batch-dependent split count: its fp32 sums differ in the last bits, and
            # re-rounding them to bf16 for the next Linear layer flips single ulps (2^-8) - measured 4.5e-3 of the largest logit
            err = float((part_logit - full_logit[lo:lo + 64]).abs().max())
            assert err <= 2e-2 * max(1.0, float(full_logit.abs().max())), err


def test_full_size_training_iteration_sanity():
    v = V()
    lr = 3e-4
    with v.compute_dtype(torch.bfloat16):
        torch.manual_seed(0)
        G, D = v.build_vae_gan(feature_size=FS, image_size=S)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        tr = v.VaeGanTrainer(G, D, lr=lr)
        p0 = [tr.fg.p.clone(), tr.fd.p.clone()]
        rm0 = D.bn1.running_mean.clone()
        x = torch.rand(B, 1, S, S, generator=torch.Generator().manual_seed(2)).to(dev())
        tr.step(x)
        l = tr.read_losses()
    assert all(math.isfinite(val) for val in l.values()), l
    # untrained discriminator: BCE(real, 1) + BCE(fake, 0) of the order of 2 ln 2
    assert 0.3 < l["d_loss"] < 5.0, l
    assert 0.0 < l["recon"] < 50.0 and l["kl"] > 0.0
    for before, flat in zip(p0, (tr.fg, tr.fd)):
        step = (flat.p - before).abs()
        assert float(step.max()) <= lr * 1.001            # the first Adam step is lr * g / (|g| + eps)
        assert float((step > 0.5 * lr).float().mean()) > 0.5, "most parameters receive a gradient"
    assert not torch.equal(D.bn1.running_mean, rm0) and int(D.bn1.num_batches_tracked) == 3     # D runs three forwards per iteration
