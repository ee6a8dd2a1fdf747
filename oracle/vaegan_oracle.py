"""CPU oracle for the VAE-GAN training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict driven) restatement of the reference notebook's hot path in plain
PyTorch fp32/fp64 ops.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it; the product package
(`vae_gan_b200/`) never does.

Where the arithmetic lives: the reference delegates every numeric op to PyTorch
(third-party, version unpinned by the reference; pinned here to the container's torch
2.11.0).  This file therefore restates the reference's *composition* of those ops and is
PINNED against the executed reference itself: `oracle/make_golden.py` runs the real
notebook classes (via `oracle/load_reference.py`) and stores inputs/weights/masks/outputs
in `tests/golden/`; `tests/test_oracle_golden.py` checks this file against those vectors,
and (inside the build container) directly against the live reference modules.

Citations are to /root/reference/README.md (byte-identical to gan.ipynb's code cells).
All functions take a flat dict `P` whose keys are exactly the reference modules'
state_dict keys; buffers in `P` are updated in place the way the
reference's modules update theirs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

BN_EPS = 1e-5          # nn.BatchNorm2d default, README.md:143
BN_MOMENTUM = 0.1
SLOPE_G = 0.01         # nn.LeakyReLU() default, README.md:172
SLOPE_D = 0.2          # README.md:394, 437
SN_EPS = 1e-12         # nn.utils.spectral_norm default, README.md:378


# ----------------------------------------------------------------------------------------
# architecture specs (names follow README.md:230,239,244,278,282,289)
# ----------------------------------------------------------------------------------------
@dataclass
class GeneratorSpec:
    depth: int = 2
    length: int = 1
    feature_size: int = 64
    in_channels: int = 1
    reconstruction_channels: int = 1

    @property
    def feature_depth(self) -> int:           # README.md:882
        return self.feature_size * (2 ** self.depth)

    def encoder_blocks(self) -> List[Tuple[str, int, int, str]]:
        """(state-dict prefix, cin, cout, mode) in forward order -- README.md:225-246."""
        out = []
        cin, fs = self.in_channels, self.feature_size
        for i in range(self.length):
            out.append((f"encoder.encoder.encoder-depth_0-level_{i}", cin, fs, "level"))
            cin = fs
        for d in range(1, self.depth + 1):
            cin = fs
            fs *= 2
            out.append((f"encoder.encoder.encoder-depth_{d}-downsample", cin, fs, "downsample"))
            for item in range(self.length - 1):
                out.append((f"encoder.encoder.encoder-depth_{d}-level_{item}", fs, fs, "level"))
        return out

    def decoder_blocks(self) -> List[Tuple[str, int, int, str]]:
        """README.md:271-291."""
        out = []
        cin = self.feature_depth
        fs = cin // 2
        for d in range(self.depth, 0, -1):
            out.append((f"decoder.decoder.decoder-depth_{d}-upsample", cin, fs, "upsample"))
            for item in range(self.length - 1):
                out.append((f"decoder.decoder.decoder-depth_{d}-level_{item}", fs, fs, "level"))
            cin = fs
            fs = cin // 2
        out.append(("decoder.decoder.decoder-depth_0-reconstruction", cin,
                    self.reconstruction_channels, "level"))
        return out


@dataclass
class DiscriminatorSpec:
    num_stride_conv1: int = 1
    num_features_conv1: int = 64
    num_blocks: Sequence[int] = (1, 1, 1)
    num_strides_res: Sequence[int] = (1, 2, 2)
    num_features_res: Sequence[int] = (128, 256, 512)
    input_size: int = 256            # the reference hard-codes 256 (README.md:435)

    def res_blocks(self) -> List[Tuple[str, int, int, int]]:
        """(prefix, cin, cout, stride) -- README.md:445-448, 488-498."""
        out = []
        cin = self.num_features_conv1
        for i, (nb, st, planes) in enumerate(zip(self.num_blocks, self.num_strides_res,
                                                 self.num_features_res)):
            out.append((f"res_layers.{i}.0", cin, planes, st))
            for j in range(1, nb):
                out.append((f"res_layers.{i}.{j}", planes, planes, 1))
            cin = planes
        return out

    @property
    def linear_len(self) -> int:     # README.md:451-454
        side = self.input_size // self.num_stride_conv1 // 4
        prod = 1
        for s in self.num_strides_res:
            prod *= s
        side = side // prod
        return side * side * self.num_features_res[-1]


# ----------------------------------------------------------------------------------------
# parameter factories (distributional restatement of init_weights, README.md:700-707)
# ----------------------------------------------------------------------------------------
def _kaiming(shape, gen, dtype):
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(shape, generator=gen, dtype=dtype) * math.sqrt(2.0 / fan_in)


def _bn_entries(P, pre, c, dtype):
    P[pre + ".weight"] = torch.ones(c, dtype=dtype)
    P[pre + ".bias"] = torch.zeros(c, dtype=dtype)
    P[pre + ".running_mean"] = torch.zeros(c, dtype=dtype)
    P[pre + ".running_var"] = torch.ones(c, dtype=dtype)
    P[pre + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def make_generator_params(spec: GeneratorSpec, seed: int = 0, dtype=torch.float32) -> Params:
    """Kaiming-normal Conv2d weights, default-uniform ConvTranspose2d weights (the reference's
    init_weights skips ConvTranspose2d -- SURVEY.md App. E), BN gamma=1 beta=0."""
    g = torch.Generator().manual_seed(seed)
    P: Params = {}
    for pre, cin, cout, mode in spec.encoder_blocks() + spec.decoder_blocks():
        _bn_entries(P, pre + ".bn1", cin, dtype)
        if mode == "upsample":
            bound = 1.0 / math.sqrt(cout * 16)
            for nm in (".conv1.weight", ".shortcut.0.weight"):
                P[pre + nm] = (torch.rand((cin, cout, 4, 4), generator=g, dtype=dtype) * 2 - 1) * bound
        else:
            P[pre + ".conv1.weight"] = _kaiming((cout, cin, 3, 3), g, dtype)
            P[pre + ".shortcut.0.weight"] = _kaiming((cout, cin, 3, 3), g, dtype)
        _bn_entries(P, pre + ".shortcut.1", cout, dtype)
        _bn_entries(P, pre + ".bn2", cout, dtype)
        P[pre + ".conv2.weight"] = _kaiming((cout, cout, 3, 3), g, dtype)
    fd = spec.feature_depth
    for nm in ("log_var", "mu"):
        P[f"code_processor.{nm}.weight"] = _kaiming((fd, fd, 3, 3), g, dtype)
        P[f"code_processor.{nm}.bias"] = torch.zeros(fd, dtype=dtype)
    return P


def make_discriminator_params(spec: DiscriminatorSpec, seed: int = 1, dtype=torch.float32) -> Params:
    g = torch.Generator().manual_seed(seed)
    P: Params = {}
    P["conv1.weight"] = _kaiming((spec.num_features_conv1, 1, 3, 3), g, dtype)
    _bn_entries(P, "bn1", spec.num_features_conv1, dtype)

    def sn_conv(pre, cout, cin, k):
        P[pre + ".weight_orig"] = _kaiming((cout, cin, k, k), g, dtype)
        P[pre + ".weight_u"] = F.normalize(torch.randn(cout, generator=g, dtype=dtype), dim=0, eps=SN_EPS)
        P[pre + ".weight_v"] = F.normalize(torch.randn(cin * k * k, generator=g, dtype=dtype), dim=0, eps=SN_EPS)

    for pre, cin, cout, st in spec.res_blocks():
        _bn_entries(P, pre + ".bn1", cin, dtype)
        sn_conv(pre + ".conv1", cout, cin, 3)
        _bn_entries(P, pre + ".bn2", cout, dtype)
        sn_conv(pre + ".conv2", cout, cout, 3)
        if st != 1 or cin != cout:
            sn_conv(pre + ".shortcut.0", cout, cin, 1)
            _bn_entries(P, pre + ".shortcut.1", cout, dtype)
    dims = [spec.linear_len, 1024, 512, 256, 1]
    for i in range(4):
        P[f"linear_{i + 1}.weight"] = _kaiming((dims[i + 1], dims[i]), g, dtype)
        P[f"linear_{i + 1}.bias"] = torch.zeros(dims[i + 1], dtype=dtype)
    return P


_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")


def is_buffer_key(k: str) -> bool:
    return k.endswith(_BUFFER_SUFFIXES)


def trainable_keys(P: Params) -> List[str]:
    return [k for k in P if not is_buffer_key(k)]


def clone_params(P: Params, dtype=None, requires_grad=False) -> Params:
    out = {}
    for k, v in P.items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        if requires_grad and not is_buffer_key(k):
            t.requires_grad_(True)
        out[k] = t
    return out


# ----------------------------------------------------------------------------------------
# primitive restatements
# ----------------------------------------------------------------------------------------
def batch_norm(x: Tensor, P: Params, pre: str, training: bool) -> Tensor:
    """nn.BatchNorm2d(eps 1e-5, momentum 0.1): batch mean / biased var in train mode, running
    stats updated with the unbiased var."""
    if training:
        P[pre + ".num_batches_tracked"] += 1
    return F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"],
                        P[pre + ".weight"], P[pre + ".bias"], training, BN_MOMENTUM, BN_EPS)


def dropout_with_mask(x: Tensor, keep: Optional[Tensor], p: float, training: bool, channelwise=False):
    """nn.Dropout / nn.Dropout2d in train mode.  `keep` is a 0/1 keep-mask (elementwise, or
    (N,C,1,1) for Dropout2d); if None the torch RNG is used like the reference does."""
    if not training or p == 0.0:
        return x
    if keep is None:
        return F.dropout2d(x, p, True) if channelwise else F.dropout(x, p, True)
    return x * (keep.to(x.dtype) / (1.0 - p))


def spectral_normed_weight(P: Params, pre: str, training: bool) -> Tensor:
    """Legacy nn.utils.spectral_norm hook (n_power_iterations=1, dim=0, eps=1e-12): one power
    iteration per *training* forward updating weight_u / weight_v in place; u, v are
    constants for autograd; weight = weight_orig / (u^T W v).  README.md:378,383,387."""
    w = P[pre + ".weight_orig"]
    u, v = P[pre + ".weight_u"], P[pre + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS))
    uu, vv = u.clone(), v.clone()
    sigma = torch.dot(uu, torch.mv(wm, vv))
    return w / sigma


# ----------------------------------------------------------------------------------------
# blocks and stacks
# ----------------------------------------------------------------------------------------
def resblock_vae(x, P, pre, mode="level", res_mode="pre-activation", training=True,
                 keep_mask=None, dropout_prob=0.5):
    """README.md:174-197."""
    def conv_like(t, w):
        if mode == "level":
            return F.conv2d(t, w, None, 1, 1)
        if mode == "downsample":
            return F.conv2d(t, w, None, 2, 1)
        if mode == "upsample":
            return F.conv_transpose2d(t, w, None, 2, 1)
        raise ValueError(mode)

    def shortcut(t):
        s = conv_like(t, P[pre + ".shortcut.0.weight"])
        return batch_norm(s, P, pre + ".shortcut.1", training)

    act = lambda t: F.leaky_relu(t, SLOPE_G)
    if res_mode == "standard":
        out = conv_like(x, P[pre + ".conv1.weight"])
        out = batch_norm(out, P, pre + ".bn1", training)
        out = act(out)
        out = dropout_with_mask(out, keep_mask, dropout_prob, training)
        out = F.conv2d(out, P[pre + ".conv2.weight"], None, 1, 1)
        out = batch_norm(out, P, pre + ".bn2", training)
        out = out + shortcut(x)
        out = act(out)
    elif res_mode == "pre-activation":
        out = batch_norm(x, P, pre + ".bn1", training)
        out = act(out)
        out = dropout_with_mask(out, keep_mask, dropout_prob, training)
        out = conv_like(out, P[pre + ".conv1.weight"])
        out = batch_norm(out, P, pre + ".bn2", training)
        out = act(out)
        out = F.conv2d(out, P[pre + ".conv2.weight"], None, 1, 1)
        out = out + shortcut(x)
    else:
        raise ValueError(res_mode)
    return out


def generator_forward(x, P, spec: GeneratorSpec, training=True, is_training_code=True,
                      eps_noise=None, keep_masks: Optional[Dict[str, Tensor]] = None):
    """UnsupervisedGeneratorNetwork.forward (README.md:640-653) with is_vae=True.
    `keep_masks[prefix]` feeds each block's dropout; `eps_noise` replaces randn_like."""
    keep_masks = keep_masks or {}
    h = x
    for pre, _, _, mode in spec.encoder_blocks():
        h = resblock_vae(h, P, pre, mode, training=training, keep_mask=keep_masks.get(pre))
    # SpatialVAECodeProcessor.forward, README.md:573-586
    log_var = torch.clamp(F.conv2d(h, P["code_processor.log_var.weight"],
                                   P["code_processor.log_var.bias"], 1, 1), -50, 50)
    mu = F.conv2d(h, P["code_processor.mu.weight"], P["code_processor.mu.bias"], 1, 1)
    if is_training_code:
        std = torch.exp(0.5 * log_var)
        e = torch.randn_like(mu) if eps_noise is None else eps_noise.to(mu.dtype)
        z = mu + std * e
    else:
        z = mu
    h = z
    for pre, _, _, mode in spec.decoder_blocks():
        h = resblock_vae(h, P, pre, mode, training=training, keep_mask=keep_masks.get(pre))
    return h, mu, log_var


def decoder_forward(z, P, spec: GeneratorSpec, training=False, keep_masks=None):
    """UnsupervisedGeneratorNetwork.decode (README.md:661-664)."""
    keep_masks = keep_masks or {}
    h = z
    for pre, _, _, mode in spec.decoder_blocks():
        h = resblock_vae(h, P, pre, mode, training=training, keep_mask=keep_masks.get(pre))
    return h


def resblock_discriminator(x, P, pre, cin, cout, stride, res_mode="pre-activation", training=True,
                           keep_mask=None, dropout_prob=0.5):
    """README.md:396-419.  keep_mask is (N, cout, 1, 1) for the channel-wise Dropout2d."""
    act = lambda t: F.leaky_relu(t, SLOPE_D)
    has_sc = (stride != 1) or (cin != cout)

    def shortcut(t):
        if not has_sc:
            return t
        w = spectral_normed_weight(P, pre + ".shortcut.0", training)
        return batch_norm(F.conv2d(t, w, None, stride, 0), P, pre + ".shortcut.1", training)

    # the hooks fire in module-call order: conv1, conv2, shortcut.0
    if res_mode == "standard":
        out = F.conv2d(x, spectral_normed_weight(P, pre + ".conv1", training), None, stride, 1)
        out = dropout_with_mask(out, keep_mask, dropout_prob, training, channelwise=True)
        out = batch_norm(out, P, pre + ".bn1", training)
        out = act(out)
        out = F.conv2d(out, spectral_normed_weight(P, pre + ".conv2", training), None, 1, 1)
        out = batch_norm(out, P, pre + ".bn2", training)
        out = out + shortcut(x)
        out = act(out)
    else:
        out = batch_norm(x, P, pre + ".bn1", training)
        out = act(out)
        out = F.conv2d(out, spectral_normed_weight(P, pre + ".conv1", training), None, stride, 1)
        out = dropout_with_mask(out, keep_mask, dropout_prob, training, channelwise=True)
        out = batch_norm(out, P, pre + ".bn2", training)
        out = act(out)
        out = F.conv2d(out, spectral_normed_weight(P, pre + ".conv2", training), None, 1, 1)
        out = out + shortcut(x)
    return out


def discriminator_forward(img, P, spec: DiscriminatorSpec, training=True,
                          keep_masks: Optional[Dict[str, Tensor]] = None):
    """Discriminator.forward (README.md:465-486); returns the raw logit (B,1)."""
    keep_masks = keep_masks or {}
    act = lambda t: F.leaky_relu(t, SLOPE_D)
    out = F.conv2d(img, P["conv1.weight"], None, spec.num_stride_conv1, 1)
    out = act(batch_norm(out, P, "bn1", training))
    for pre, cin, cout, st in spec.res_blocks():
        out = resblock_discriminator(out, P, pre, cin, cout, st, training=training,
                                     keep_mask=keep_masks.get(pre))
    out = F.avg_pool2d(out, 4)
    out = out.reshape(out.size(0), -1)
    out = act(F.linear(out, P["linear_1.weight"], P["linear_1.bias"]))
    out = act(F.linear(out, P["linear_2.weight"], P["linear_2.bias"]))
    out = act(F.linear(out, P["linear_3.weight"], P["linear_3.bias"]))
    out = F.linear(out, P["linear_4.weight"], P["linear_4.bias"])
    return out


# ----------------------------------------------------------------------------------------
# losses (README.md:792-798, 816-831) and the BCE variant named by BASELINE.json north_star
# ----------------------------------------------------------------------------------------
def kl_divergence(mu, log_var):
    """-0.5 * SUM over batch and latent (README.md:822-825; the trailing .mean() is a no-op)."""
    return -0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp())


def reconstruction_loss(gen, real):
    """L1Loss + MSELoss, both mean-reduced (README.md:818-819, 921)."""
    return F.l1_loss(gen, real) + F.mse_loss(gen, real)


def d_loss_terms(d_real, d_fake, loss_mode):
    if loss_mode == "bce":
        return (F.binary_cross_entropy_with_logits(d_real, torch.ones_like(d_real)),
                F.binary_cross_entropy_with_logits(d_fake, torch.zeros_like(d_fake)))
    return -torch.mean(d_real), torch.mean(d_fake)            # README.md:792-793


def g_adv_loss(d_fake, loss_mode):
    if loss_mode == "bce":
        return F.binary_cross_entropy_with_logits(d_fake, torch.ones_like(d_fake))
    return -torch.mean(d_fake)                                # README.md:816


def gradient_penalty(P_d, spec_d, real, fake, alpha, keep_masks=None):
    """compute_gradient_penalty (README.md:717-739); `alpha` (B,1,1,1) replaces np.random."""
    inter = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    d_inter = discriminator_forward(inter, P_d, spec_d, True, keep_masks)
    grads = torch.autograd.grad(d_inter, inter, torch.ones_like(d_inter), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
    grads = grads.view(grads.size(0), -1)
    return ((grads.norm(2, dim=1) - 1) ** 2).mean()


# ----------------------------------------------------------------------------------------
# optimizers (torch.optim semantics, restated)
# ----------------------------------------------------------------------------------------
@dataclass
class OptState:
    kind: str = "adam"            # "adam" | "rmsprop"
    lr: float = 3e-4
    betas: Tuple[float, float] = (0.9, 0.999)
    eps: float = 1e-8
    alpha: float = 0.99
    weight_decay: float = 0.0
    step: int = 0
    m: Dict[str, Tensor] = field(default_factory=dict)
    v: Dict[str, Tensor] = field(default_factory=dict)


def optimizer_step(P: Params, grads: Dict[str, Tensor], st: OptState):
    st.step += 1
    with torch.no_grad():
        for k, g in grads.items():
            if g is None:
                continue
            p = P[k]
            if k not in st.v:
                st.v[k] = torch.zeros_like(p)
                st.m[k] = torch.zeros_like(p)
            if st.weight_decay != 0.0:
                g = g + st.weight_decay * p
            if st.kind == "adam":
                b1, b2 = st.betas
                st.m[k].mul_(b1).add_(g, alpha=1 - b1)
                st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                bc1 = 1 - b1 ** st.step
                bc2 = 1 - b2 ** st.step
                denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(st.eps)
                p.addcdiv_(st.m[k], denom, value=-st.lr / bc1)
            elif st.kind == "rmsprop":
                st.v[k].mul_(st.alpha).addcmul_(g, g, value=1 - st.alpha)
                p.addcdiv_(g, st.v[k].sqrt().add_(st.eps), value=-st.lr)
            else:
                raise ValueError(st.kind)


# ----------------------------------------------------------------------------------------
# one training iteration (README.md:775-834), randomness injected
# ----------------------------------------------------------------------------------------
def train_step(Pg: Params, Pd: Params, opt_g: OptState, opt_d: OptState, real: Tensor,
               spec_g: GeneratorSpec, spec_d: DiscriminatorSpec, *,
               eps_noise=None, g_masks=None, d_masks_real=None, d_masks_fake=None,
               d_masks_gp=None, d_masks_gen=None, loss_mode="bce",
               weights=(1.0, 10.0, 0.1), clip_value=0.01, lambda_gp=10.0, gp_alpha=None,
               return_grads=False):
    """Order is result-affecting and follows the reference exactly: G fwd -> D(real),
    D(fake.detach) [, GP] -> D backward -> D step [-> clamp] -> D(fake) with the UPDATED D ->
    G backward -> G step.  loss_mode: "bce" (north_star) | "wgan" (ref. critic loss + clamp,
    no GP) | "wgan_gp" (the reference as written: critic + 10*GP + clamp)."""
    for P in (Pg, Pd):
        for k in trainable_keys(P):
            P[k].requires_grad_(True)
            P[k].grad = None
    gkeys, dkeys = trainable_keys(Pg), trainable_keys(Pd)

    gen, mu, log_var = generator_forward(real, Pg, spec_g, True, True, eps_noise, g_masks)

    d_real = discriminator_forward(real, Pd, spec_d, True, d_masks_real)
    d_fake = discriminator_forward(gen.detach(), Pd, spec_d, True, d_masks_fake)
    real_loss, fake_loss = d_loss_terms(d_real, d_fake, loss_mode)
    d_loss = real_loss + fake_loss
    gp = None
    if loss_mode == "wgan_gp":
        gp = gradient_penalty(Pd, spec_d, real.detach(), gen.detach(), gp_alpha, d_masks_gp)
        d_loss = d_loss + lambda_gp * gp
    d_grads = torch.autograd.grad(d_loss, [Pd[k] for k in dkeys], allow_unused=True)
    d_grads = dict(zip(dkeys, d_grads))
    optimizer_step(Pd, d_grads, opt_d)
    if loss_mode in ("wgan", "wgan_gp"):
        with torch.no_grad():
            for k in dkeys:                      # every parameter, README.md:805-806
                Pd[k].clamp_(-clip_value, clip_value)

    d_gen = discriminator_forward(gen, Pd, spec_d, True, d_masks_gen)
    adv = g_adv_loss(d_gen, loss_mode)
    recon = reconstruction_loss(gen, real)
    kl = kl_divergence(torch.flatten(mu, 1), torch.flatten(log_var, 1))
    g_loss = weights[0] * adv + weights[1] * recon + weights[2] * kl
    g_grads = torch.autograd.grad(g_loss, [Pg[k] for k in gkeys], allow_unused=True)
    g_grads = dict(zip(gkeys, g_grads))
    optimizer_step(Pg, g_grads, opt_g)

    out = dict(d_loss=d_loss.detach(), g_loss=g_loss.detach(), recon=recon.detach(),
               kl=kl.detach(), real_loss=real_loss.detach(), fake_loss=fake_loss.detach(),
               adv=adv.detach(), gen=gen.detach(), mu=mu.detach(), log_var=log_var.detach(),
               d_real=d_real.detach(), d_fake=d_fake.detach(), d_gen=d_gen.detach())
    if gp is not None:
        out["gp"] = gp.detach()
    if return_grads:
        out["d_grads"] = {k: (None if g is None else g.detach()) for k, g in d_grads.items()}
        out["g_grads"] = {k: (None if g is None else g.detach()) for k, g in g_grads.items()}
    for P in (Pg, Pd):
        for k in trainable_keys(P):
            P[k].requires_grad_(False)
    return out


# ----------------------------------------------------------------------------------------
# Philox4x32-10 restatement (integer work; bit-exact against the CUDA generator).
# Counter = (lo32(idx/4), hi32(idx/4), offset_lo, offset_hi), key = (seed_lo, seed_hi);
# element idx uses output word idx%4.  Keep-mask: word >= p * 2^32.
# ----------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    import numpy as np

    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c = [np.asarray(x, dtype=np.uint32).copy() for x in counter]
    k0 = np.asarray(key[0], dtype=np.uint32).copy()
    k1 = np.asarray(key[1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c


def _philox_blocks(blk, seed: int, offset: int):
    import numpy as np

    n = blk.shape[0]
    ctr = [(blk & np.uint64(0xFFFFFFFF)).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
           np.full(n, offset & 0xFFFFFFFF, dtype=np.uint32),
           np.full(n, (offset >> 32) & 0xFFFFFFFF, dtype=np.uint32)]
    key = (np.full(n, seed & 0xFFFFFFFF, dtype=np.uint32),
           np.full(n, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint32))
    return np.stack(philox4x32_10(ctr, key), axis=1)        # [n][4]


def philox_uint32(n_elems: int, seed: int, offset: int, start: int = 0):
    """32-bit word per linear element index e in [start, start+n): block e>>2, word e&3
    (used by the Dropout2d per-(n,c) scale)."""
    import numpy as np

    idx = np.arange(start, start + n_elems, dtype=np.uint64)
    words = _philox_blocks(idx >> np.uint64(2), seed, offset)
    return words[np.arange(n_elems), (idx & np.uint64(3)).astype(np.int64)]


def philox_uniform(n_elems: int, seed: int, offset: int, start: int = 0):
    """float32 uniform [0,1): (word >> 8) * 2^-24 (the gradient-penalty interpolation weights)."""
    import numpy as np

    w = philox_uint32(n_elems, seed, offset, start)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def philox_keep_mask(n_elems: int, seed: int, offset: int, p: float, start: int = 0):
    """Elementwise-dropout keep mask (uint8), identical to the CUDA kernels: element e uses the
    16-bit half-word (e & 7) of Philox block (e >> 3); keep iff half-word >= floor(p * 65536)."""
    import numpy as np

    idx = np.arange(start, start + n_elems, dtype=np.uint64)
    words = _philox_blocks(idx >> np.uint64(3), seed, offset)
    j = (idx & np.uint64(7)).astype(np.int64)
    w = words[np.arange(n_elems), j >> 1]
    h = np.where((j & 1) == 1, w >> np.uint32(16), w & np.uint32(0xFFFF))
    thr = min(int(float(np.float32(p)) * 65536.0), 65535)
    return (h >= np.uint32(thr)).astype(np.uint8)


def philox_keep_scale2d(n: int, c: int, seed: int, offset: int, p: float, sample_offset: int = 0):
    """Dropout2d scale per (n, c): 0 or 1/(1-p); keep iff word >= floor(p * 2^32)."""
    import numpy as np

    w = philox_uint32(n * c, seed, offset, start=sample_offset * c)
    thr = min(int(float(np.float32(p)) * 4294967296.0), 0xFFFFFFFF)
    keep = (w >= np.uint32(thr)) if p > 0 else np.ones(n * c, dtype=bool)
    return (keep.astype(np.float32) / np.float32(1.0 - p)).reshape(n, c)
