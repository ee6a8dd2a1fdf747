"""Drop-in nn.Modules with the reference notebook's constructor signatures and state_dict keys
(SURVEY.md section 8b / Appendix B), whose forward/backward run on this package's sm_100a kernels.

The torch modules held inside (nn.Conv2d, nn.BatchNorm2d, nn.Linear, ...) are PARAMETER
CONTAINERS only - they give identical parameter names, shapes and default initialisation (so
the reference's `init_weights` and `load_state_dict(ref.state_dict())` work unchanged) - and
their own forward is never called.
"""
from __future__ import annotations

import contextlib
import threading
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import functional as VF
from .functional import ConvGeom

SLOPE_G = 0.01   # nn.LeakyReLU() default, README.md:172
SLOPE_D = 0.2    # README.md:394,437

_tls = threading.local()


@contextlib.contextmanager
def _scope():
    """Outermost module converts to user-facing fp32 on exit; nested modules hand the internal
    (channels_last, compute-dtype) activation straight to the next one."""
    depth = getattr(_tls, "depth", 0)
    _tls.depth = depth + 1
    try:
        yield depth == 0
    finally:
        _tls.depth = depth


def _pop_stats(x, training):
    s = getattr(x, "_vg_stats", None)
    return s if training else None


def _new_stats(c, device, training):
    return VF.zeros_f64(2 * c, device) if training else None


class _NoForward:
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the enclosing vae_gan_b200 module runs the kernels")


class ParamConv2d(_NoForward, nn.Conv2d):
    pass


class ParamConvTranspose2d(_NoForward, nn.ConvTranspose2d):
    pass


class SpectralNormConv2d(_NoForward, nn.Conv2d):
    """Same parameters/buffers as nn.utils.spectral_norm(nn.Conv2d(...)) (legacy hook version):
    `weight_orig` (Parameter), `weight_u` (c_out), `weight_v` (c_in*k*k); `.weight` is kept as a
    plain alias of weight_orig's storage so the reference's init_weights reaches it."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        w = self._parameters.pop("weight")
        self.register_parameter("weight_orig", w)
        object.__setattr__(self, "weight", w.data)
        with torch.no_grad():
            h = w.shape[0]
            wd = w.numel() // h
            u = nn.functional.normalize(w.new_empty(h).normal_(0, 1), dim=0, eps=VF.SN_EPS)
            v = nn.functional.normalize(w.new_empty(wd).normal_(0, 1), dim=0, eps=VF.SN_EPS)
        self.register_buffer("weight_u", u)
        self.register_buffer("weight_v", v)


def _conv_apply(mod, x, geom, *, colscale=None, stats_out=None, training=True, out_dtype=None):
    if isinstance(mod, SpectralNormConv2d):
        # `_vg_sn`: (sigma, u, v) of the batched power iteration the enclosing Discriminator ran for this forward
        return VF.conv(x, mod.weight_orig, mod.bias, geom=geom, sn=(mod.weight_u, mod.weight_v), colscale=colscale,
                       stats_out=stats_out, training=training, out_dtype=out_dtype, sn_pre=getattr(mod, "_vg_sn", None))
    return VF.conv(x, mod.weight, mod.bias, geom=geom, colscale=colscale, stats_out=stats_out, training=training,
                   out_dtype=out_dtype)


_MODE_GEOM = {
    "level": ConvGeom(3, 1, 1, False),
    "downsample": ConvGeom(3, 2, 1, False),
    "upsample": ConvGeom(4, 2, 1, True),
}


class ResBlockVAE(nn.Module):
    """README.md:126-197."""

    def __init__(self, in_channels, out_channels, mode="level", res_mode="pre-activation", dropout_prob=0.5):
        super().__init__()
        if mode not in _MODE_GEOM:
            raise ValueError(f"unknown mode {mode!r}")
        if res_mode not in ("pre-activation", "standard"):
            raise ValueError(f"unknown res_mode {res_mode!r}")
        self.res_mode = res_mode
        self.mode = mode
        self.bn1 = nn.BatchNorm2d(in_channels) if res_mode == "pre-activation" else nn.BatchNorm2d(out_channels)
        self.dropout = nn.Dropout(p=dropout_prob)
        if mode == "upsample":
            self.conv1 = ParamConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1, bias=False)
            sc = ParamConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1, bias=False)
        else:
            st = 1 if mode == "level" else 2
            self.conv1 = ParamConv2d(in_channels, out_channels, kernel_size=3, stride=st, padding=1, bias=False)
            sc = ParamConv2d(in_channels, out_channels, kernel_size=3, stride=st, padding=1, bias=False)
        self.shortcut = nn.Sequential(sc, nn.BatchNorm2d(out_channels))
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.conv2 = ParamConv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.activation_fun = nn.LeakyReLU(inplace=False)

    def forward(self, x):
        with _scope() as outer:
            x = VF.to_act(x)
            tr = self.training
            g1 = _MODE_GEOM[self.mode]
            g2 = _MODE_GEOM["level"]
            slope = self.activation_fun.negative_slope
            p = self.dropout.p
            dev = x.device
            c_out = self.bn2.num_features
            tag = f"ResBlockVAE:{self.mode}"
            if self.res_mode == "pre-activation":
                a, x = VF.bn_act_fork(x, self.bn1, slope=slope, drop_p=p, training=tr, sums=_pop_stats(x, tr), tag=tag)
                s2 = _new_stats(c_out, dev, tr)
                c1 = _conv_apply(self.conv1, a, g1, stats_out=s2, training=tr)
                b = VF.bn_act(c1, self.bn2, slope=slope, training=tr, sums=s2)
                c2 = _conv_apply(self.conv2, b, g2, training=tr)
                ss = _new_stats(c_out, dev, tr)
                sc = _conv_apply(self.shortcut[0], x, g1, stats_out=ss, training=tr)
                so = _new_stats(c_out, dev, tr)
                out = VF.bn_add(c2, sc, None, self.shortcut[1], slope=1.0, training=tr, sums_b=ss, stats_out=so)
            else:
                s1 = _new_stats(c_out, dev, tr)
                c1 = _conv_apply(self.conv1, x, g1, stats_out=s1, training=tr)
                a = VF.bn_act(c1, self.bn1, slope=slope, drop_p=p, training=tr, sums=s1, tag=tag)
                s2 = _new_stats(c_out, dev, tr)
                c2 = _conv_apply(self.conv2, a, g2, stats_out=s2, training=tr)
                ss = _new_stats(c_out, dev, tr)
                sc = _conv_apply(self.shortcut[0], x, g1, stats_out=ss, training=tr)
                so = _new_stats(c_out, dev, tr)
                out = VF.bn_add(c2, sc, self.bn2, self.shortcut[1], slope=slope, training=tr, sums_a=s2, sums_b=ss,
                                stats_out=so)
            if outer:
                return VF.from_act(out)
            if so is not None:
                out._vg_stats = so
            return out


class Encoder(nn.Module):
    """README.md:204-249."""

    def __init__(self, in_channels, depth, length, feature_size, block=ResBlockVAE):
        super().__init__()
        encoder = OrderedDict()
        for i in range(length):
            encoder["encoder-depth_0-level_" + str(i)] = block(in_channels, feature_size, mode="level")
            in_channels = feature_size
        for d in range(1, depth + 1):
            in_channels = feature_size
            feature_size *= 2
            encoder["encoder-depth_" + str(d) + "-downsample"] = block(in_channels, feature_size, mode="downsample")
            for item in range(0, length - 1):
                encoder["encoder-depth_" + str(d) + "-level_" + str(item)] = block(feature_size, feature_size, mode="level")
        self.encoder = nn.Sequential(encoder)

    def forward(self, x):
        with _scope() as outer:
            out = self.encoder(VF.to_act(x))
            return VF.from_act(out) if outer else out


class Decoder(nn.Module):
    """README.md:252-294."""

    def __init__(self, in_channels, depth, length, reconstruction_channels, block=ResBlockVAE):
        super().__init__()
        decoder = OrderedDict()
        feature_size = in_channels // 2
        for d in range(depth, 0, -1):
            decoder["decoder-depth_" + str(d) + "-upsample"] = block(in_channels, feature_size, mode="upsample")
            for item in range(0, length - 1):
                decoder["decoder-depth_" + str(d) + "-level_" + str(item)] = block(feature_size, feature_size, mode="level")
            in_channels = feature_size
            feature_size = in_channels // 2
        decoder["decoder-depth_0-reconstruction"] = block(in_channels, reconstruction_channels, mode="level")
        self.decoder = nn.Sequential(decoder)

    def forward(self, x):
        with _scope() as outer:
            out = self.decoder(VF.to_act(x))
            return VF.from_act(out) if outer else out


class ResBlockDiscriminator(nn.Module):
    """README.md:356-419."""

    def __init__(self, in_channels, out_channels, res_stride=1, res_mode="pre-activation", dropout_prob=0.5):
        super().__init__()
        if res_mode not in ("pre-activation", "standard"):
            raise ValueError(f"unknown res_mode {res_mode!r}")
        self.res_mode = res_mode
        self.res_stride = res_stride
        self.bn1 = nn.BatchNorm2d(in_channels) if res_mode == "pre-activation" else nn.BatchNorm2d(out_channels)
        self.conv1 = SpectralNormConv2d(in_channels, out_channels, kernel_size=3, stride=res_stride, padding=1, bias=False)
        self.dropout = nn.Dropout2d(p=dropout_prob)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.conv2 = SpectralNormConv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)
        if res_stride != 1 or out_channels != in_channels:
            self.shortcut = nn.Sequential(
                SpectralNormConv2d(in_channels, out_channels, kernel_size=1, stride=res_stride, bias=False),
                nn.BatchNorm2d(out_channels),
            )
        else:
            self.shortcut = nn.Sequential()
        self.activation_fun = nn.LeakyReLU(0.2, inplace=False)

    def forward(self, x):
        with _scope() as outer:
            x = VF.to_act(x)
            tr = self.training
            slope = self.activation_fun.negative_slope
            p = self.dropout.p
            dev = x.device
            c_out = self.bn2.num_features
            g1 = ConvGeom(3, self.res_stride, 1, False)
            g2 = ConvGeom(3, 1, 1, False)
            gs = ConvGeom(1, self.res_stride, 0, False)
            has_sc = len(self.shortcut) > 0
            scale = VF.dropout2d_scale(x.shape[0], c_out, p, dev, tag="ResBlockDiscriminator") if (tr and p > 0) else None
            if self.res_mode == "pre-activation":
                a, x = VF.bn_act_fork(x, self.bn1, slope=slope, training=tr, sums=_pop_stats(x, tr))
                s2 = _new_stats(c_out, dev, tr)
                c1 = _conv_apply(self.conv1, a, g1, colscale=scale, stats_out=s2, training=tr)
                b = VF.bn_act(c1, self.bn2, slope=slope, training=tr, sums=s2, out_colscale=scale)
                c2 = _conv_apply(self.conv2, b, g2, training=tr)
                so = _new_stats(c_out, dev, tr)
                if has_sc:
                    ss = _new_stats(c_out, dev, tr)
                    sc = _conv_apply(self.shortcut[0], x, gs, stats_out=ss, training=tr)
                    out = VF.bn_add(c2, sc, None, self.shortcut[1], slope=1.0, training=tr, sums_b=ss, stats_out=so)
                else:
                    out = VF.bn_add(c2, x, None, None, slope=1.0, training=tr, stats_out=so)
            else:
                s1 = _new_stats(c_out, dev, tr)
                c1 = _conv_apply(self.conv1, x, g1, colscale=scale, stats_out=s1, training=tr)
                a = VF.bn_act(c1, self.bn1, slope=slope, training=tr, sums=s1, out_colscale=scale)
                s2 = _new_stats(c_out, dev, tr)
                c2 = _conv_apply(self.conv2, a, g2, stats_out=s2, training=tr)
                so = _new_stats(c_out, dev, tr)
                if has_sc:
                    ss = _new_stats(c_out, dev, tr)
                    sc = _conv_apply(self.shortcut[0], x, gs, stats_out=ss, training=tr)
                    out = VF.bn_add(c2, sc, self.bn2, self.shortcut[1], slope=slope, training=tr, sums_a=s2, sums_b=ss,
                                    stats_out=so)
                else:
                    out = VF.bn_add(c2, x, self.bn2, None, slope=slope, training=tr, sums_a=s2, stats_out=so)
            if outer:
                return VF.from_act(out)
            if so is not None:
                out._vg_stats = so
            return out


class Discriminator(nn.Module):
    """README.md:422-498.  `input_size` (our only extra kwarg, default 256 like the reference's
    hard-coded value) sizes linear_1 for other image sizes."""

    def __init__(self, block, num_stride_conv1: int, num_features_conv1: int, num_blocks, num_strides_res,
                 num_features_res, input_size: int = 256):
        super().__init__()
        assert len(num_blocks) == len(num_strides_res) == len(num_features_res), "length of lists must be equal"
        isz = np.array([1, input_size, input_size])
        self.block = block
        self.activation_fun = nn.LeakyReLU(0.2, inplace=False)
        self.num_stride_conv1 = num_stride_conv1
        self.in_planes = num_features_conv1
        self.conv1 = ParamConv2d(int(isz[0]), num_features_conv1, kernel_size=3, stride=num_stride_conv1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(num_features_conv1)
        res_layers = []
        for i in range(len(num_blocks)):
            res_layers.append(self._make_layer(planes=num_features_res[i], num_blocks=num_blocks[i], stride=num_strides_res[i]))
        self.res_layers = nn.Sequential(*res_layers)
        linear_len = isz // num_stride_conv1 // 4
        linear_len = np.floor_divide(linear_len, np.prod(num_strides_res))
        linear_len[0] = 1
        self.linear_len = int(np.prod(linear_len) * num_features_res[-1])
        self.linear_1 = nn.Linear(self.linear_len, 1024)
        self.linear_2 = nn.Linear(1024, 512)
        self.linear_3 = nn.Linear(512, 256)
        self.linear_4 = nn.Linear(256, 1)

    def _make_layer(self, planes, num_blocks, stride):
        layers = [self.block(in_channels=self.in_planes, out_channels=planes, res_stride=stride)]
        for _ in np.arange(num_blocks - 1):
            layers.append(self.block(in_channels=planes, out_channels=planes))
        self.in_planes = planes
        return nn.Sequential(*layers)

    def forward(self, img):
        with _scope():
            x = VF.to_act(img)
            tr = self.training
            slope = self.activation_fun.negative_slope
            c1 = self.bn1.num_features
            # ONE batched power iteration for every spectral-normed convolution of this forward (the hooks of the
            # reference fire once per module call, in any order: the weights are independent)
            sn_convs = [m for m in self.res_layers.modules() if isinstance(m, SpectralNormConv2d)]
            pre = VF.spectral_norm_batched([(m.weight_orig, m.weight_u, m.weight_v) for m in sn_convs], tr)
            for m, p in zip(sn_convs, pre):
                m._vg_sn = p
            try:
                s1 = _new_stats(c1, x.device, tr)
                out = _conv_apply(self.conv1, x, ConvGeom(3, self.num_stride_conv1, 1, False), stats_out=s1, training=tr)
                out = VF.bn_act(out, self.bn1, slope=slope, training=tr, sums=s1)
                out = self.res_layers(out)
            finally:
                for m in sn_convs:
                    m._vg_sn = None
            out = VF.AvgPoolFlattenFn.apply(out, 4)
            wd = VF.config.compute_dtype
            out = VF.linear(out, self.linear_1.weight, self.linear_1.bias, slope, wd)
            out = VF.linear(out, self.linear_2.weight, self.linear_2.bias, slope, wd)
            out = VF.linear(out, self.linear_3.weight, self.linear_3.bias, slope, wd)
            out = VF.linear(out, self.linear_4.weight, self.linear_4.bias, 1.0, wd)
            return out


class SpatialVAECodeProcessor(nn.Module):
    """README.md:522-597."""

    def __init__(self, feature_depth, is_training):
        super().__init__()
        self.log_vars_upper_bound = 50
        self.log_vars_lower_bound = -self.log_vars_upper_bound
        self.is_training = is_training
        self.log_var = ParamConv2d(feature_depth, feature_depth, kernel_size=3, stride=1, padding=1)
        self.mu = ParamConv2d(feature_depth, feature_depth, kernel_size=3, stride=1, padding=1)
        self.eps_override = None      # tests inject the reparameterisation noise here

    def forward(self, x):
        with _scope() as outer:
            x = VF.to_act(x)
            g = _MODE_GEOM["level"]
            lv_raw = VF.conv(x, self.log_var.weight, self.log_var.bias, geom=g, out_dtype=torch.float32)
            mu = VF.conv(x, self.mu.weight, self.mu.bias, geom=g, out_dtype=torch.float32)
            eps = None
            if self.is_training:
                if self.eps_override is not None:
                    eps = VF.as_act(self.eps_override.to(mu.device), torch.float32)
                else:
                    eps = VF.philox_normal(tuple(mu.shape), mu.device)
            z, lv = VF.ReparamFn.apply(mu, lv_raw, eps, bool(self.is_training), x.dtype)
            if outer:
                z, mu, lv = VF.from_act(z), VF.from_act(mu), VF.from_act(lv)
            return z, mu, lv

    def encode(self, x):
        with _scope() as outer:
            x = VF.to_act(x)
            mu = VF.conv(x, self.mu.weight, self.mu.bias, geom=_MODE_GEOM["level"], out_dtype=torch.float32)
            return VF.from_act(mu) if outer else mu      # fp32 (the latent mean is not rounded to the compute dtype)

    def decode(self, x):
        return x

    def set_is_training(self, is_training):
        self.is_training = is_training


class UnsupervisedGeneratorNetwork(nn.Module):
    """README.md:600-667."""

    def __init__(self, encoder, code_processor, decoder, is_vae):
        super().__init__()
        self.is_vae = is_vae
        self.is_training = True
        self.encoder = encoder
        self.code_processor = code_processor
        self.decoder = decoder

    def forward(self, x):
        with _scope() as outer:
            x = self.encoder(x)
            if self.is_vae:
                x, mu, log_var = self.code_processor(x)
            else:
                x = self.code_processor(x)
            x = self.decoder(x)
            if outer:
                x = VF.from_act(x)
                if self.is_vae:
                    mu, log_var = VF.from_act(mu), VF.from_act(log_var)
            if self.is_vae:
                return x, mu, log_var
            return x

    def encode(self, x):
        with _scope() as outer:
            x = self.encoder(x)
            mu = self.code_processor.encode(x)
            return VF.from_act(mu) if outer else mu

    def decode(self, x):
        with _scope() as outer:
            x = self.code_processor.decode(x)
            x = self.decoder(x)
            return VF.from_act(x) if outer else x

    def set_is_training(self, is_training):
        self.code_processor.set_is_training(is_training)


def init_weights(module):
    """README.md:700-707 (note: nn.ConvTranspose2d is not an nn.Conv2d, so it keeps torch's
    default init, exactly as in the reference)."""
    if isinstance(module, nn.Conv2d) or isinstance(module, nn.Linear):
        nn.init.kaiming_normal_(module.weight)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.BatchNorm2d):
        module.weight.data.fill_(1)
        module.bias.data.zero_()


def build_vae_gan(depth=2, length=1, feature_size=64, image_size=96, disc_params=None, is_vae=True):
    """`experiment()`'s model construction (README.md:882-907)."""
    feature_depth = feature_size * (2 ** depth)
    G = UnsupervisedGeneratorNetwork(
        encoder=Encoder(in_channels=1, depth=depth, length=length, feature_size=feature_size),
        decoder=Decoder(in_channels=feature_depth, depth=depth, length=length, reconstruction_channels=1),
        code_processor=SpatialVAECodeProcessor(feature_depth=feature_depth, is_training=True),
        is_vae=is_vae,
    )
    if disc_params is None:
        disc_params = dict(num_stride_conv1=1, num_features_conv1=feature_size, num_blocks=[1, 1, 1],
                           num_strides_res=[1, 2, 2],
                           num_features_res=[2 * feature_size, 4 * feature_size, 8 * feature_size])
    D = Discriminator(block=ResBlockDiscriminator, input_size=image_size, **disc_params)
    G.apply(init_weights)
    D.apply(init_weights)
    return G, D
