"""Shared helpers for the -m gpu parity tests (oracle = checker only)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from oracle import vaegan_oracle as O


def dev():
    return torch.device("cuda", 0)


def relmax(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| normalised by max |b| (the per-tensor max-abs metric SURVEY.md section 7 item 7 asks for)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = float(b.abs().max())
    if denom == 0.0:
        denom = 1.0
    return float((a - b).abs().max()) / denom


def assert_close(a, b, tol, name=""):
    e = relmax(a, b)
    assert e <= tol, f"{name}: max-normalised error {e:.3e} > {tol:.1e}"
    return e


def nchw(t: torch.Tensor) -> torch.Tensor:
    """Internal activation -> contiguous fp32 NCHW on the CPU."""
    return t.detach().float().contiguous().cpu()


def philox_mask_nchw(shape, seed, offset, p, step=0, sample_offset=0):
    n, c, h, w = shape
    start = sample_offset * c * h * w
    m = O.philox_keep_mask(n * h * w * c, seed, offset + 65536 * step, p, start=start).reshape(n, h, w, c)
    return torch.from_numpy(m).permute(0, 3, 1, 2).contiguous()


def philox_keep2d(n, c, seed, offset, p, step=0, sample_offset=0):
    sc = O.philox_keep_scale2d(n, c, seed, offset + 65536 * step, p, sample_offset=sample_offset)
    return torch.from_numpy((sc > 0).astype(np.uint8)).reshape(n, c, 1, 1)


def generator_masks(spec_g, B, S, seed, first_site=0, step=0, skip_after_encoder=0, sample_offset=0, p=0.5):
    """Keep-masks per generator block as the product draws them: one Philox site per block in forward
    order; `skip_after_encoder`=1 when the reparameterisation noise also consumed a site."""
    masks, h = {}, S
    site = first_site
    n_enc = len(spec_g.encoder_blocks())
    for i, (pre, cin, cout, mode) in enumerate(spec_g.encoder_blocks() + spec_g.decoder_blocks()):
        if i == n_enc:
            site += skip_after_encoder
        masks[pre] = philox_mask_nchw((B, cin, h, h), seed, site, p, step, sample_offset)
        site += 1
        h = h // 2 if mode == "downsample" else (h * 2 if mode == "upsample" else h)
    return masks, site


def discriminator_masks(spec_d, B, seed, first_site, step=0, sample_offset=0, p=0.5):
    masks = {}
    site = first_site
    for pre, cin, cout, st in spec_d.res_blocks():
        masks[pre] = philox_keep2d(B, cout, seed, site, p, step, sample_offset)
        site += 1
    return masks, site


def load_params_into(module: torch.nn.Module, P):
    sd = {k: v.detach().clone() for k, v in P.items()}
    module.load_state_dict(sd, strict=True)
    return module
