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


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2 - robust to the isolated LeakyReLU / |x| kink flips that change single rows
    of a gradient by O(1) in any finite-precision run."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    d = float(b.norm())
    return float((a - b).norm()) / (d if d > 0 else 1.0)


def assert_close(a, b, tol, name=""):
    e = relmax(a, b)
    assert e <= tol, f"{name}: max-normalised error {e:.3e} > {tol:.1e}"
    return e


def analytically_zero(want: torch.Tensor, global_max: float, thresh: float = 1e-9) -> bool:
    """Gradients that are exactly zero in exact arithmetic (e.g. the bias of a BatchNorm whose output
    only feeds other BatchNorms: a per-channel shift is cancelled) show up as pure rounding noise -
    ~1e-16 of the largest gradient in the fp64 oracle.  They carry no signal to compare."""
    return float(want.detach().abs().max()) <= thresh * global_max


def compare_grads(named_got, want: dict, tol: float, label="", zero_thresh: float = 1e-9, ref_lp: dict = None,
                  slack: float = 2.0, metric=relmax, skip=(), allowance: float = 0.0):
    """Per-tensor max-normalised comparison of parameter gradients against the (fp64) oracle.

    Conditioning-aware: a tensor passes if its error is <= tol, OR (when `ref_lp` is given) if it is
    no worse than `slack` x the error the REFERENCE ITSELF makes on that tensor when run in the same
    precision class (`ref_lp` = the oracle's gradients in fp32, or under torch bf16 autocast).  At
    random initialisation several gradients of this network amplify rounding noise by 1e2-1e5x
    (BatchNorm-backward cancellations), so a fixed tolerance is not attainable by ANY
    implementation in that precision; what parity means there is "as accurate as the reference".
    Returns ({name: (err, ref_err)}, skipped)."""
    def full(w):
        return w["sample"] if isinstance(w, dict) else w

    gmax = max(float(full(w).detach().abs().max()) for w in want.values())
    errs, bad, skipped = {}, [], []
    for k, got in named_got:
        w = want[k]
        if isinstance(w, dict):          # summarised golden tensor: strided sample
            got = got.detach().float().flatten().cpu()[:: w["stride"]][:4096]
            w = w["sample"]
        if analytically_zero(w, gmax, zero_thresh) or k in skip:
            skipped.append(k)
            continue
        e = metric(got, w)
        er = metric(ref_lp[k], w) if ref_lp is not None else 0.0
        errs[k] = (e, er)
        if e > tol and e > slack * er and e > allowance:
            bad.append((k, e, er))
    assert not bad, f"{label}: {len(bad)} gradient tensors above {tol:.1e} and above {slack}x the reference's own error: " + \
        ", ".join(f"{k}={e:.2e} (ref {er:.2e})" for k, e, er in bad[:12])
    return errs, skipped


def summarize_errs(errs: dict) -> str:
    worst = max(errs, key=lambda k: errs[k][0])
    over = sum(1 for e, er in errs.values() if e > er)
    return (f"worst {errs[worst][0]:.2e} on {worst} (reference's own error there {errs[worst][1]:.2e}); "
            f"{over}/{len(errs)} tensors less accurate than the low-precision reference")


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
