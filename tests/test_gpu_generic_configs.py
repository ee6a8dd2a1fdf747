"""Generic configurations (-m gpu): the ranges of the reference's hyper-parameter search
(README.md:1084-1096) - feature sizes 8..64 in steps of 8, depth 1..n, length > 1, discriminators with other block counts -
run through the same kernels (CUDA-core path below 64 channels, tensor cores from 64 on) and match the fp64 oracle like
the BASELINE configuration does; shapes the library cannot take fail with a VgError (VG_EINVAL / VG_EUNSUPPORTED), never
with a wrong result; and the tile table (vae_gan_b200/tune.py, vg_conv_tune_*) changes kernel choices, not results.
"""
import ctypes as C

import pytest
import torch

from tests.gpu_util import dev
from tests.test_gpu_modules import _run_trainer_vs_oracle

pytestmark = pytest.mark.gpu


def V():
    import vae_gan_b200 as v
    return v


@pytest.fixture(scope="module", autouse=True)
def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import vae_gan_b200  # noqa: F401
    yield


# (feature_size, depth, length, image size, discriminator (num_blocks, strides, features) or None for experiment()'s)
GENERIC = [
    (16, 1, 1, 32, ((1, 1), (1, 2), (32, 64))),                 # depth 1, two-stage discriminator
    (24, 2, 2, 32, None),                                       # non-power-of-two width, two blocks per level
    (8, 3, 1, 64, ((2, 1, 1), (1, 2, 2), (16, 32, 64))),        # depth 3, a stage with two discriminator blocks
    (40, 2, 1, 32, None),                                       # 40 / 80 / 160 / 320 channels: C % 8 == 0, C % 64 != 0
    (32, 1, 3, 32, ((1, 3, 1), (1, 2, 2), (64, 64, 128))),      # length 3, num_blocks 3, repeated width
]


@pytest.mark.parametrize("fs,depth,length,S,disc", GENERIC)
def test_generic_config_fp32_train_step_vs_oracle(fs, depth, length, S, disc):
    # ONE iteration at the flat fp32 tolerance (forward, both backward passes, both optimizer steps, and `adv` evaluated after
    # the discriminator update).  A second iteration is not compared at this tolerance: after two Adam updates the
    # lr * sign(g) steps have amplified fp32 rounding noise on near-zero gradients (measured 1.5e-4 .. 1.2e-3 on `adv` for
    # the 320-channel and the length-3 configurations) - in ANY implementation, the fp32 CPU oracle included.
    _run_trainer_vs_oracle(torch.float32, "bce", "adam", B=2, S=S, fs=fs, steps=1, tol_loss=5e-5, max_bad_frac=5e-3,
                           depth=depth, length=length, disc=disc)


def test_generic_config_bf16_depth1_tensor_cores_vs_oracle():
    """depth 1 at feature size 64: 64 / 128-channel tensor-core layers at 48x48 / 24x24 that the BASELINE model never runs."""
    _run_trainer_vs_oracle(torch.bfloat16, "bce", "adam", B=4, S=48, fs=64, steps=1, tol_loss=2e-2, max_bad_frac=0.05,
                           depth=1, length=1, disc=((1, 1), (1, 2), (128, 256)))


def test_invalid_descriptors_are_errors_not_wrong_results():
    """A descriptor whose output size contradicts its geometry, a fused inference epilogue on an fp32 output, a tile-table
    entry with an impossible N tile: the library answers with an error code (raised as VgError) - it never guesses."""
    import vae_gan_b200.functional as VF
    from vae_gan_b200 import _lib, tune
    from vae_gan_b200._lib import VgError
    x = torch.randn(2, 8, 8, 64, device=dev()).to(torch.bfloat16)
    y = torch.empty(2, 8, 8, 64, device=dev(), dtype=torch.bfloat16)
    pk = torch.zeros(9 * 64 * 64, device=dev(), dtype=torch.bfloat16)
    s = _lib.stream_ptr()
    bad = _lib.VgConvDesc(2, 8, 8, 64, 7, 8, 64, 3, 3, 1, 1, 0, _lib.VG_BF16, _lib.VG_BF16)       # h_out should be 8
    with pytest.raises(VgError, match="inconsistent"):
        _lib.call("vg_conv_forward", C.byref(bad), x.data_ptr(), pk.data_ptr(), pk.data_ptr(), None, None, y.data_ptr(), None, s)
    good32 = _lib.VgConvDesc(2, 8, 8, 64, 8, 8, 64, 3, 3, 1, 1, 0, _lib.VG_BF16, _lib.VG_F32)
    y32 = torch.empty(2, 8, 8, 64, device=dev())
    ep = _lib.VgConvEpilogue(None, None, None, 0, 0.2, None, None, None, None, 1.0)                 # LeakyReLU epilogue on fp32 output
    with pytest.raises(VgError):
        _lib.call("vg_conv_forward_fused", C.byref(good32), x.data_ptr(), pk.data_ptr(), pk.data_ptr(), C.byref(ep), y32.data_ptr(), None, s)
    with pytest.raises(VgError):
        tune.set_entry((0, 2, 8, 8, 64, 64, 3, 1, 1, 0), 96, 1)
    with pytest.raises(VgError):
        tune.set_entry((0, 2, 8, 8, 64, 64, 3, 1, 1, 0), 64, 7)
    torch.cuda.synchronize()


def test_tile_table_changes_kernels_not_results():
    """Every (N tile, form) the table can pin computes the same convolution (bf16 rounding of an fp32 accumulator: the
    tile shape does not change the per-output reduction order inside one CTA, so results are bit-identical)."""
    import vae_gan_b200.functional as VF
    from vae_gan_b200 import _lib, tune
    v = V()
    g = torch.Generator().manual_seed(3)
    for (n, cin, cout, h, k, st, pad, tr) in [(4, 128, 256, 24, 3, 1, 1, 0), (8, 256, 128, 24, 3, 2, 1, 0), (4, 128, 64, 24, 4, 2, 1, 1)]:
        geom = VF.ConvGeom(k, st, pad, bool(tr))
        x = VF.as_act(torch.randn(n, cin, h, h, generator=g).to(dev()), torch.bfloat16)
        w = (torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), generator=g) / (cin * k * k) ** 0.5).to(dev())
        d, ho, wo = VF._conv_desc(x.shape, cout, geom, torch.bfloat16, torch.bfloat16)
        pk = torch.empty(w.numel(), dtype=torch.bfloat16, device=dev())
        pn = torch.empty_like(pk)
        s = _lib.stream_ptr()
        _lib.call("vg_conv_pack_weights", C.byref(d), w.data_ptr(), None, pk.data_ptr(), pn.data_ptr(), s)
        dy = VF.as_act(torch.randn(n, cout, ho, wo, generator=g).to(dev()), torch.bfloat16)
        key_f = (0, n, h, h, cin, cout, k, st, pad, tr)
        key_d = (1, n, h, h, cin, cout, k, st, pad, tr)
        outs = []
        try:
            for bn, form in [(0, 0)] + [(b, f) for b in (64, 128, 256) for f in (1, 2, 3)]:
                if bn and (cout % bn or cin % bn):
                    continue
                tune.set_entry(key_f, bn, form)
                tune.set_entry(key_d, bn, form)
                y = VF.empty_act(n, cout, ho, wo, torch.bfloat16, dev())
                dx = torch.empty_like(x)
                _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), pk.data_ptr(), pn.data_ptr(), None, None, y.data_ptr(), None, s)
                _lib.call("vg_conv_dgrad", C.byref(d), dy.data_ptr(), pk.data_ptr(), pn.data_ptr(), dx.data_ptr(), s)
                torch.cuda.synchronize()
                outs.append(((bn, form), y.clone(), dx.clone()))
        finally:
            tune.set_entry(key_f, 0, 0)
            tune.set_entry(key_d, 0, 0)
        assert len(outs) >= 4
        (_, y0, dx0) = outs[0]
        for cfg, y, dx in outs[1:]:
            # same per-output reduction order in every tile shape: expected bit-identical; bound = 1 bf16 ulp of the largest value
            for name, a, b in (("forward", y, y0), ("dgrad", dx, dx0)):
                err = float((a.float() - b.float()).abs().max())
                assert err <= 2.0 ** -7 * float(b.float().abs().max()), f"{name} differs under tile-table entry {cfg}: {err}"
                if err != 0.0:
                    print(f"note: {name} not bit-identical under {cfg}: max abs diff {err}")


def test_autotune_records_and_pins(tmp_path):
    """tune.record sees the tensor-core shapes of a step; autotune returns entries only for measured wins; save / load /
    apply round-trip; clear() restores the heuristics."""
    from vae_gan_b200 import tune
    v = V()
    with v.compute_dtype(torch.bfloat16):
        torch.manual_seed(0)
        G, D = v.build_vae_gan(feature_size=64, image_size=32)
        G, D = G.to(dev()).train(), D.to(dev()).train()
        tr = v.VaeGanTrainer(G, D)
        x = torch.rand(4, 1, 32, 32, device=dev())
        tr.step(x)
        keys = tune.record(lambda: tr.step(x))
    assert len(keys) >= 10 and all(len(k) == 10 for k in keys)
    assert any(k[0] == 1 for k in keys) and any(k[0] == 0 for k in keys)          # forward and dgrad shapes
    try:
        table = tune.autotune(keys=keys[:4], iters=5, min_gain=0.0)
        for k, e in table.items():
            assert e["bn"] in (64, 128, 256) and e["form"] in (1, 2, 3) and e["ms"] <= e["heuristic_ms"]
        path = tmp_path / "t.json"
        tune.save(table, path, meta={"test": True})
        assert tune.load(path) == table
        tune.apply(tune.load(path))
    finally:
        tune.clear()
        tune.apply_default()
