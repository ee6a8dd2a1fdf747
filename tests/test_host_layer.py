"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol the header
declares (no compute calls), ctypes struct layouts match the C structs, and the drop-in modules keep
the reference's state-dict keys."""
import ctypes as C
import re
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from vae_gan_b200 import _lib
    from vae_gan_b200.build import build
    build()
    lib = _lib.load()
    hdr = (ROOT / "include" / "vaegan_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 38
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert lib.vg_version() == 100


def test_ctypes_struct_layouts_match_header():
    """Compile a tiny C program against the header and compare sizeof / offsetof."""
    import subprocess, tempfile, json
    from vae_gan_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "vaegan_b200.h"
int main(void){
  printf("{\"VgConvDesc\":%zu,\"VgBnDesc\":%zu,\"VgLossDesc\":%zu,\"VgOptDesc\":%zu,"
         "\"bn_step_ptr\":%zu,\"bn_seed\":%zu,\"loss_xhat_dtype\":%zu,\"opt_grad_scale\":%zu}\n",
         sizeof(VgConvDesc), sizeof(VgBnDesc), sizeof(VgLossDesc), sizeof(VgOptDesc),
         offsetof(VgBnDesc, step_ptr), offsetof(VgBnDesc, seed), offsetof(VgLossDesc, xhat_dtype),
         offsetof(VgOptDesc, grad_scale));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "t.c"
        p.write_text(src)
        subprocess.check_call(["gcc", "-I", str(ROOT / "include"), str(p), "-o", str(Path(td) / "t")])
        got = json.loads(subprocess.check_output([str(Path(td) / "t")]))
    assert got["VgConvDesc"] == C.sizeof(_lib.VgConvDesc)
    assert got["VgBnDesc"] == C.sizeof(_lib.VgBnDesc)
    assert got["VgLossDesc"] == C.sizeof(_lib.VgLossDesc)
    assert got["VgOptDesc"] == C.sizeof(_lib.VgOptDesc)
    assert got["bn_step_ptr"] == _lib.VgBnDesc.step_ptr.offset
    assert got["bn_seed"] == _lib.VgBnDesc.seed.offset
    assert got["loss_xhat_dtype"] == _lib.VgLossDesc.xhat_dtype.offset
    assert got["opt_grad_scale"] == _lib.VgOptDesc.grad_scale.offset


def test_modules_keep_reference_state_dict_keys(golden_dir):
    from vae_gan_b200 import build_vae_gan
    s = torch.load(golden_dir / "structure_96.pt")
    G, D = build_vae_gan(image_size=96)
    assert sorted((k, tuple(v.shape)) for k, v in G.state_dict().items()) == sorted(s["g_keys"])
    assert sorted((k, tuple(v.shape)) for k, v in D.state_dict().items()) == sorted(s["d_keys"])
    assert sum(p.numel() for p in G.parameters()) == s["g_params"]
    assert sum(p.numel() for p in D.parameters()) == s["d_params"]
    # reference hard-codes 256x256 (README.md:435): the default input_size reproduces linear_len
    from vae_gan_b200 import Discriminator, ResBlockDiscriminator
    D256 = Discriminator(ResBlockDiscriminator, 1, 64, [1, 1, 1], [1, 2, 2], [128, 256, 512])
    assert D256.linear_len == 131072


def test_product_never_imports_the_oracle():
    for p in (ROOT / "vae_gan_b200").glob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from vae_gan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    import pytest
    with pytest.raises(_lib.VgError):
        _lib.load()


def test_bn_double_backward_closed_form_matches_autograd():
    """vae_gan_b200.gp.bn_double_backward (the math behind vg_bn_act_double_backward_*) against torch's own double
    backward of batch_norm + leaky_relu in fp64 (README.md:717-739 needs it for the gradient penalty)."""
    import torch
    import torch.nn.functional as F
    from vae_gan_b200 import gp
    torch.manual_seed(0)
    N, Cc, H, W = 3, 5, 4, 4
    for slope in (0.2, 1.0):
        x = torch.randn(N, Cc, H, W, dtype=torch.float64, requires_grad=True)
        gamma = torch.randn(Cc, dtype=torch.float64, requires_grad=True)
        beta = torch.randn(Cc, dtype=torch.float64, requires_grad=True)
        dy = torch.randn(N, Cc, H, W, dtype=torch.float64, requires_grad=True)
        G = torch.randn(N, Cc, H, W, dtype=torch.float64)
        y = F.leaky_relu(F.batch_norm(x, None, None, gamma, beta, True, 0.1, 1e-5), slope)
        (dx,) = torch.autograd.grad(y, x, dy, create_graph=True)
        g_dy, g_x, g_gamma = torch.autograd.grad((dx * G).sum(), [dy, x, gamma])
        mean, var = x.mean((0, 2, 3)), x.var((0, 2, 3), unbiased=False)
        o_dy, o_x, o_g = gp.bn_double_backward(G, dy.detach(), x.detach(), gamma.detach(), beta.detach(), mean.detach(),
                                               (var + 1e-5).rsqrt().detach(), slope, N * H * W, True)
        assert (o_dy - g_dy).abs().max() < 1e-12
        assert (o_x - g_x).abs().max() < 1e-12
        assert (o_g.double() - g_gamma).abs().max() < 1e-5      # returned in fp32


def test_tile_table_and_deterministic_switch_host_logic(tmp_path):
    """Host-side state of the library that needs no GPU: the tile table (set / record / list / clear, argument checks, the
    JSON round trip of vae_gan_b200.tune, the committed B200 table parses and applies) and the argument checks of
    vg_set_deterministic."""
    from vae_gan_b200 import _lib, tune
    from vae_gan_b200.build import build
    build()
    lib = _lib.load()
    # --- tile table
    tune.clear()
    key = (0, 32, 24, 24, 256, 256, 3, 1, 1, 0)
    tune.set_entry(key, 128, 1)
    tune.set_entry((1,) + key[1:], 256, 3)
    tune.set_entry(key, 0, 0)                       # removes the entry again
    for bad in ((96, 1), (64, 7), (512, 2), (128, -1)):
        try:
            tune.set_entry(key, *bad)
            raise AssertionError(f"accepted N tile / form {bad}")
        except _lib.VgError as e:
            assert "tile" in str(e) or "form" in str(e)
    cnt = C.c_int(-1)
    _lib.call("vg_conv_tune_record", 1)
    _lib.call("vg_conv_tune_record", 0)
    _lib.call("vg_conv_tune_seen", None, 0, C.byref(cnt))
    assert cnt.value == 0                           # nothing was launched while recording
    table = {key: {"bn": 128, "form": 1, "ms": 0.025, "heuristic_ms": 0.027}}
    path = tmp_path / "table.json"
    tune.save(table, path, meta={"gpu": "none"})
    assert tune.load(path) == table
    tune.apply(tune.load(path))
    shipped = tune.load(tune.DEFAULT_TABLE)
    assert len(shipped) >= 20
    for k, e in shipped.items():
        assert len(k) == 10 and k[0] in (0, 1) and e["bn"] in (64, 128, 256) and e["form"] in (1, 2, 3)
        assert e["ms"] <= 0.95 * e["heuristic_ms"]          # only >= 5 % measured wins are shipped
        n_out = k[4] if k[0] else k[5]
        assert n_out % e["bn"] == 0
    tune.apply(shipped)
    tune.clear()
    # --- deterministic mode: argument checks happen before anything touches a device
    assert lib.vg_get_deterministic() == 0
    assert lib.vg_set_deterministic(1, None, 0, None, 0) != 0
    assert b"scratch" in lib.vg_last_error()
    assert lib.vg_set_deterministic(1, C.c_void_p(16), 1 << 10, C.c_void_p(16), 4096) != 0       # scratch too small
    assert lib.vg_set_deterministic(1, C.c_void_p(24), 1 << 20, C.c_void_p(16), 4096) != 0       # scratch misaligned
    assert lib.vg_get_deterministic() == 0
    assert lib.vg_set_deterministic(0, None, 0, None, 0) == 0
