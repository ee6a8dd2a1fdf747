"""Host-logic dry run (CPU, no GPU, no kernels): the whole Python side of a training iteration - modules,
autograd.Functions, the trainer's flat buffers / stats arena / weight-pack cache / batched spectral norm - executed on
CPU tensors with every C-ABI call replaced by a checker that validates the call against the ctypes prototype
(argument count, pointer-vs-scalar kinds, struct types) and does nothing else.  The numbers are garbage by
construction; what this pins is that every entry point is called with a well-formed argument list on every code path
(BCE / WGAN / WGAN-GP, eval forward, n_critics, stock-autograd use after a trainer) before a GPU minute is spent."""
import ctypes as C

import pytest
import torch

import vae_gan_b200 as V
from vae_gan_b200 import _lib


class _CallChecker:
    def __init__(self):
        self.calls = {}

    def __call__(self, name, *args):
        assert name in _lib._PROTOS, f"unknown entry point {name}"
        _, argtypes = _lib._PROTOS[name]
        assert len(args) == len(argtypes), f"{name}: {len(args)} arguments for a prototype of {len(argtypes)}"
        for i, (a, t) in enumerate(zip(args, argtypes)):
            if t is C.c_void_p:
                ok = a is None or isinstance(a, int) or isinstance(a, (C.Array, C.c_void_p)) or hasattr(a, "_obj")
                assert ok, f"{name}: argument {i} should be a pointer, got {type(a)}"
            elif isinstance(t, type) and issubclass(t, C._Pointer):
                want = t._type_
                ok = a is None or (hasattr(a, "_obj") and isinstance(a._obj, want)) or isinstance(a, C.Array)
                assert ok, f"{name}: argument {i} should be byref({want.__name__}), got {type(a)}"
            elif t in (C.c_int, C.c_longlong, C.c_ulonglong, C.c_size_t):
                assert isinstance(a, (int, bool)), f"{name}: argument {i} should be an integer, got {type(a)} ({a!r})"
            elif t in (C.c_float, C.c_double):
                assert isinstance(a, (int, float)), f"{name}: argument {i} should be a number, got {type(a)}"
        self.calls[name] = self.calls.get(name, 0) + 1


@pytest.fixture()
def dry(monkeypatch):
    chk = _CallChecker()
    import vae_gan_b200.functional as VF
    import vae_gan_b200.gp as GP
    import vae_gan_b200.sampling as SA
    for mod in (_lib, VF, GP, SA):
        monkeypatch.setattr(mod, "call", chk, raising=False)
    monkeypatch.setattr(SA, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "ensure_device", lambda device: None)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(VF, "stream_ptr", lambda: 0)
    monkeypatch.setattr(GP, "stream_ptr", lambda: 0, raising=False)
    # CPU tensors have no torch.cuda.current_device(): the Philox step tensor is keyed by device
    monkeypatch.setattr(VF.PhiloxRng, "step_tensor",
                        lambda self, device: self._step.setdefault("cpu", torch.zeros(1, dtype=torch.int64)))
    yield chk
    VF.config.process_group = None
    VF.config.trainer_active = False
    VF.arena.active = False
    VF.arena.buf = None


def _models(fs=8, S=32, dtype=torch.float32):
    torch.manual_seed(0)
    G, D = V.build_vae_gan(feature_size=fs, image_size=S)
    return G.train(), D.train()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("loss_mode,opt,n_critics", [("bce", "adam", 1), ("wgan", "rmsprop", 2), ("wgan_gp", "rmsprop", 1)])
def test_trainer_iteration_host_logic(dry, dtype, loss_mode, opt, n_critics):
    B, S = 2, 32
    x = torch.rand(B, 1, S, S)
    with V.compute_dtype(dtype):
        G, D = _models()
        tr = V.VaeGanTrainer(G, D, loss_mode=loss_mode, optimizer=opt, n_critics=n_critics)
        for _ in range(3):
            losses = tr.step(x)
        assert "d_loss" in losses and "g_loss" in losses
    c = dry.calls
    # one batched pack per network at the top of an iteration + one for D after its update, never a per-layer pack
    assert c.get("vg_conv_pack_weights_batched", 0) >= 6
    if loss_mode != "wgan_gp":
        assert c.get("vg_conv_pack_weights", 0) == 0, "a convolution missed the pack cache"
        assert c.get("vg_spectral_norm_sigma", 0) == 0, "a spectral-normed convolution missed the batched power iteration"
    assert c.get("vg_spectral_norm_sigma_batched", 0) >= 6
    assert c.get("vg_bn_finalize", 0) == 0 and c.get("vg_bn_param_grads", 0) == 0, "finalize / param-grad kernels are folded in"
    assert c.get("vg_bn_act_forward_fused", 0) > 0 and c.get("vg_bn_act_backward_apply_fused", 0) > 0
    assert c.get("vg_optimizer_step", 0) >= 4


def test_modules_standalone_and_eval_host_logic(dry):
    B, S = 2, 32
    x = torch.rand(B, 1, S, S)
    with V.compute_dtype(torch.bfloat16):
        G, D = _models()
        y, mu, lv = G(x)
        assert y.shape == (B, 1, S, S) and mu.shape == (B, 32, S // 4, S // 4)
        logits = D(x)
        assert logits.shape == (B, 1)
        (logits.sum() + y.sum()).backward()
        assert all(p.grad is not None for p in D.parameters())
        G.eval(); D.eval()
        G.set_is_training(False)
        with torch.no_grad():
            G.decode(G.encode(x))
            D(x)
        # stand-alone blocks (per-weight spectral norm path, both res_modes)
        blk = V.ResBlockDiscriminator(8, 16, res_stride=2, res_mode="standard").train()
        blk(torch.rand(B, 8, 16, 16)).sum().backward()
        blk2 = V.ResBlockVAE(8, 16, mode="upsample", res_mode="standard").train()
        blk2(torch.rand(B, 8, 8, 8)).sum().backward()
    assert dry.calls.get("vg_spectral_norm_sigma", 0) == 3      # the stand-alone block's three convolutions
    assert dry.calls.get("vg_nhwc_to_nchw", 0) > 0               # module outputs are NCHW-contiguous


def test_stock_loop_after_trainer_host_logic(dry):
    B, S = 2, 32
    x = torch.rand(B, 1, S, S)
    with V.compute_dtype(torch.float32):
        G, D = _models()
        tr = V.VaeGanTrainer(G, D)
        tr.step(x)
        opt = torch.optim.SGD(D.parameters(), lr=0.0)
        opt.zero_grad()
        D(x).sum().backward()
        assert all(p.grad is not None for p in D.parameters()), "gradients swallowed by the trainer's hidden flat buffer"
        # loss Functions outside the trainer scale with the incoming gradient
        import vae_gan_b200.functional as VF
        n0 = dry.calls.get("vg_scale", 0)
        dr = torch.randn(B, 1, requires_grad=True)
        df = torch.randn(B, 1, requires_grad=True)
        t, _, _ = VF.DiscriminatorLossFn.apply(dr, df, 0)
        (0.5 * t).backward()
        assert dry.calls.get("vg_scale", 0) == n0 + 2


def test_input_pipeline_table_and_pack_items(dry):
    """ctypes tables of the batched entry points are well formed (sizes, field order)."""
    assert C.sizeof(_lib.VgPackItem) == 3 * 8 + 4 * 4
    assert C.sizeof(_lib.VgSnItem) == 6 * 8 + 2 * 4
    assert C.sizeof(_lib.VgBnChannel) == 4 * 8 + 8 + 3 * 8 + 2 * 4


def test_folded_sampler_host_logic(dry):
    from vae_gan_b200.sampling import FoldedGenerator
    with V.compute_dtype(torch.bfloat16):
        torch.manual_seed(0)
        G, _ = V.build_vae_gan(feature_size=64, image_size=32)
        G = G.eval()
        G.set_is_training(False)
        fg = FoldedGenerator(G)
        assert [b.tc for b in fg.enc + fg.dec] == [False, True, True, True, True, False]
        assert fg.decode(torch.randn(2, 256, 8, 8)).shape == (2, 1, 32, 32)
        assert fg.encode(torch.rand(2, 1, 32, 32)).shape == (2, 256, 8, 8)
        assert fg.reconstruct(torch.rand(2, 1, 32, 32)).shape == (2, 1, 32, 32)
    assert dry.calls.get("vg_fold_bn_into_conv", 0) == 8 and dry.calls.get("vg_conv_forward_fused", 0) > 0
