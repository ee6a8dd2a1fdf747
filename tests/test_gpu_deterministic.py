"""Deterministic-reduction mode (-m gpu; include/vaegan_b200.h vg_set_deterministic, SURVEY.md section 7 hard part 1).

With `vae_gan_b200.set_deterministic(True)` no cross-block floating-point sum uses atomics, so
* two runs of the trainer from the same state produce IDENTICAL bits (parameters of both networks, optimizer state,
  BatchNorm buffers, losses) - fp32 parity path, bf16 tensor-core path and the gradient-penalty mode;
* a CUDA-graph replay is bit-identical to the eager step from the same state (the test round 1 could only state with a
  tolerance, tests/test_gpu_modules.py::test_cuda_graph_replay_matches_eager);
* the results agree with the default (atomic) mode to rounding - the mode changes the ORDER of the sums, nothing else.
"""
import pytest
import torch

from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


def V():
    import vae_gan_b200 as v
    return v


@pytest.fixture()
def deterministic():
    v = V()
    v.set_deterministic(True, dev())
    assert v.is_deterministic()
    yield
    v.set_deterministic(False)
    assert not v.is_deterministic()


def _make(v, fs, S, **kw):
    torch.manual_seed(0)
    G, D = v.build_vae_gan(feature_size=fs, image_size=S)
    G, D = G.to(dev()).train(), D.to(dev()).train()
    v.rng.seed = 11
    v.rng.step_tensor(dev()).zero_()          # the Philox step counter is process-global
    return v.VaeGanTrainer(G, D, **kw)


def _state(tr):
    bufs = [b.clone() for net in (tr.G, tr.D) for b in net.buffers()]
    return [tr.fg.p.clone(), tr.fd.p.clone(), tr.fg.m.clone(), tr.fd.m.clone(), tr.fg.v.clone(), tr.fd.v.clone()] + bufs


def _run(v, cdt, B, S, fs, steps, graph=False, **kw):
    with v.compute_dtype(cdt):
        tr = _make(v, fs, S, **kw)
        gen = torch.Generator().manual_seed(5)
        xs = [torch.rand(B, 1, S, S, generator=gen).to(dev()) for _ in range(steps)]
        if graph:
            tr.capture(xs[0], warmup=1)
        losses = []
        for x in xs:
            tr.step(x)
            losses.append(tr.read_losses())
        torch.cuda.synchronize()
        return _state(tr), losses


def _assert_identical(a, b, what):
    sa, la = a
    sb, lb = b
    for i, (x, y) in enumerate(zip(sa, sb)):
        assert torch.equal(x, y), f"{what}: state tensor {i} differs in {int((x != y).sum())} of {x.numel()} elements"
    assert la == lb, f"{what}: losses differ {la} vs {lb}"


@pytest.mark.parametrize("cdt,B,S,fs", [(torch.float32, 4, 32, 64), (torch.bfloat16, 8, 96, 64), (torch.float32, 3, 32, 8)])
def test_reruns_are_bit_identical(deterministic, cdt, B, S, fs):
    v = V()
    if fs == 8:
        # feature_size 8 has BatchNorms over C % 8 == 0 only from width 8 on; every layer here is 8/16/32/1 wide: fine
        pass
    a = _run(v, cdt, B, S, fs, steps=3)
    b = _run(v, cdt, B, S, fs, steps=3)
    _assert_identical(a, b, f"rerun {cdt}")


def test_wgan_gp_reruns_are_bit_identical(deterministic):
    v = V()
    kw = dict(loss_mode="wgan_gp", optimizer="rmsprop")
    a = _run(v, torch.float32, 4, 32, 64, steps=2, **kw)
    b = _run(v, torch.float32, 4, 32, 64, steps=2, **kw)
    _assert_identical(a, b, "wgan_gp rerun fp32")
    a = _run(v, torch.bfloat16, 4, 96, 64, steps=2, **kw)
    b = _run(v, torch.bfloat16, 4, 96, 64, steps=2, **kw)
    _assert_identical(a, b, "wgan_gp rerun bf16")


@pytest.mark.parametrize("cdt,B,S", [(torch.float32, 2, 32), (torch.bfloat16, 8, 96)])
def test_graph_replay_is_bit_identical_to_eager(deterministic, cdt, B, S):
    v = V()
    eager = _run(v, cdt, B, S, 64, steps=3)
    graph = _run(v, cdt, B, S, 64, steps=3, graph=True)
    _assert_identical(eager, graph, f"graph vs eager {cdt}")


def test_deterministic_mode_changes_only_the_summation_order():
    v = V()
    v.set_deterministic(False)
    base_state, base_losses = _run(v, torch.float32, 4, 32, 64, steps=1)
    v.set_deterministic(True, dev())
    try:
        det_state, det_losses = _run(v, torch.float32, 4, 32, 64, steps=1)
    finally:
        v.set_deterministic(False)
    for k in base_losses[0]:
        a, b = base_losses[0][k], det_losses[0][k]
        assert abs(a - b) <= 1e-4 * max(1.0, abs(a)), (k, a, b)
    # one Adam step moves every weight by ~lr * sign(g): compare where the gradient is not at rounding level
    for x, y in zip(base_state[:2], det_state[:2]):
        frac = float(((x - y).abs() > 1.5e-4).float().mean())
        assert frac < 0.05, f"{frac:.3%} of parameters differ between the atomic and the ordered mode"


def test_unsupported_configuration_is_loud(deterministic):
    """BatchNorm over a channel count that is neither 1 nor a multiple of 8 has no ordered reduction: VG_EUNSUPPORTED."""
    import vae_gan_b200.functional as VF
    from vae_gan_b200._lib import VgError
    v = V()
    with v.compute_dtype(torch.float32):
        bn = torch.nn.BatchNorm2d(12).to(dev()).train()
        x = torch.randn(2, 12, 8, 8, device=dev())
        with pytest.raises(VgError, match="deterministic mode"):
            VF.bn_act(VF.to_act(x, torch.float32), bn, slope=1.0)
