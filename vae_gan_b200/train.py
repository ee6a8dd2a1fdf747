"""The VAE-GAN training iteration (README.md:775-834) on this package's kernels.

Order follows the reference exactly (it is result-affecting): G forward -> D(real), D(fake.detach)
-> D backward -> D optimizer step [-> clamp in WGAN mode] -> D(fake) with the UPDATED D -> G
backward (D dgrad only) -> G optimizer step.  Differences, all result-neutral or named by
BASELINE.json north_star: BCE-with-logits adversarial loss and Adam are the default (the
reference's critic loss / RMSprop+clamp are `loss_mode="wgan"`, `optimizer="rmsprop"`; the notebook exactly as
written - critic loss + 10 x gradient penalty + clamp - is `loss_mode="wgan_gp"`, see gp.py); the unused
D weight gradients of the G step are not computed; losses stay on the device (no per-step sync).

Parameters, gradients and optimizer state of each network live in ONE flat fp32 buffer: weight
gradients are accumulated by the wgrad kernels straight into the flat gradient buffer, the
data-parallel all-reduce is one NCCL call per network, and the optimizer is one fused kernel.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as VF
from . import modules as M


class FlatParams:
    """Re-homes a module's parameters into one flat buffer (+ flat grad, + optimizer state)."""

    ALIGN = 64

    def __init__(self, net: nn.Module):
        params = [p for p in net.parameters()]
        assert params, "module has no parameters"
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.params, self.offsets, self.total = params, offs, total
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(params, offs):
                view = self.p[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = self.g[o:o + p.numel()].view(p.shape)
                p.grad = gview
                p._vg_grad_buf = gview       # wgrad kernels accumulate here directly

    def zero_grad(self):
        VF.call("vg_fill_zero", VF.ptr(self.g), self.g.numel() * 4, VF.stream_ptr())

    def set_requires_grad(self, flag: bool):
        for p in self.params:
            p.requires_grad_(flag)


class WeightPacks:
    """The GEMM-layout copies (`pack_kn`, `pack_nk`, compute dtype) of every convolution / Linear weight of one network
    in persistent buffers, re-made by ONE batched launch (functional.pack_weights_batched) instead of one pack kernel per
    layer per forward.  ConvFn picks them up through `weight._vg_packs` while a trainer iteration is active; the
    spectral norm (a per-forward scalar) is applied in the convolution epilogue, so the discriminator's three forwards of
    an iteration share the packs made after its optimizer step."""

    def __init__(self, net: nn.Module, dtype):
        self.dtype = dtype
        self.items = []
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                w = m.weight_orig if isinstance(m, M.SpectralNormConv2d) else m.weight
                transposed = isinstance(m, nn.ConvTranspose2d)
                wv = w
            elif isinstance(m, nn.Linear) and dtype == torch.bfloat16 and m.in_features % 64 == 0 and m.out_features % 64 == 0:
                w, transposed = m.weight, False           # runs as a 1x1 convolution on the tensor cores (functional.linear)
                wv = w.view(m.out_features, m.in_features, 1, 1)
            else:
                continue
            kn = torch.empty(w.numel(), dtype=dtype, device=w.device)
            nk = torch.empty(w.numel(), dtype=dtype, device=w.device)
            w._vg_packs = (kn, nk, dtype)
            self.items.append((wv.detach(), kn, nk, transposed))
        self.table = None

    def repack(self):
        self.table = VF.pack_weights_batched(self.items, self.dtype)


class GradBuckets:
    """Bucketed, backward-overlapped gradient all-reduce over one network's flat gradient buffer (SURVEY.md section 8e,
    collective 2; north_star: "bucketed NCCL allreduce ... overlapped with backward on side streams").

    The flat buffer is cut into contiguous buckets of ~`bucket_mb` in REVERSE parameter order (the order in which
    the backward finishes them: the discriminator's Linear layers - 80 % of its bytes - come first).  Every Function
    that accumulates a weight gradient into the flat buffer reports `use` at forward and `done` at backward
    (functional.note_use / note_done); when the last pending contribution of a bucket has been ENQUEUED on the compute
    stream, an event is recorded there, the communication stream waits for it and launches NCCL's all-reduce of that
    bucket, so the transfer runs under the remaining backward kernels.  `flush()` (called where the un-overlapped
    all-reduce used to be) reduces whatever is left and makes the compute stream wait for the communication stream.
    All of it is captured in the iteration's CUDA graph as a fork/join of the two streams."""

    def __init__(self, flat: "FlatParams", pg, comm_stream, bucket_mb: float = 25.0):
        self.flat, self.pg, self.comm = flat, pg, comm_stream
        cap = int(bucket_mb * (1 << 20) / 4)
        self.ranges = []               # (start, end) element ranges of flat.g, in launch order
        self.bucket_of = {}            # id(param) -> bucket index
        end = flat.total
        cur_start, cur_end, members = end, end, []
        for p, o in zip(reversed(flat.params), reversed(flat.offsets)):
            if cur_end - o > cap and members:
                self.ranges.append((cur_start, cur_end))
                for q in members:
                    self.bucket_of[id(q)] = len(self.ranges) - 1
                cur_end, members = cur_start, []
            cur_start = o
            members.append(p)
        if members:
            self.ranges.append((cur_start, cur_end))
            for q in members:
                self.bucket_of[id(q)] = len(self.ranges) - 1
        self.pending = [0] * len(self.ranges)
        self.sent = [False] * len(self.ranges)
        self.order = []                # buckets in the order they were launched (diagnostics / tests)
        self.armed = False

    def begin(self):
        """Start of a forward whose backward will be overlapped."""
        self.pending = [0] * len(self.ranges)
        self.sent = [False] * len(self.ranges)
        self.order = []
        self.armed = True

    def use(self, param):
        if self.armed:
            b = self.bucket_of.get(id(param))
            if b is not None:
                self.pending[b] += 1

    def done(self, param):
        if not self.armed:
            return
        b = self.bucket_of.get(id(param))
        if b is None:
            return
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.sent[b]:
            self._send(b)

    def _send(self, b):
        self.sent[b] = True
        self.order.append(b)
        s, e = self.ranges[b]
        if self.comm is None:          # host-only use (gloo tests): no streams to fork / join
            dist.all_reduce(self.flat.g[s:e], op=dist.ReduceOp.SUM, group=self.pg)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm.wait_event(ev)
        with torch.cuda.stream(self.comm):
            dist.all_reduce(self.flat.g[s:e], op=dist.ReduceOp.SUM, group=self.pg)

    def flush(self):
        """Reduce every bucket not sent yet (e.g. parameters without a fused gradient path), then join the streams."""
        for b in range(len(self.ranges)):
            if not self.sent[b]:
                self._send(b)
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.armed = False


class VaeGanTrainer:
    def __init__(self, generator: nn.Module, discriminator: nn.Module, *, loss_mode: str = "bce",
                 optimizer: str = "adam", lr: float = 3e-4, weights=(1.0, 10.0, 0.1), clip_value: float = 0.01,
                 weight_decay: Optional[float] = None, betas=(0.9, 0.999), process_group=None,
                 local_batch: Optional[int] = None, peer_syncbn: Optional[bool] = None, lambda_gp: float = 10.0,
                 n_critics: int = 1):
        assert loss_mode in ("bce", "wgan", "wgan_gp")
        assert optimizer in ("adam", "rmsprop")
        assert n_critics >= 1
        # README.md:812 `if i % n_critics == 0:` - the generator is updated on every n_critics-th iteration (the
        # first one included); its forward runs every iteration because the D step needs the fakes (README.md:789)
        self.n_critics = int(n_critics)
        self.iteration = 0
        self.G, self.D = generator, discriminator
        self.loss_mode, self.opt_kind, self.lr, self.weights = loss_mode, optimizer, lr, tuple(weights)
        self.clip = clip_value if loss_mode in ("wgan", "wgan_gp") else 0.0
        self.lambda_gp = lambda_gp
        self.gp_alpha_override = None        # tests inject the interpolation weights (B,1,1,1) here
        self.weight_decay = weight_decay if weight_decay is not None else (1e-5 if optimizer == "rmsprop" else 0.0)
        self.betas = betas
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if process_group is not None else 1
        self.rank = dist.get_rank(process_group) if process_group is not None else 0
        self.device = next(generator.parameters()).device
        VF._lib.ensure_device(self.device)
        self.fg = FlatParams(generator)
        self.fd = FlatParams(discriminator)
        self.opt_step = torch.zeros(1, dtype=torch.int64, device=self.device)   # device-side Adam t
        self.losses: Dict[str, torch.Tensor] = {}
        self.graph = None
        self.graph_d_only = None      # n_critics > 1: the iteration without the generator update
        self.static_real = None
        self.local_batch = local_batch
        self.peer = None
        if process_group is not None:
            VF.config.process_group = process_group
            if peer_syncbn is None:
                peer_syncbn = os.environ.get("VG_PEER_SYNCBN", "1") == "1"
            if peer_syncbn and self.world > 1:
                try:
                    from .dist import PeerExchange
                    self.peer = PeerExchange(process_group, self.device)
                except Exception as e:   # no P2P / IPC on this box: NCCL carries the statistics instead
                    if self.rank == 0:
                        print(f"[vae_gan_b200] NVLink peer exchange unavailable ({type(e).__name__}: {e}); SyncBN uses NCCL")
                    self.peer = None
                # all ranks must agree
                ok = torch.tensor([1 if self.peer is not None else 0], device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
                if int(ok) == 0:
                    self.peer = None
        # persistent weight packs, re-made by one batched launch per network (VG_PACK_CACHE=0: per-layer packs as before)
        self.packs_g = self.packs_d = None
        if os.environ.get("VG_PACK_CACHE", "1") == "1":
            self.packs_g = WeightPacks(generator, VF.config.compute_dtype)
            self.packs_d = WeightPacks(discriminator, VF.config.compute_dtype)
        self._packs_stale = True      # packs are re-made right after each optimizer step; stale only before the first iteration,
                                      # after a state restore, or after weights_changed()
        # data parallel: bucketed gradient all-reduce on a side stream, overlapped with the backward (GradBuckets).
        # The gradient-penalty mode accumulates part of its gradients through stock autograd (double backward), so it
        # keeps the single all-reduce after the backward.
        self.buckets_g = self.buckets_d = None
        if self.world > 1 and os.environ.get("VG_GRAD_BUCKETS", "1") == "1" and loss_mode != "wgan_gp":
            self.comm_stream = torch.cuda.Stream(device=self.device)
            mb = float(os.environ.get("VG_BUCKET_MB", "25"))
            self.buckets_g = GradBuckets(self.fg, process_group, self.comm_stream, mb)
            self.buckets_d = GradBuckets(self.fd, process_group, self.comm_stream, mb)
        # num_batches_tracked of every BatchNorm: G's are used once per iteration, D's three times
        self._nbt_g = [m.num_batches_tracked for m in generator.modules() if isinstance(m, nn.BatchNorm2d)]
        self._nbt_d = [m.num_batches_tracked for m in discriminator.modules() if isinstance(m, nn.BatchNorm2d)]

    # ------------------------------------------------------------------------------------------
    def _opt(self, flat: FlatParams, clamp: float):
        VF.optimizer_step(flat.p, flat.g, flat.m, flat.v, kind=self.opt_kind, lr=self.lr, betas=self.betas,
                          weight_decay=self.weight_decay, clamp=clamp, step_tensor=self.opt_step)

    def _allreduce(self, flat: FlatParams, buckets: Optional[GradBuckets] = None):
        if os.environ.get("VG_DIAG_NO_GRAD_AR", "0") == "1":     # timing diagnosis only (wrong numerics)
            if buckets is not None:
                buckets.armed = False
            return
        if self.world > 1:
            if buckets is not None:
                buckets.flush()          # most buckets were launched from inside the backward already
            else:
                dist.all_reduce(flat.g, op=dist.ReduceOp.SUM, group=self.pg)

    def _step_impl(self, real: torch.Tensor, g_step: bool = True):
        dev = self.device
        if self.world > 1:
            VF.config.sample_offset = self.rank * real.shape[0]
        adv_mode = 0 if self.loss_mode == "bce" else 1
        VF.rng.reset_sites()
        VF.rng.advance(dev)
        VF.call("vg_counter_add", VF.ptr(self.opt_step), 1, VF.stream_ptr())
        if self.G.training and self._nbt_g:
            torch._foreach_add_(self._nbt_g, 1)
        if self.D.training and self._nbt_d:
            torch._foreach_add_(self._nbt_d, (3 if self.loss_mode == "wgan_gp" else 2) + (1 if g_step else 0))
        VF.config.defer_num_batches_tracked = True
        VF.config.trainer_active = True
        VF.config.peer = self.peer
        if self.peer is not None:
            self.peer.reset()
        VF.arena.begin(dev)           # ONE memset for every fp64 accumulator of the iteration
        if self.packs_g is not None and self._packs_stale:
            self.packs_g.repack()
            self.packs_d.repack()
            self._packs_stale = False
        try:
            return self._iteration(real, adv_mode, g_step)
        finally:
            VF.arena.end()
            VF.config.defer_num_batches_tracked = False
            VF.config.trainer_active = False
            VF.config.peer = None

    def _iteration(self, real, adv_mode, g_step=True):
        with M._scope():
            with M._scope():          # depth >= 1 everywhere: modules hand over internal activations
                # ---- generator forward (graph kept for the G step) ----
                if self.buckets_g is not None and g_step:
                    self.buckets_g.begin()
                    VF.config.grad_tracker = self.buckets_g
                gen, mu, log_var = self.G(real)
                VF.config.grad_tracker = None
                # ---- discriminator step ----
                self.fd.set_requires_grad(True)
                self.fd.zero_grad()
                if self.buckets_d is not None and os.environ.get("VG_DIAG_NO_GRAD_AR", "0") != "1":
                    self.buckets_d.begin()
                    VF.config.grad_tracker = self.buckets_d
                d_real = self.D(real)
                d_fake = self.D(gen.detach())
                d_total, d_rl, d_fl = VF.DiscriminatorLossFn.apply(d_real, d_fake, adv_mode)
                gp_term = None
                if self.loss_mode == "wgan_gp":
                    # README.md:796-798: + lambda_gp * gradient_penalty(D, real.data, gen.data)
                    from .gp import gradient_penalty
                    b = real.shape[0]
                    alpha = self.gp_alpha_override
                    if alpha is None:
                        alpha = VF.philox_uniform(b, self.device, tag="gp_alpha").view(b, 1, 1, 1)
                    gp_term = gradient_penalty(self.D, real.detach(), gen.detach().float(), alpha)
                    # a mean over the LOCAL samples: 1/world of it is this rank's share of the global mean
                    d_total = d_total + (self.lambda_gp / self.world) * gp_term
                d_total.backward()
                VF.config.grad_tracker = None
                self._allreduce(self.fd, self.buckets_d)
                self._opt(self.fd, self.clip)
                if self.packs_d is not None:
                    self.packs_d.repack()          # D(gen) below and the next iteration run with the UPDATED discriminator (README.md:816)
                # ---- generator step (README.md:812: every n_critics-th iteration) ----
                d_gen = None
                if g_step:
                    self.fd.set_requires_grad(False)
                    self.fg.zero_grad()
                    d_gen = self.D(gen)
                    g_total, recon, kl, adv = VF.GeneratorLossFn.apply(gen, real, mu, log_var, d_gen, adv_mode,
                                                                       self.weights[0], self.weights[1], self.weights[2])
                    VF.config.grad_tracker = self.buckets_g          # G's uses were counted during its forward
                    g_total.backward()
                    VF.config.grad_tracker = None
                    self._allreduce(self.fg, self.buckets_g)
                    self._opt(self.fg, 0.0)
                    if self.packs_g is not None:
                        self.packs_g.repack()      # for the next iteration's generator forward
                    self.fd.set_requires_grad(True)
        self.losses = dict(d_loss=d_total.detach(), real_loss=d_rl.detach(), fake_loss=d_fl.detach())
        if g_step:
            self.losses.update(g_loss=g_total.detach(), recon=recon.detach(), kl=kl.detach(), adv=adv.detach())
            self._last_g_losses = {k: self.losses[k] for k in ("g_loss", "recon", "kl", "adv")}
        elif getattr(self, "_last_g_losses", None) is not None:
            # like the notebook's print (README.md:837), a skipped generator step reports the previous values
            self.losses.update(self._last_g_losses)
        if gp_term is not None:
            self.losses["gp"] = (gp_term / self.world).detach()
        self.last = dict(gen=gen.detach(), mu=mu.detach(), log_var=log_var.detach(), d_real=d_real.detach(),
                         d_fake=d_fake.detach())
        if d_gen is not None:
            self.last["d_gen"] = d_gen.detach()
        return self.losses

    # ------------------------------------------------------------------------------------------
    def _state_snapshot(self):
        """Everything one iteration mutates: parameters and optimizer state (flat buffers), Adam's step counter, the
        BatchNorm buffers, spectral-norm u / v, the Philox step counter."""
        snap = dict(flat=[(f, f.p.clone(), f.m.clone(), f.v.clone()) for f in (self.fg, self.fd)],
                    opt_step=self.opt_step.clone(), rng=VF.rng.step_tensor(self.device).clone(),
                    bufs=[(b, b.clone()) for net in (self.G, self.D) for b in net.buffers()],
                    iteration=self.iteration, last_g=getattr(self, "_last_g_losses", None))
        return snap

    def _state_restore(self, snap):
        with torch.no_grad():
            for f, p, m, v in snap["flat"]:
                f.p.copy_(p); f.m.copy_(m); f.v.copy_(v)
            self.opt_step.copy_(snap["opt_step"])
            VF.rng.step_tensor(self.device).copy_(snap["rng"])
            for b, saved in snap["bufs"]:
                b.copy_(saved)
        self.iteration = snap["iteration"]
        self._last_g_losses = snap["last_g"]
        self._packs_stale = True

    def capture(self, real_example: torch.Tensor, warmup: int = 3):
        """Capture the whole iteration (fwd + bwd + collectives + optimizers) in one CUDA graph (two when
        n_critics > 1: with and without the generator update).

        The warm-up iterations and the capture pass itself are NOT training steps: parameters, optimizer state,
        Adam's step counter, BatchNorm running statistics / num_batches_tracked, spectral-norm u / v and the
        Philox step are restored afterwards, so a graph-mode run starts from exactly the state an eager run
        starts from (capture does not execute kernels, but the warm-up does)."""
        self.static_real = real_example.clone()
        snap = self._state_snapshot()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self.static_real)
                if self.n_critics > 1:
                    self._step_impl(self.static_real, g_step=False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_impl(self.static_real)
        self._losses_full = self.losses
        if self.n_critics > 1:
            self.graph_d_only = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_d_only, pool=self.graph.pool()):
                self._step_impl(self.static_real, g_step=False)
            self._losses_d_only = self.losses
        self._state_restore(snap)
        self.weights_changed()
        torch.cuda.synchronize(self.device)
        return self

    def weights_changed(self):
        """Call after modifying parameters from OUTSIDE the trainer (load_state_dict, manual clamps, ...): the cached GEMM
        packs of the weights are re-made (immediately, so that a captured graph - which only re-packs after its own
        optimizer steps - sees them too)."""
        if self.packs_g is not None:
            self.packs_g.repack()
            self.packs_d.repack()
            self._packs_stale = False

    def step(self, real: torch.Tensor):
        g_step = (self.iteration % self.n_critics) == 0
        self.iteration += 1
        if self.graph is not None:
            if real is not self.static_real:
                self.static_real.copy_(real, non_blocking=True)
            if g_step:
                self.graph.replay()
                self.losses = self._losses_full
            else:
                self.graph_d_only.replay()
                self.losses = self._losses_d_only
            return self.losses
        return self._step_impl(real, g_step)

    def read_losses(self) -> Dict[str, float]:
        """One device->host read of the last step's scalars."""
        keys = list(self.losses)
        vals = torch.stack([self.losses[k].float() for k in keys])
        if self.world > 1:
            # each rank holds its share (means are normalised by the GLOBAL count, KL is a sum)
            vals = vals.clone()
            dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.pg)
        return dict(zip(keys, vals.cpu().tolist()))
