"""torch.autograd.Functions over the C ABI (include/vaegan_b200.h).

Activations are torch tensors of LOGICAL shape (N, C, H, W) stored channels_last (physically
NHWC) in the compute dtype (bf16 tensor-core path, or fp32 parity path).  Every Function
launches this library's kernels on torch's current CUDA stream; there is no torch/cuDNN/cuBLAS
fallback - a missing library or an unsupported configuration raises.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import VgBnDesc, VgConvDesc, VgLossDesc, VgOptDesc, call, ptr, stream_ptr, vg_dtype

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SN_EPS = 1e-12


# ----------------------------------------------------------------------------------------------
# global configuration / randomness
# ----------------------------------------------------------------------------------------------
class _Config:
    compute_dtype = torch.bfloat16
    process_group = None          # data-parallel group for SyncBN statistics (None = single GPU)
    sample_offset = 0             # global index of this rank's first sample
    defer_num_batches_tracked = False   # the trainer bumps every BN's counter with ONE foreach kernel per step
    peer = None                   # dist.PeerExchange: NVLink one-shot exchange of the SyncBN sums (else NCCL)
    trainer_active = False        # inside VaeGanTrainer._iteration: loss backward is called with grad 1 (no rescale kernels)
    grad_tracker = None           # train.GradBuckets: told when a parameter's fused gradient is used / final (DP overlap)


config = _Config()


@contextlib.contextmanager
def compute_dtype(dtype):
    old = config.compute_dtype
    config.compute_dtype = dtype
    try:
        yield
    finally:
        config.compute_dtype = old


class PhiloxRng:
    """Counter-based randomness: every dropout / noise site draws stream id `offset` (a host
    counter, identical on every rank and every replay) + 65536 * step (a DEVICE counter), so a
    captured CUDA graph produces fresh masks on each replay and tests can regenerate any mask."""

    def __init__(self, seed: int = 0x5EED5EED):
        self.seed = seed
        self.site = 0
        self._step = {}
        self.record = False
        self.trace = []

    def reset_sites(self):
        self.site = 0

    def next_site(self, tag: str = "", shape=None) -> int:
        s = self.site
        self.site += 1
        if self.record:
            self.trace.append((tag, s, shape))
        return s

    def step_tensor(self, device) -> torch.Tensor:
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in self._step:
            self._step[key] = torch.zeros(1, dtype=torch.int64, device=device)
        return self._step[key]

    def advance(self, device, inc: int = 1):
        t = self.step_tensor(device)
        call("vg_counter_add", ptr(t), inc, stream_ptr())


rng = PhiloxRng()


def _world():
    pg = config.process_group
    if pg is None:
        return 1
    return dist.get_world_size(pg)


import os as _os
_DIAG_NO_SYNCBN = bool(int(_os.environ.get("VG_DIAG_NO_SYNCBN", "0")))      # timing diagnosis only (wrong numerics)


def _allreduce_sums(t: torch.Tensor):
    if config.process_group is not None and _world() > 1 and not _DIAG_NO_SYNCBN:
        if config.peer is not None:
            config.peer.allreduce_(t, rng.step_tensor(t.device))
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=config.process_group)


def live_grad_buf(param):
    """The trainer's flat gradient view for `param` (FlatParams sets `param._vg_grad_buf` and `param.grad` to the SAME
    storage) - or None when the fused accumulation must not be used: the parameter is driven by stock machinery
    (e.g. `optimizer.zero_grad()` replaced / dropped `.grad`), in which case the backward returns ordinary gradients
    and autograd accumulates them.  Evaluated at BACKWARD time."""
    if param is None:
        return None
    buf = getattr(param, "_vg_grad_buf", None)
    if buf is None:
        return None
    owner = getattr(param, "_vg_owner", param)
    g = owner.grad
    if g is None or g.data_ptr() != buf.data_ptr():
        return None
    return buf


def note_use(ctx, indexed_params):
    """Forward of a Function whose backward will ACCUMULATE into the trainer's flat gradient buffer: one pending
    contribution per (parameter, use) whose gradient autograd will ask for (ctx.needs_input_grad).  The data-parallel
    trainer all-reduces a bucket of the flat buffer as soon as every contribution to it has been launched
    (train.GradBuckets).  Returns the owners noted; the backward hands exactly that list to note_done."""
    t = config.grad_tracker
    if t is None:
        return ()
    noted = []
    for idx, p in indexed_params:
        if p is not None and ctx.needs_input_grad[idx] and getattr(p, "_vg_grad_buf", None) is not None:
            owner = getattr(p, "_vg_owner", p)
            t.use(owner)
            noted.append(owner)
    return tuple(noted)


def note_done(noted):
    """Backward counterpart of note_use: the kernels adding these contributions are enqueued on the current stream."""
    t = config.grad_tracker
    if t is None or not noted:
        return
    for owner in noted:
        t.done(owner)


# ----------------------------------------------------------------------------------------------
# tensor helpers
# ----------------------------------------------------------------------------------------------
def is_act(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous()


def empty_act(n, c, h, w, dtype, device):
    return torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def as_act(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """Make `t` an internal activation (channels_last, `dtype`) using our own kernels."""
    dtype = dtype or t.dtype
    _lib.ensure_device(t.device)
    n, c, h, w = t.shape
    if is_act(t):
        if t.dtype == dtype:
            return t
        out = empty_act(n, c, h, w, dtype, t.device)
        call("vg_cast", ptr(t), vg_dtype(t.dtype), ptr(out), vg_dtype(dtype), t.numel(), stream_ptr())
        return out
    if not t.is_contiguous():
        t = t.contiguous()
    if t.dtype != torch.float32:
        tmp = torch.empty(t.shape, dtype=torch.float32, device=t.device)
        call("vg_cast", ptr(t), vg_dtype(t.dtype), ptr(tmp), _lib.VG_F32, t.numel(), stream_ptr())
        t = tmp
    out = empty_act(n, c, h, w, dtype, t.device)
    call("vg_nchw_to_nhwc", ptr(t), n, c, h, w, vg_dtype(dtype), ptr(out), stream_ptr())
    return out


class StatsArena:
    """Zero-initialised fp64 accumulators (BatchNorm sums, loss sums) for ONE training iteration, carved from a
    single buffer that the trainer clears with ONE memset at the start of the iteration - instead of one
    memset node per accumulator (~150 per iteration).  Outside a trainer iteration (`active` False) every
    request falls back to its own zero-filled tensor."""

    CAPACITY = 1 << 19          # doubles (4 MB): ~150 accumulators of <= 2 * 1024 channels (+ the 5C second-order sums)

    def __init__(self):
        self.buf = None
        self.pos = 0
        self.active = False
        self._lock = threading.Lock()

    def begin(self, device):
        if self.buf is None or self.buf.device != device:
            self.buf = torch.empty(self.CAPACITY, dtype=torch.float64, device=device)
        call("vg_fill_zero", ptr(self.buf), self.buf.numel() * 8, stream_ptr())
        self.pos = 0
        self.active = True

    def end(self):
        self.active = False

    def take(self, n, device):
        if not self.active or self.buf is None or self.buf.device != device:
            return None
        with self._lock:
            n2 = (n + 1) // 2 * 2                   # keep every slice 16-byte aligned
            if self.pos + n2 > self.buf.numel():
                return None
            t = self.buf[self.pos:self.pos + n]
            self.pos += n2
        return t


arena = StatsArena()


def zeros_f64(n, device):
    t = arena.take(n, device)
    if t is not None:
        return t
    t = torch.empty(n, dtype=torch.float64, device=device)
    call("vg_fill_zero", ptr(t), t.numel() * 8, stream_ptr())
    return t


def zeros_f32(shape, device):
    t = torch.empty(shape, dtype=torch.float32, device=device)
    call("vg_fill_zero", ptr(t), t.numel() * 4, stream_ptr())
    return t


class ToActFn(Function):
    """Module-boundary conversion: any (N,C,H,W) tensor -> internal activation dtype/layout."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.in_dtype = x.dtype
        ctx.in_cl = is_act(x)
        return as_act(x, dtype)

    @staticmethod
    def backward(ctx, g):
        dt = ctx.in_dtype if ctx.in_dtype in (torch.float32, torch.bfloat16) else torch.float32
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty): stay differentiable
            return (g if (is_act(g) and g.dtype == dt) else ToActFn.apply(g, dt)), None
        return as_act(g, dt), None


def to_act(x, dtype=None):
    dtype = dtype or config.compute_dtype
    if is_act(x) and x.dtype == dtype:
        return x
    return ToActFn.apply(x, dtype)


class FromActFn(Function):
    """Module-boundary conversion back to what the reference modules return: a CONTIGUOUS NCHW fp32 tensor (so
    reference-style code such as `out.view(out.size(0), -1)` works on it)."""

    @staticmethod
    def forward(ctx, y):
        _lib.ensure_device(y.device)
        n, c, h, w = y.shape
        ctx.in_dtype = y.dtype
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=y.device)
        call("vg_nhwc_to_nchw", ptr(y), vg_dtype(y.dtype), n, c, h, w, ptr(out), stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty): stay differentiable
            return ToActFn.apply(g, ctx.in_dtype)
        return as_act(g, ctx.in_dtype)


def from_act(y, dtype=torch.float32):
    """Internal activation -> user-facing tensor: contiguous NCHW fp32, like the reference modules' outputs.
    (A 1-channel tensor is bit-identical in NCHW and NHWC, so only the dtype changes there.)"""
    if y.shape[1] > 1 and is_act(y) and not y.is_contiguous() and dtype == torch.float32:
        return FromActFn.apply(y)
    if y.dtype == dtype:
        return y
    return ToActFn.apply(y, dtype)


# ----------------------------------------------------------------------------------------------
# convolution
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvGeom:
    k: int
    stride: int
    pad: int
    transposed: bool = False


def _conv_desc(x_shape, c_out, g: ConvGeom, act_dtype, out_dtype):
    n, c_in, h, w = x_shape
    if g.transposed:
        ho = (h - 1) * g.stride - 2 * g.pad + g.k
        wo = (w - 1) * g.stride - 2 * g.pad + g.k
    else:
        ho = (h + 2 * g.pad - g.k) // g.stride + 1
        wo = (w + 2 * g.pad - g.k) // g.stride + 1
    d = VgConvDesc(n, h, w, c_in, ho, wo, c_out, g.k, g.k, g.stride, g.pad, int(g.transposed),
                   vg_dtype(act_dtype), vg_dtype(out_dtype))
    return d, ho, wo


# VG_SN_FUSED_WGRAD=0: the two-call sequence vg_conv_wgrad + vg_spectral_norm_backward (A/B, and the reference for the test)
_SN_FUSED_WGRAD = _os.environ.get("VG_SN_FUSED_WGRAD", "1") == "1"


class ConvFn(Function):
    """nn.Conv2d / nn.ConvTranspose2d (optionally spectral-normed, optionally followed by the
    per-(n,c) Dropout2d scale) - README.md:148-170, 378-387, 441, 556-571.

    weight is the fp32 parameter in torch layout.  With `sn=(u, v)` the legacy spectral-norm
    hook semantics apply: one power iteration (training) updating u, v in place, weight/sigma.
    `stats_out` (double[2*c_out], zeroed) receives sum / sum-of-squares of the output for the
    BatchNorm that follows.  When `colscale` is given the matching BnActFn must be called with
    out_colscale=colscale: it returns the gradient w.r.t. the UNSCALED conv output.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, sn_u, sn_v, colscale, geom, out_dtype, stats_out, training, sn_pre):
        _lib.ensure_device(x.device)
        assert is_act(x), "ConvFn expects a channels_last activation"
        dev = x.device
        c_out = weight.shape[1] if geom.transposed else weight.shape[0]
        out_dtype = out_dtype or x.dtype
        d, ho, wo = _conv_desc(x.shape, c_out, geom, x.dtype, out_dtype)
        s = stream_ptr()
        sigma = None
        u_saved = v_saved = None
        w = weight.detach()
        if not w.is_contiguous():
            w = w.contiguous()
        if sn_pre is not None:
            # the discriminator ran ONE batched power iteration for all of its weights at the top of its forward
            sigma, u_saved, v_saved = sn_pre
        elif sn_u is not None:
            rows, cols = w.shape[0], w.numel() // w.shape[0]
            sigma = torch.empty(1, dtype=torch.float32, device=dev)
            ws = torch.empty(rows + cols + 4, dtype=torch.float32, device=dev)
            call("vg_spectral_norm_sigma", ptr(w), rows, cols, ptr(sn_u), ptr(sn_v), int(training), SN_EPS,
                 ptr(sigma), ptr(ws), s)
            u_saved, v_saved = sn_u.clone(), sn_v.clone()
        packs = cached_packs(weight, x.dtype)
        if packs is not None:
            # the trainer packed every weight of the network once after the optimizer step; W / sigma is applied in
            # the epilogue (conv is linear), so the same pack serves every forward until the next update
            pack_kn, pack_nk = packs
            sigma_ep = sigma
        else:
            numel = w.numel()
            pack_kn = torch.empty(numel, dtype=x.dtype, device=dev)
            pack_nk = torch.empty(numel, dtype=x.dtype, device=dev)
            call("vg_conv_pack_weights", C.byref(d), ptr(w), ptr(sigma), ptr(pack_kn), ptr(pack_nk), s)
            sigma_ep = None
        y = empty_act(x.shape[0], c_out, ho, wo, out_dtype, dev)
        b = bias.detach() if bias is not None else None
        call("vg_conv_forward_scaled", C.byref(d), ptr(x), ptr(pack_kn), ptr(pack_nk), ptr(b), ptr(colscale), ptr(sigma_ep), 0,
             ptr(y), ptr(stats_out), s)
        ctx.d = d
        ctx.geom = geom
        ctx.has_bias = bias is not None
        ctx.has_sn = sigma is not None
        ctx.sigma_in_epilogue = sigma_ep is not None
        ctx.wshape = tuple(weight.shape)
        ctx.weight_ref, ctx.bias_ref = weight, bias          # live_grad_buf() is evaluated at backward time
        ctx.noted = note_use(ctx, ((1, weight), (2, bias)))
        ctx.save_for_backward(x, pack_kn, pack_nk, w if ctx.has_sn else None, sigma, u_saved, v_saved, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pack_kn, pack_nk, w, sigma, u, v, weight = ctx.saved_tensors
        d = ctx.d
        if torch.is_grad_enabled():          # create_graph=True: differentiable input gradient (vae_gan_b200/gp.py)
            from . import gp
            # NOTE: this pass yields the INPUT gradient only (autograd.grad(..., inputs=x, create_graph=True) as in
            # compute_gradient_penalty); ctx.needs_input_grad cannot tell whether a weight gradient was requested,
            # so higher-order derivatives with respect to parameters are not available through create_graph.
            if not ctx.needs_input_grad[0]:
                return (None,) * 11
            w_eff = gp.sn_effective_weight(weight, u, v) if ctx.has_sn else weight
            dx = gp.ConvDgradFn.apply(dy, w_eff, d, x.dtype)
            return (dx,) + (None,) * 10
        s = stream_ptr()
        dy = as_act(dy, x.dtype)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = empty_act(d.n, d.c_in, d.h_in, d.w_in, x.dtype, x.device)
            call("vg_conv_dgrad_scaled", C.byref(d), ptr(dy), ptr(pack_kn), ptr(pack_nk),
                 ptr(sigma) if ctx.sigma_in_epilogue else None, 0, ptr(dx), s)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            wbuf = live_grad_buf(ctx.weight_ref)          # trainer's flat gradient view (fused accumulation) or None
            bbuf = live_grad_buf(ctx.bias_ref)
            fused = wbuf is not None
            direct = fused and not ctx.has_sn
            if ctx.has_bias:
                db = bbuf if bbuf is not None else zeros_f32((d.c_out,), x.device)
            if ctx.has_sn and x.dtype == torch.bfloat16 and _SN_FUSED_WGRAD and _lib.load().vg_conv_wgrad_sn_supported(C.byref(d)):
                # tensor-core layer: weight gradient + spectral-norm backward in one call (the gradient stays in the kernel's
                # packed layout and is corrected while it is transposed: no unpack pass, no temporary dW, one memset)
                dw = wbuf if fused else zeros_f32(ctx.wshape, x.device)
                numel = 1
                for n_ in ctx.wshape:
                    numel *= n_
                ws = torch.empty(numel + 4, dtype=torch.float32, device=x.device)
                call("vg_conv_wgrad_sn", C.byref(d), ptr(x), ptr(dy), ptr(w), ptr(u), ptr(v), ptr(sigma), ptr(dw), ptr(db), ptr(ws), s)
                if fused:
                    dw = None
                if bbuf is not None:
                    db = None
                note_done(ctx.noted)
                return dx, dw, db, None, None, None, None, None, None, None, None
            dwh = wbuf if direct else zeros_f32(ctx.wshape, x.device)
            ws = torch.empty(dwh.numel(), dtype=torch.float32, device=x.device) if x.dtype == torch.bfloat16 else None
            call("vg_conv_wgrad", C.byref(d), ptr(x), ptr(dy), ptr(dwh), ptr(db), ptr(ws), s)
            if ctx.has_sn:
                rows, cols = ctx.wshape[0], dwh.numel() // ctx.wshape[0]
                dw = wbuf if fused else zeros_f32(ctx.wshape, x.device)
                ws = torch.empty(4, dtype=torch.float32, device=x.device)
                call("vg_spectral_norm_backward", ptr(dwh), ptr(w), ptr(u), ptr(v), ptr(sigma), rows, cols,
                     ptr(dw), ptr(ws), s)
            else:
                dw = dwh
            if fused:
                dw = None
            if bbuf is not None:
                db = None
        note_done(ctx.noted)
        return dx, dw, db, None, None, None, None, None, None, None, None


def conv(x, weight, bias=None, *, geom: ConvGeom, sn=None, colscale=None, out_dtype=None, stats_out=None,
         training=True, sn_pre=None):
    u, v = sn if sn is not None else (None, None)
    return ConvFn.apply(x, weight, bias, u, v, colscale, geom, out_dtype, stats_out, training, sn_pre)


def cached_packs(weight, dtype):
    """(pack_kn, pack_nk) made by the trainer's batched pack (train.WeightPacks) - valid only inside a trainer iteration
    (the trainer re-packs after every optimizer step) and for the compute dtype they were made in."""
    if not config.trainer_active:
        return None
    pk = getattr(weight, "_vg_packs", None)
    if pk is None or pk[2] != dtype:
        return None
    return pk[0], pk[1]


def pack_weights_batched(items, dtype):
    """items: [(weight fp32 tensor (torch layout), pack_kn, pack_nk, transposed)] -> ONE launch (per 32 weights)."""
    if not items:
        return None
    _lib.ensure_device(items[0][0].device)
    table = (_lib.VgPackItem * len(items))()
    for i, (w, kn, nk, transposed) in enumerate(items):
        taps = 1
        for dim in w.shape[2:]:
            taps *= dim
        table[i] = _lib.VgPackItem(ptr(w), ptr(kn), ptr(nk), w.shape[0], w.shape[1], taps, int(transposed))
    call("vg_conv_pack_weights_batched", table, len(items), vg_dtype(dtype), stream_ptr())
    return table


def spectral_norm_batched(weights, training):
    """One batched power iteration (README.md:378,383,387; legacy nn.utils.spectral_norm hook: n_power_iterations 1,
    eps 1e-12) for a list of (weight_orig, weight_u, weight_v): three launches for all of them.  Returns, per weight,
    (sigma, u_saved, v_saved) - the post-iteration vectors this forward's backward needs."""
    if not weights:
        return []
    dev = weights[0][0].device
    _lib.ensure_device(dev)
    n = len(weights)
    sizes = [(w.shape[0], w.numel() // w.shape[0]) for w, _, _ in weights]
    total = sum(1 + r + c for r, c in sizes)
    ws_floats = sum(r + c for r, c in sizes)
    buf = torch.empty(total + ws_floats + 16, dtype=torch.float32, device=dev)
    table = (_lib.VgSnItem * n)()
    out, off = [], 0
    for i, ((w, u, v), (r, c)) in enumerate(zip(weights, sizes)):
        sigma = buf[off:off + 1]
        u_s = buf[off + 1:off + 1 + r]
        v_s = buf[off + 1 + r:off + 1 + r + c]
        off += 1 + r + c
        wd = w.detach()
        assert wd.is_contiguous()
        table[i] = _lib.VgSnItem(ptr(wd), ptr(u), ptr(v), ptr(u_s), ptr(v_s), ptr(sigma), r, c)
        out.append((sigma, u_s, v_s))
    work = buf[off:]
    call("vg_spectral_norm_sigma_batched", table, n, int(training), SN_EPS, ptr(work), ws_floats, stream_ptr())
    return out


# ----------------------------------------------------------------------------------------------
# BatchNorm + LeakyReLU + Dropout
# ----------------------------------------------------------------------------------------------
def _bn_desc(x, slope=1.0, drop_p=0.0, offset=0, training=True):
    n, c, h, w = x.shape
    step_t = rng.step_tensor(x.device) if drop_p > 0 else None
    d = VgBnDesc(n * h * w, c, h * w, vg_dtype(x.dtype), float(slope), float(drop_p), rng.seed, int(offset),
                 int(config.sample_offset), int(training), ptr(step_t))
    return d


def _bn_channel(x, gamma, beta, running_mean, running_var, sums, training, mr_out):
    """VgBnChannel for the fused kernels: batch statistics (`sums`, already all-reduced under SyncBN) in training,
    running statistics in eval.  The returned struct holds raw pointers - the caller keeps the tensors alive."""
    n, c, h, w = x.shape
    use_batch = training or running_mean is None
    count = float(n * h * w * (_world() if training else 1))
    return _lib.VgBnChannel(ptr(gamma), ptr(beta), None, ptr(sums) if use_batch else None, count,
                            ptr(running_mean), ptr(running_var), ptr(mr_out), BN_EPS, BN_MOMENTUM)


def _bn_prepare_sums(x, sums, training, running_mean):
    """Training: make sure the per-channel sums of x exist (the producer may have accumulated them already) and are
    global (SyncBN all-reduce).  Eval: nothing."""
    if not (training or running_mean is None):
        return None
    if sums is None:
        sums = zeros_f64(2 * x.shape[1], x.device)
        d = _bn_desc(x)
        call("vg_bn_stats", ptr(x), C.byref(d), ptr(sums), stream_ptr())
    if training:
        _allreduce_sums(sums)
    return sums


def bn_batch_stats(x, sums, running_mean, running_var, training, momentum=BN_MOMENTUM, eps=BN_EPS):
    """Per-channel (mean, rstd) of `x` (float[2C]) as a stand-alone step - batch statistics (+ SyncBN all-reduce, +
    running-stat update) in training, running statistics in eval.  The hot path folds this into the consuming
    kernel (vg_bn_act_forward_fused / vg_bn_add_forward_fused); kept for callers that need the numbers alone."""
    n, c, h, w = x.shape
    s = stream_ptr()
    mr = torch.empty(2 * c, dtype=torch.float32, device=x.device)
    if training:
        sums = _bn_prepare_sums(x, sums, True, running_mean)
        count = float(n * h * w * _world())
        call("vg_bn_finalize", ptr(sums), count, c, eps, momentum, ptr(running_mean), ptr(running_var), ptr(mr), s)
    else:
        call("vg_bn_eval_stats", ptr(running_mean), ptr(running_var), c, eps, ptr(mr), s)
    return mr


def _bn_act_forward(x, g, b, running_mean, running_var, sums, d, training):
    """finalize (or eval statistics) + BatchNorm + LeakyReLU + dropout in ONE launch; returns (y, mean_rstd)."""
    c = x.shape[1]
    sums = _bn_prepare_sums(x, sums, training, running_mean)
    mr = torch.empty(2 * c, dtype=torch.float32, device=x.device)
    ch = _bn_channel(x, g, b, running_mean, running_var, sums, training, mr)
    y = torch.empty_like(x)
    call("vg_bn_act_forward_fused", ptr(x), C.byref(ch), C.byref(d), ptr(y), stream_ptr())
    return y, mr


def _bn_backward(dy, x, mr, gamma, beta, d, out_colscale=None, addend=None, need_dx=True, need_params=True,
                 gbuf=None, bbuf=None):
    """Shared BN(+act+dropout) backward: returns dx, dgamma, dbeta (None when accumulated into the
    trainer's flat gradient views gbuf / bbuf).  Two launches: the reduction, then the apply - which also adds the
    parameter gradients (block 0) and the shortcut gradient `addend` of a residual fork."""
    s = stream_ptr()
    c = d.c
    sums = zeros_f64(2 * c, x.device)
    call("vg_bn_act_backward_reduce", ptr(dy), ptr(x), ptr(mr), ptr(gamma), ptr(beta), C.byref(d), ptr(sums), s)
    if d.training:
        _allreduce_sums(sums)          # SyncBN: global sums enter dx; param grads get reduced again
    fused = gbuf is not None and bbuf is not None
    dgamma = dbeta = None
    if need_params:
        dgamma = gbuf if fused else zeros_f32((c,), x.device)
        dbeta = bbuf if fused else zeros_f32((c,), x.device)
    # the parameter gradients are all-reduced (summed) later with the flat gradient buffer, so each rank
    # contributes 1/world of the already-global sums
    pscale = 1.0 / _world() if (d.training and _world() > 1) else 1.0
    dx = None
    if need_dx:
        dx = torch.empty_like(x)
        count = float(d.rows * _world())
        call("vg_bn_act_backward_apply_fused", ptr(dy), ptr(x), ptr(mr), ptr(gamma), ptr(beta), ptr(sums), count,
             C.byref(d), ptr(out_colscale), ptr(addend), ptr(dx), ptr(dgamma), ptr(dbeta), pscale, s)
    elif need_params:
        call("vg_bn_param_grads_scaled", ptr(sums), c, pscale, ptr(dgamma), ptr(dbeta), s)
    if fused or not need_params:
        return dx, None, None
    return dx, dgamma, dbeta


class BnActFn(Function):
    """y = dropout(leaky_relu(batch_norm(x)))  - README.md:188-190, 192-193, 410-411, 414-415,
    467-468.  `sums`: optional pre-accumulated double[2C] statistics from the producer."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, sums, slope, drop_p, offset, training,
                out_colscale):
        _lib.ensure_device(x.device)
        assert is_act(x)
        g, b = gamma.detach(), beta.detach()
        d = _bn_desc(x, slope, drop_p if training else 0.0, offset, training)
        y, mr = _bn_act_forward(x, g, b, running_mean, running_var, sums, d, training)
        ctx.d = d
        ctx.noted = note_use(ctx, ((1, gamma), (2, beta)))
        ctx.save_for_backward(x, mr, g, b, out_colscale, gamma, beta)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mr, g, b, ocs, gamma, beta = ctx.saved_tensors
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            assert ctx.d.drop_p == 0.0, "second-order path: elementwise dropout is not on the discriminator path"
            if not ctx.needs_input_grad[0]:
                return (None,) * 11
            dx = gp.BnBwdFn.apply(dy, x, gamma, beta, mr, ctx.d, ocs)
            return (dx,) + (None,) * 10
        dy = as_act(dy, x.dtype)
        dx, dgamma, dbeta = _bn_backward(dy, x, mr, g, b, ctx.d, out_colscale=ocs,
                                         need_dx=ctx.needs_input_grad[0], need_params=ctx.needs_input_grad[1],
                                         gbuf=live_grad_buf(gamma), bbuf=live_grad_buf(beta))
        note_done(ctx.noted)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None, None


class BnActForkFn(Function):
    """(y, x_pass) = (dropout(leaky_relu(batch_norm(x))), x): the fork at the top of a pre-activation residual
    block, where x feeds both the main path and the shortcut (README.md:188-195, 410-417).  Autograd hands the
    backward BOTH gradients at once, and the shortcut's gradient is added inside the BatchNorm backward-apply
    kernel (`addend`) instead of by a separate full-tensor accumulation kernel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, sums, slope, drop_p, offset, training):
        _lib.ensure_device(x.device)
        assert is_act(x)
        g, b = gamma.detach(), beta.detach()
        d = _bn_desc(x, slope, drop_p if training else 0.0, offset, training)
        y, mr = _bn_act_forward(x, g, b, running_mean, running_var, sums, d, training)
        ctx.d = d
        ctx.noted = note_use(ctx, ((1, gamma), (2, beta)))
        ctx.save_for_backward(x, mr, g, b, gamma, beta)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dpass):
        x, mr, g, b, gamma, beta = ctx.saved_tensors
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            assert ctx.d.drop_p == 0.0, "second-order path: elementwise dropout is not on the discriminator path"
            if not ctx.needs_input_grad[0]:
                return (None,) * 10
            dx = gp.BnBwdFn.apply(dy, x, gamma, beta, mr, ctx.d, None) if dy is not None else None
            if dpass is not None:
                dx = dpass.to(x.dtype) if dx is None else dx + dpass.to(dx.dtype)
            return (dx,) + (None,) * 9
        if dy is None:                       # only the pass-through branch was used
            note_done(ctx.noted)
            return (as_act(dpass, x.dtype) if dpass is not None else None,) + (None,) * 9
        dy = as_act(dy, x.dtype)
        addend = as_act(dpass, x.dtype) if dpass is not None else None
        dx, dgamma, dbeta = _bn_backward(dy, x, mr, g, b, ctx.d, addend=addend, need_dx=ctx.needs_input_grad[0],
                                         need_params=ctx.needs_input_grad[1], gbuf=live_grad_buf(gamma),
                                         bbuf=live_grad_buf(beta))
        note_done(ctx.noted)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None


def bn_act_fork(x, bn, *, slope=1.0, drop_p=0.0, training=True, sums=None, tag=""):
    """bn_act that also returns x for the block's shortcut; see BnActForkFn."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        # nothing flows back into x: keep the plain tensor so the shortcut does not compute a useless input gradient
        return bn_act(x, bn, slope=slope, drop_p=drop_p, training=training, sums=sums, tag=tag), x
    offset = rng.next_site(tag, tuple(x.shape)) if (training and drop_p > 0) else 0
    if training and bn.track_running_stats and not config.defer_num_batches_tracked:
        bn.num_batches_tracked += 1
    return BnActForkFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, sums, slope, drop_p, offset, training)


def bn_act(x, bn, *, slope=1.0, drop_p=0.0, training=True, sums=None, out_colscale=None, tag=""):
    """`bn` is an nn.BatchNorm2d used as a parameter/buffer container."""
    offset = rng.next_site(tag, tuple(x.shape)) if (training and drop_p > 0) else 0
    if training and bn.track_running_stats and not config.defer_num_batches_tracked:
        bn.num_batches_tracked += 1
    return BnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, sums, slope, drop_p, offset,
                         training, out_colscale)


class BnAddFn(Function):
    """out = leaky_relu(bnA(a) + bnB(b)), either BN optional: the residual add of
    README.md:183-184, 195, 405-406, 417 fused with the shortcut's BatchNorm."""

    @staticmethod
    def forward(ctx, a, b, ga, ba, rma, rva, sums_a, gb, bb, rmb, rvb, sums_b, slope, training, stats_out):
        _lib.ensure_device(a.device)
        assert is_act(a) and is_act(b) and a.shape == b.shape and a.dtype == b.dtype
        mra = mrb = None
        cha = chb = None
        ga_p, ba_p, gb_p, bb_p = ga, ba, gb, bb
        c = a.shape[1]
        if ga is not None:
            ga, ba = ga.detach(), ba.detach()
            sums_a = _bn_prepare_sums(a, sums_a, training, rma)
            mra = torch.empty(2 * c, dtype=torch.float32, device=a.device)
            cha = _bn_channel(a, ga, ba, rma, rva, sums_a, training, mra)
        if gb is not None:
            gb, bb = gb.detach(), bb.detach()
            sums_b = _bn_prepare_sums(b, sums_b, training, rmb)
            mrb = torch.empty(2 * c, dtype=torch.float32, device=a.device)
            chb = _bn_channel(b, gb, bb, rmb, rvb, sums_b, training, mrb)
        d = _bn_desc(a, slope, 0.0, 0, training)
        out = torch.empty_like(a)
        # both finalizes (running-stat updates included) + affine + add + LeakyReLU + next block's statistics: one launch
        call("vg_bn_add_forward_fused", ptr(a), C.byref(cha) if cha is not None else None, ptr(b),
             C.byref(chb) if chb is not None else None, C.byref(d), ptr(out), ptr(stats_out), stream_ptr())
        ctx.d = d
        ctx.slope = slope
        ctx.noted = note_use(ctx, ((2, ga_p), (3, ba_p), (7, gb_p), (8, bb_p)))
        ctx.save_for_backward(a if mra is not None else None, mra, ga, ba, b if mrb is not None else None, mrb, gb, bb,
                              out if slope != 1.0 else None, ga_p, ba_p, gb_p, bb_p)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, mra, ga, ba, b, mrb, gb, bb, out, ga_p, ba_p, gb_p, bb_p = ctx.saved_tensors
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            dpre = dout * gp.lrelu_mask(out, ctx.slope).to(dout.dtype) if ctx.slope != 1.0 else dout
            d1 = VgBnDesc.from_buffer_copy(ctx.d)
            d1.slope = 1.0
            da = db = None
            if ctx.needs_input_grad[0]:
                da = gp.BnBwdFn.apply(dpre, a, ga_p, ba_p, mra, d1, None) if mra is not None else dpre
            if ctx.needs_input_grad[1]:
                db = gp.BnBwdFn.apply(dpre, b, gb_p, bb_p, mrb, d1, None) if mrb is not None else dpre
            return (da, db) + (None,) * 13
        ref = a if a is not None else (b if b is not None else out)
        dtype = ref.dtype if ref is not None else dout.dtype
        dout = as_act(dout, dtype)
        s = stream_ptr()
        if ctx.slope != 1.0:
            dpre = torch.empty_like(dout)
            call("vg_lrelu_backward", ptr(dout), ptr(out), dout.numel(), vg_dtype(dout.dtype), float(ctx.slope),
                 ptr(dpre), s)
        else:
            dpre = dout
        d = VgBnDesc.from_buffer_copy(ctx.d)
        d.slope = 1.0
        da = db = dga = dba = dgb = dbb = None
        if mra is not None:
            da, dga, dba = _bn_backward(dpre, a, mra, ga, ba, d, need_dx=ctx.needs_input_grad[0],
                                        need_params=ctx.needs_input_grad[2], gbuf=live_grad_buf(ga_p),
                                        bbuf=live_grad_buf(ba_p))
        elif ctx.needs_input_grad[0]:
            da = dpre
        if mrb is not None:
            db, dgb, dbb = _bn_backward(dpre, b, mrb, gb, bb, d, need_dx=ctx.needs_input_grad[1],
                                        need_params=ctx.needs_input_grad[7], gbuf=live_grad_buf(gb_p),
                                        bbuf=live_grad_buf(bb_p))
        elif ctx.needs_input_grad[1]:
            db = dpre
        note_done(ctx.noted)
        return da, db, dga, dba, None, None, None, dgb, dbb, None, None, None, None, None, None


def bn_add(a, b, bn_a=None, bn_b=None, *, slope=1.0, training=True, sums_a=None, sums_b=None, stats_out=None):
    def unpack(bn):
        if bn is None:
            return None, None, None, None
        if training and bn.track_running_stats and not config.defer_num_batches_tracked:
            bn.num_batches_tracked += 1
        return bn.weight, bn.bias, bn.running_mean, bn.running_var

    ga, ba, rma, rva = unpack(bn_a)
    gb, bb, rmb, rvb = unpack(bn_b)
    return BnAddFn.apply(a, b, ga, ba, rma, rva, sums_a, gb, bb, rmb, rvb, sums_b, slope, training, stats_out)


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        call("vg_add", ptr(a), ptr(b), a.numel(), vg_dtype(a.dtype), ptr(out), stream_ptr())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


# ----------------------------------------------------------------------------------------------
# randomness helpers
# ----------------------------------------------------------------------------------------------
def dropout2d_scale(n, c, p, device, tag="dropout2d"):
    """Per-(n,c) Dropout2d scale (0 or 1/(1-p)) from Philox - nn.Dropout2d, README.md:381."""
    _lib.ensure_device(device)
    offset = rng.next_site(tag, (n, c))
    out = torch.empty((n, c), dtype=torch.float32, device=device)
    call("vg_dropout2d_scale", ptr(out), n, c, float(p), rng.seed, offset, ptr(rng.step_tensor(device)),
         int(config.sample_offset), stream_ptr())
    return out


def philox_normal(shape, device, tag="randn"):
    """Standard-normal noise in NHWC element order of a (N,C,H,W) activation (fp32)."""
    _lib.ensure_device(device)
    n, c, h, w = shape
    offset = rng.next_site(tag, tuple(shape))
    out = empty_act(n, c, h, w, torch.float32, device)
    start = int(config.sample_offset) * c * h * w
    call("vg_philox_normal", ptr(out), out.numel(), rng.seed, offset, ptr(rng.step_tensor(device)), start,
         stream_ptr())
    return out


def philox_uniform(n, device, tag="uniform"):
    """`n` uniform [0,1) fp32 numbers indexed by the GLOBAL sample (partition-invariant)."""
    _lib.ensure_device(device)
    offset = rng.next_site(tag, (n,))
    out = torch.empty(n, dtype=torch.float32, device=device)
    call("vg_philox_uniform", ptr(out), n, rng.seed, offset, ptr(rng.step_tensor(device)), int(config.sample_offset),
         stream_ptr())
    return out


def export_dropout_mask(shape, drop_p, offset, device, step=None, sample_offset=0, seed=None):
    """Keep-mask bytes (logical NCHW view) for a BnActFn dropout site - used by the tests to feed
    the oracle the very same mask."""
    _lib.ensure_device(device)
    n, c, h, w = shape
    st = None
    if step is not None:
        st = torch.full((1,), int(step), dtype=torch.int64, device=device)
    d = VgBnDesc(n * h * w, c, h * w, _lib.VG_F32, 1.0, float(drop_p), seed if seed is not None else rng.seed,
                 int(offset), int(sample_offset), 1, ptr(st))
    m = torch.empty((n, h, w, c), dtype=torch.uint8, device=device)
    call("vg_dropout_mask", C.byref(d), ptr(m), stream_ptr())
    return m.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# discriminator head
# ----------------------------------------------------------------------------------------------
class AvgPoolFlattenFn(Function):
    """F.avg_pool2d(x, k) + view(B, -1) in NCHW order (README.md:471-473); output fp32."""

    @staticmethod
    def forward(ctx, x, k):
        _lib.ensure_device(x.device)
        assert is_act(x)
        n, c, h, w = x.shape
        out = torch.empty((n, c * (h // k) * (w // k)), dtype=torch.float32, device=x.device)
        call("vg_avgpool_flatten_forward", ptr(x), n, h, w, c, k, vg_dtype(x.dtype), ptr(out), stream_ptr())
        ctx.shape, ctx.k, ctx.dtype = (n, c, h, w), k, x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = ctx.shape
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            return gp.AvgPoolBwdFn.apply(g, ctx.shape, ctx.k, ctx.dtype), None
        g = g.contiguous().float()
        dx = empty_act(n, c, h, w, ctx.dtype, g.device)
        call("vg_avgpool_flatten_backward", ptr(g), n, h, w, c, ctx.k, vg_dtype(ctx.dtype), ptr(dx), stream_ptr())
        return dx, None


class LinearFn(Function):
    """leaky_relu(nn.Linear(x)) (README.md:474-483).  Activations fp32; the weight is streamed in
    the compute dtype (bf16 halves the 75 MB linear_1 read)."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope, wdtype):
        _lib.ensure_device(x.device)
        x = x.contiguous()
        m, k = x.shape
        n = weight.shape[0]
        s = stream_ptr()
        w = weight.detach()
        if wdtype != torch.float32:
            wq = torch.empty(w.shape, dtype=wdtype, device=w.device)
            call("vg_cast", ptr(w), _lib.VG_F32, ptr(wq), vg_dtype(wdtype), w.numel(), s)
            w = wq
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        b = bias.detach() if bias is not None else None
        call("vg_linear_forward", ptr(x), ptr(w), ptr(b), m, n, k, vg_dtype(wdtype), float(slope), ptr(y), s)
        ctx.dims, ctx.slope, ctx.wdtype, ctx.has_bias = (m, n, k), slope, wdtype, bias is not None
        ctx.weight_ref, ctx.bias_ref = weight, bias
        ctx.noted = note_use(ctx, ((1, weight), (2, bias)))
        ctx.save_for_backward(x, w, y if slope != 1.0 else None, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y, weight = ctx.saved_tensors
        m, n, k = ctx.dims
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            if not ctx.needs_input_grad[0]:          # input gradient only, see ConvFn.backward
                return (None,) * 5
            dpre = dy * gp.lrelu_mask(y, ctx.slope) if ctx.slope != 1.0 else dy
            return gp.LinearDgradFn.apply(dpre, weight, ctx.wdtype), None, None, None, None
        s = stream_ptr()
        dy = dy.contiguous().float()
        if ctx.slope != 1.0:
            dpre = torch.empty_like(dy)
            call("vg_lrelu_backward", ptr(dy), ptr(y), dy.numel(), _lib.VG_F32, float(ctx.slope), ptr(dpre), s)
            dy = dpre
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((m, k), dtype=torch.float32, device=dy.device)
            call("vg_linear_dgrad", ptr(dy), ptr(w), m, n, k, vg_dtype(ctx.wdtype), ptr(dx), s)
        if ctx.needs_input_grad[1]:
            wbuf, bbuf = live_grad_buf(ctx.weight_ref), live_grad_buf(ctx.bias_ref)
            fused = wbuf is not None and (not ctx.has_bias or bbuf is not None)
            dw = wbuf if fused else zeros_f32((n, k), dy.device)
            db = (bbuf if fused else zeros_f32((n,), dy.device)) if ctx.has_bias else None
            call("vg_linear_wgrad", ptr(x), ptr(dy), m, n, k, vg_dtype(ctx.wdtype), ptr(dw), ptr(db), s)
            if fused:
                dw = db = None
        note_done(ctx.noted)
        return dx, dw, db, None, None


class LeakyReluFn(Function):
    """nn.LeakyReLU on the discriminator head's fp32 activations (README.md:475-481)."""

    @staticmethod
    def forward(ctx, x, slope):
        x = x.contiguous() if not (x.is_contiguous() or is_act(x)) else x
        y = torch.empty_like(x)
        call("vg_lrelu_forward", ptr(x), x.numel(), _lib.VG_F32, float(slope), ptr(y), stream_ptr())
        ctx.slope = slope
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        if torch.is_grad_enabled():          # create_graph=True (gradient penalty)
            from . import gp
            return dy * gp.lrelu_mask(y, ctx.slope).to(dy.dtype), None
        if dy.stride() != y.stride():
            dy = dy.contiguous() if y.is_contiguous() else as_act(dy, torch.float32)
        dx = torch.empty_like(y)
        call("vg_lrelu_backward", ptr(dy), ptr(y), y.numel(), _lib.VG_F32, float(ctx.slope), ptr(dx), stream_ptr())
        return dx, None


def linear(x, weight, bias, slope, wdtype):
    """leaky_relu(nn.Linear(x)).  On the bf16 path, layers whose sizes fit the tensor-core tiles run
    as 1x1 convolutions on [B, C, 1, 1] activations through the tcgen05 kernels (forward with
    split-K, dgrad, wgrad); everything else takes the CUDA-core GEMM."""
    m, k = x.shape
    n = weight.shape[0]
    if wdtype == torch.bfloat16 and k % 64 == 0 and n % 64 == 0:
        xa = to_act(x.view(m, k, 1, 1), torch.bfloat16)
        wv = weight.view(n, k, 1, 1)
        gb = getattr(weight, "_vg_grad_buf", None)
        if gb is not None:
            wv._vg_grad_buf = gb.view(n, k, 1, 1)      # wgrad accumulates straight into the flat buffer
            wv._vg_owner = weight                      # ... while weight.grad still IS that buffer (live_grad_buf)
        wv._vg_packs = getattr(weight, "_vg_packs", None)
        y = conv(xa, wv, bias, geom=ConvGeom(1, 1, 0, False), out_dtype=torch.float32)
        if slope != 1.0:
            y = LeakyReluFn.apply(y, slope)
        return y.view(m, n)
    return LinearFn.apply(x, weight, bias, slope, wdtype)


# ----------------------------------------------------------------------------------------------
# reparameterisation and losses
# ----------------------------------------------------------------------------------------------
class ReparamFn(Function):
    """README.md:575-582: log_var = clamp(raw, -50, 50); z = mu + exp(0.5 log_var) * eps."""

    @staticmethod
    def forward(ctx, mu, lv_raw, eps, training, z_dtype):
        _lib.ensure_device(mu.device)
        assert is_act(mu) and is_act(lv_raw) and mu.dtype == torch.float32 and lv_raw.dtype == torch.float32
        n, c, h, w = mu.shape
        z = empty_act(n, c, h, w, z_dtype, mu.device)
        lv = torch.empty_like(lv_raw)
        call("vg_reparam_forward", ptr(mu), ptr(lv_raw), ptr(eps), mu.numel(), int(training), vg_dtype(z_dtype),
             ptr(z), ptr(lv), stream_ptr())
        ctx.training, ctx.z_dtype = training, z_dtype
        ctx.save_for_backward(lv_raw, eps)
        return z, lv

    @staticmethod
    @once_differentiable
    def backward(ctx, dz, dlv):
        lv_raw, eps = ctx.saved_tensors
        d_mu = torch.empty_like(lv_raw)
        d_lv = torch.empty_like(lv_raw)
        if dz is None:
            dz = zeros_f32(lv_raw.shape, lv_raw.device).permute(0, 1, 2, 3)
            dz = as_act(dz, ctx.z_dtype)
        else:
            dz = as_act(dz, ctx.z_dtype)
        if dlv is not None:
            dlv = as_act(dlv, torch.float32)
        call("vg_reparam_backward", ptr(dz), ptr(lv_raw), ptr(eps), ptr(dlv), lv_raw.numel(), int(ctx.training),
             vg_dtype(ctx.z_dtype), ptr(d_mu), ptr(d_lv), stream_ptr())
        return d_mu, d_lv, None, None, None


def _scaled(t, scale):
    """t * scale with a DEVICE scalar `scale` (no host sync, CUDA-graph safe) on this library's kernels."""
    out = torch.empty_like(t)
    call("vg_scale", ptr(t), ptr(scale.reshape(1).contiguous()), t.numel(), vg_dtype(t.dtype), ptr(out), stream_ptr())
    return out


class GeneratorLossFn(Function):
    """One fused kernel for the generator objective AND its gradients (README.md:816-831):
    adv (BCE-with-logits vs 1, or -mean D) + w_recon*(L1+MSE) + w_kl*KL.  Returns
    (total, recon, kl, adv) as fp32 scalars; backward only scales the stored gradients."""

    @staticmethod
    def forward(ctx, xhat, x, mu, lv, logits, adv_mode, w_adv, w_recon, w_kl):
        _lib.ensure_device(xhat.device)
        dev = xhat.device
        assert is_act(xhat) and is_act(mu) and is_act(lv)
        xf = x if (x.dtype == torch.float32 and is_act(x)) else as_act(x, torch.float32)
        world = _world()
        nlog = logits.numel() if logits is not None else 0
        d = VgLossDesc(xhat.numel(), xhat.numel() * world, mu.numel(), nlog, nlog * world, int(adv_mode),
                       float(w_adv), float(w_recon), float(w_kl), vg_dtype(xhat.dtype))
        d_xhat = torch.empty_like(xhat)
        d_mu = torch.empty_like(mu)
        d_lv = torch.empty_like(lv)
        d_log = torch.empty_like(logits) if logits is not None else None
        losses = zeros_f64(4, dev)
        lg = logits.detach().contiguous() if logits is not None else None
        call("vg_generator_loss", ptr(xhat), ptr(xf), ptr(mu), ptr(lv), ptr(lg), C.byref(d), ptr(d_xhat), ptr(d_mu),
             ptr(d_lv), ptr(d_log), ptr(losses), stream_ptr())
        ctx.save_for_backward(d_xhat, d_mu, d_lv, d_log)
        ctx.unit_grad = config.trainer_active
        out = losses.to(torch.float32)
        total, recon, kl, adv = out[0], out[1], out[2], out[3]
        # only `total` carries gradients; backprop from the reported parts raises instead of silently dropping them
        ctx.mark_non_differentiable(recon, kl, adv)
        return total, recon, kl, adv

    @staticmethod
    @once_differentiable
    def backward(ctx, g_total, g_recon, g_kl, g_adv):
        d_xhat, d_mu, d_lv, d_log = ctx.saved_tensors
        if ctx.unit_grad:
            # VaeGanTrainer calls total.backward() (incoming gradient exactly 1): no rescale kernels in the hot loop
            return d_xhat, None, d_mu, d_lv, d_log, None, None, None, None
        # general use (loss / k, GradScaler, 0.5 * loss, ...): the stored gradients scale with the incoming one
        sc = g_total.to(torch.float32)
        return (_scaled(d_xhat, sc), None, _scaled(d_mu, sc), _scaled(d_lv, sc),
                _scaled(d_log, sc) if d_log is not None else None, None, None, None, None)


class DiscriminatorLossFn(Function):
    """README.md:792-793 (critic) or BCE-with-logits (north_star): returns (total, real, fake)."""

    @staticmethod
    def forward(ctx, d_real, d_fake, adv_mode):
        _lib.ensure_device(d_real.device)
        n = d_real.numel()
        g_real = torch.empty_like(d_real)
        g_fake = torch.empty_like(d_fake)
        losses = zeros_f64(3, d_real.device)
        call("vg_discriminator_loss", ptr(d_real.detach().contiguous()), ptr(d_fake.detach().contiguous()), n,
             n * _world(), int(adv_mode), ptr(g_real), ptr(g_fake), ptr(losses), stream_ptr())
        ctx.save_for_backward(g_real, g_fake)
        ctx.unit_grad = config.trainer_active
        out = losses.to(torch.float32)
        total, lr, lf = out[0], out[1], out[2]
        ctx.mark_non_differentiable(lr, lf)
        return total, lr, lf

    @staticmethod
    @once_differentiable
    def backward(ctx, g_total, g_r, g_f):
        g_real, g_fake = ctx.saved_tensors
        if ctx.unit_grad:
            return g_real, g_fake, None
        sc = g_total.to(torch.float32)
        return _scaled(g_real, sc), _scaled(g_fake, sc), None


# ----------------------------------------------------------------------------------------------
# fused optimizer over flat buffers
# ----------------------------------------------------------------------------------------------
def optimizer_step(p, g, m, v, *, kind="adam", lr=3e-4, betas=(0.9, 0.999), alpha=0.99, eps=1e-8, weight_decay=0.0,
                   step=1, clamp=0.0, grad_scale=1.0, step_tensor: Optional[torch.Tensor] = None):
    _lib.ensure_device(p.device)
    k = 0 if kind == "adam" else 1
    d = VgOptDesc(k, lr, betas[0], betas[1], alpha, eps, weight_decay, 1 - betas[0] ** max(step, 1),
                  1 - betas[1] ** max(step, 1), clamp, grad_scale)
    call("vg_optimizer_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), C.byref(d), ptr(step_tensor), stream_ptr())
