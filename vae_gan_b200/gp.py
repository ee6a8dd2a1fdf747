"""Second-order (create_graph=True) backward of the discriminator path: what `compute_gradient_penalty`
(README.md:717-739) needs.

The reference calls `autograd.grad(D(interpolates), interpolates, create_graph=True)` and then
backpropagates the penalty through that gradient.  Here every first-order backward of the functions
in `functional.py` has a *differentiable* twin below, used only when autograd runs the backward with
grad mode enabled; the twins are again `autograd.Function`s whose forward is the same CUDA kernel the
ordinary backward launches and whose backward (the double backward) is built from the same kernels:

    conv dgrad      dx = dgrad(dy, W)            ->  d/d(dy) = conv_forward(G, W),  d/dW = wgrad(x := G, dy)
    BatchNorm bwd   dx = g*r*P(dy*m)             ->  closed form below (`bn_double_backward`)
    avg-pool bwd    linear                       ->  avg-pool forward
    Linear dgrad    dx = dy @ W                  ->  d/d(dy) = G @ W^T,             d/dW = dy^T @ G

LeakyReLU and dropout masks are piecewise constant, so they enter only as fixed multipliers.  Spectral
norm enters through `W / sigma(W)` with `sigma = u^T W v` (u, v constants - torch's hook semantics),
written with differentiable elementwise ops on the fp32 parameter.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import call, ptr, stream_ptr, vg_dtype


def _F():
    from . import functional as VF
    return VF


# ----------------------------------------------------------------------------------------------
# BatchNorm (+ LeakyReLU mask) backward and its derivative
# ----------------------------------------------------------------------------------------------
def bn_double_backward(G, dy, x, gamma, beta, mean, rstd, slope, count, training=True, allreduce=None):
    """Derivative of the BatchNorm(+LeakyReLU) input gradient.

    First backward (README.md:188-190 via torch's batch_norm backward), per channel with M = `count`:
        pre = gamma*xh + beta,  m = 1 if pre > 0 else slope,  dyb = dy*m,  xh = (x - mean)*rstd
        dx  = gamma*rstd*(dyb - mean(dyb) - xh*mean(dyb*xh))                         (training)
        dx  = gamma*rstd*dyb                                                         (eval)
    Given G = dL/d(dx) returns (dL/d(dy), dL/dx, dL/dgamma).  All tensors (N,C,H,W); statistics over
    (N,H,W).  `allreduce(t)` sums a double[5,C] tensor over data-parallel ranks (SyncBN).
    """
    c = x.shape[1]
    sh = (1, c, 1, 1)
    f = torch.float64 if x.dtype == torch.float64 else torch.float32
    xf, dyf, Gf = x.to(f), dy.to(f), G.to(f)
    mean, r, g, b = (t.to(f).view(sh) for t in (mean, rstd, gamma, beta))
    xh = (xf - mean) * r
    m = torch.where(g * xh + b > 0, torch.ones((), dtype=f, device=x.device), torch.full((), slope, dtype=f, device=x.device))
    dyb = dyf * m
    if not training:
        g_dy = g * r * Gf * m
        g_gamma = (Gf * r * dyb).sum((0, 2, 3))
        return g_dy.to(dy.dtype), None, g_gamma.float()
    dims = (0, 2, 3)
    sums = torch.stack([dyb.sum(dims, dtype=torch.float64), (dyb * xh).sum(dims, dtype=torch.float64),
                        Gf.sum(dims, dtype=torch.float64), (Gf * xh).sum(dims, dtype=torch.float64),
                        (Gf * dyb).sum(dims, dtype=torch.float64)])
    if allreduce is not None:
        allreduce(sums)
    M = float(count)
    s_dyb, s_dybx, s_G, s_Gx, s_Gdyb = (sums[i].to(f).view(sh) for i in range(5))
    a, bb, cc, mG = s_dyb / M, s_dybx / M, s_Gx / M, s_G / M
    S1 = s_Gdyb - a * s_G - bb * s_Gx
    g_gamma = (r * S1).view(c)
    g_dy = g * r * (Gf - mG - xh * cc) * m
    Q = -g * r * (bb * Gf + cc * dyb)
    mQ = -g * r * (bb * mG + cc * a)
    mQx = -2.0 * g * r * bb * cc
    g_x = r * (Q - mQ - xh * mQx) - (r * r * g * S1 / M) * xh
    return g_dy.to(dy.dtype), g_x.to(x.dtype), g_gamma.float()


class BnBwdFn(Function):
    """dx of BatchNorm(+LeakyReLU) as a differentiable function of (dy, x, gamma); forward = the fused
    reduce + apply kernels of the ordinary backward.

    `ocs` is the Dropout2d column scale of the convolution that PRODUCED x (README.md:412): ConvFn stores the
    scaled output and expects the gradient with respect to the UNSCALED one, so both the returned dx and the
    adjoint handed back for x carry that factor."""

    @staticmethod
    def forward(ctx, dy, x, gamma, beta, mr, d, ocs):
        VF = _F()
        ctx.dy_dtype = dy.dtype
        dy = VF.as_act(dy, x.dtype)
        dx, _, _ = VF._bn_backward(dy, x, mr, gamma.detach(), beta.detach(), d, out_colscale=ocs, need_dx=True,
                                   need_params=False)
        ctx.d = d
        ctx.save_for_backward(dy, x, gamma, beta, mr, ocs)
        return dx

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        VF = _F()
        dy, x, gamma, beta, mr, ocs = ctx.saved_tensors
        d = ctx.d
        c = d.c
        world = VF._world() if d.training else 1
        count = float(d.rows * world)

        def allreduce(t):                      # SyncBN: the five per-channel sums are global
            flat = t.view(-1)
            for i in range(0, flat.numel(), 2048):      # the NVLink exchange moves <= 2048 doubles per call
                VF._allreduce_sums(flat[i:i + 2048])

        if d.training and c % 8 == 0 and c // 8 <= 256 and d.drop_p == 0.0:
            # fused kernels: one reduction pass (five per-channel sums) + one apply pass (both outputs)
            s = stream_ptr()
            G = VF.as_act(G, x.dtype)
            g, b = gamma.detach(), beta.detach()
            sums = VF.zeros_f64(5 * c, x.device)
            call("vg_bn_act_double_backward_reduce", ptr(dy), ptr(x), ptr(G), ptr(mr), ptr(g), ptr(b), C.byref(d), ptr(ocs), ptr(sums), s)
            if world > 1:
                allreduce(sums)
            g_dy, g_x = torch.empty_like(x), torch.empty_like(x)
            call("vg_bn_act_double_backward_apply", ptr(dy), ptr(x), ptr(G), ptr(mr), ptr(g), ptr(b), ptr(sums), count, C.byref(d),
                 ptr(ocs), ptr(g_dy), ptr(g_x), s)
            sv = sums.view(5, c)
            g_gamma = (mr[c:].double() * (sv[4] - sv[0] * sv[2] / count - sv[1] * sv[3] / count)).float()
        else:
            scale = ocs.view(x.shape[0], c, 1, 1) if ocs is not None else None
            if scale is not None:
                G = G.float() * scale
            g_dy, g_x, g_gamma = bn_double_backward(G, dy, x, gamma, beta, mr[:c], mr[c:], float(d.slope), count,
                                                    bool(d.training), allreduce if world > 1 else None)
            if scale is not None and g_x is not None:
                g_x = (g_x.float() * scale).to(x.dtype)
        if world > 1:
            g_gamma = g_gamma / world          # summed again with the flat gradient all-reduce
        return g_dy.to(ctx.dy_dtype), g_x, g_gamma, None, None, None, None


# ----------------------------------------------------------------------------------------------
# convolution input gradient
# ----------------------------------------------------------------------------------------------
class ConvDgradFn(Function):
    """dx = dgrad(dy, W) as a differentiable function of (dy, W); W is the EFFECTIVE fp32 weight in torch
    layout (already divided by sigma for spectral-normed layers)."""

    @staticmethod
    def forward(ctx, dy, w_eff, d, act_dtype):
        VF = _F()
        dev = dy.device
        ctx.dy_dtype = dy.dtype
        dy = VF.as_act(dy, act_dtype)
        s = stream_ptr()
        w = w_eff.detach().contiguous()
        pack_kn = torch.empty(w.numel(), dtype=act_dtype, device=dev)
        pack_nk = torch.empty(w.numel(), dtype=act_dtype, device=dev)
        call("vg_conv_pack_weights", C.byref(d), ptr(w), None, ptr(pack_kn), ptr(pack_nk), s)
        dx = VF.empty_act(d.n, d.c_in, d.h_in, d.w_in, act_dtype, dev)
        call("vg_conv_dgrad", C.byref(d), ptr(dy), ptr(pack_kn), ptr(pack_nk), ptr(dx), s)
        ctx.d, ctx.act_dtype, ctx.wshape = d, act_dtype, tuple(w_eff.shape)
        ctx.save_for_backward(dy, pack_kn, pack_nk)
        return dx

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        VF = _F()
        dy, pack_kn, pack_nk = ctx.saved_tensors
        d, dt = ctx.d, ctx.act_dtype
        dev = dy.device
        s = stream_ptr()
        G = VF.as_act(G, dt)
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            # <G, dgrad(dy, W)> = <conv(G, W), dy>
            g_dy = VF.empty_act(d.n, d.c_out, d.h_out, d.w_out, dt, dev)
            d2 = _lib.VgConvDesc.from_buffer_copy(d)
            d2.out_dtype = vg_dtype(dt)
            call("vg_conv_forward", C.byref(d2), ptr(G), ptr(pack_kn), ptr(pack_nk), None, None, ptr(g_dy), None, s)
            g_dy = g_dy.to(ctx.dy_dtype)
        if ctx.needs_input_grad[1]:
            g_w = VF.zeros_f32(ctx.wshape, dev)
            ws = torch.empty(g_w.numel(), dtype=torch.float32, device=dev) if dt == torch.bfloat16 else None
            call("vg_conv_wgrad", C.byref(d), ptr(G), ptr(dy), ptr(g_w), None, ptr(ws), s)
        return g_dy, g_w, None, None


def sn_effective_weight(weight, u, v):
    """W / sigma with sigma = u^T W v (u, v are constants): differentiable in W (README.md:378-387)."""
    w2 = weight.reshape(weight.shape[0], -1)
    sigma = (u.unsqueeze(1) * w2 * v.unsqueeze(0)).sum()
    return weight / sigma


# ----------------------------------------------------------------------------------------------
# discriminator head
# ----------------------------------------------------------------------------------------------
class AvgPoolBwdFn(Function):
    @staticmethod
    def forward(ctx, g, shape, k, dtype):
        VF = _F()
        n, c, h, w = shape
        g = g.contiguous().float()
        dx = VF.empty_act(n, c, h, w, dtype, g.device)
        call("vg_avgpool_flatten_backward", ptr(g), n, h, w, c, k, vg_dtype(dtype), ptr(dx), stream_ptr())
        ctx.shape, ctx.k, ctx.dtype = shape, k, dtype
        return dx

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        VF = _F()
        n, c, h, w = ctx.shape
        G = VF.as_act(G, ctx.dtype)
        out = torch.empty((n, c * (h // ctx.k) * (w // ctx.k)), dtype=torch.float32, device=G.device)
        call("vg_avgpool_flatten_forward", ptr(G), n, h, w, c, ctx.k, vg_dtype(ctx.dtype), ptr(out), stream_ptr())
        return out, None, None, None


class LinearDgradFn(Function):
    """dx = dy @ W (W: [n, k] fp32 parameter) on the CUDA-core GEMM kernels."""

    @staticmethod
    def forward(ctx, dy, weight, wdtype):
        VF = _F()
        dy = dy.contiguous().float()
        m, n = dy.shape
        k = weight.shape[1]
        s = stream_ptr()
        w = weight.detach()
        if wdtype != torch.float32:
            wq = torch.empty(w.shape, dtype=wdtype, device=w.device)
            call("vg_cast", ptr(w), _lib.VG_F32, ptr(wq), vg_dtype(wdtype), w.numel(), s)
            w = wq
        dx = torch.empty((m, k), dtype=torch.float32, device=dy.device)
        call("vg_linear_dgrad", ptr(dy), ptr(w), m, n, k, vg_dtype(wdtype), ptr(dx), s)
        ctx.dims, ctx.wdtype = (m, n, k), wdtype
        ctx.save_for_backward(dy, w)
        return dx

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        VF = _F()
        dy, w = ctx.saved_tensors
        m, n, k = ctx.dims
        s = stream_ptr()
        G = G.contiguous().float()
        g_dy = g_w = None
        if ctx.needs_input_grad[0]:
            g_dy = torch.empty((m, n), dtype=torch.float32, device=G.device)
            call("vg_linear_forward", ptr(G), ptr(w), None, m, n, k, vg_dtype(ctx.wdtype), 1.0, ptr(g_dy), s)
        if ctx.needs_input_grad[1]:
            g_w = VF.zeros_f32((n, k), G.device)
            call("vg_linear_wgrad", ptr(G), ptr(dy), m, n, k, vg_dtype(ctx.wdtype), ptr(g_w), None, s)
        return g_dy, g_w, None


def lrelu_mask(y, slope):
    """The constant multiplier of LeakyReLU's backward, from the OUTPUT sign (slope > 0)."""
    return torch.where(y > 0, torch.ones((), dtype=y.dtype, device=y.device), torch.full((), slope, dtype=y.dtype, device=y.device))


# ----------------------------------------------------------------------------------------------
# the penalty itself
# ----------------------------------------------------------------------------------------------
def gradient_penalty(discriminator, real, fake, alpha):
    """`compute_gradient_penalty` (README.md:717-739) with the interpolation weights given:
    `alpha` (B,1,1,1) replaces the reference's np.random draw."""
    inter = (alpha * real + (1 - alpha) * fake).detach().requires_grad_(True)
    d_inter = discriminator(inter)
    ones = torch.ones_like(d_inter)
    grads = torch.autograd.grad(outputs=d_inter, inputs=inter, grad_outputs=ones, create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
    grads = grads.reshape(grads.size(0), -1)
    return ((grads.norm(2, dim=1) - 1) ** 2).mean()
