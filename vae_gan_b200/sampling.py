"""Eval-mode / sampling path with the BatchNorms folded away.

What the reference runs here: `generator.eval(); generator.set_is_training(False)` followed by the forward
(`visualize_reconstructions`, README.md:1215-1226), `decode(z)` (README.md:661-664, BASELINE config 5) and `encode(x)`
(README.md:655-659).  In eval mode every BatchNorm is a fixed per-channel affine map, so for a pre-activation block
(README.md:188-195)

    a   = lrelu(bn1(x))              c1 = conv1(a)          b = lrelu(bn2(c1))
    out = conv2(b) + bnS(convS(x))

the two BatchNorms that FOLLOW a convolution fold into its weights (`vg_fold_bn_into_conv`: w * gamma * rstd, bias =
beta - mean * gamma * rstd), LeakyReLU and the residual add move into the convolution epilogue, and the NEXT block's
lrelu(bn1(.)) is emitted as a second output of conv2 (`VgConvEpilogue`).  A block is then three tensor-core launches
and no elementwise pass (training-mode kernels: 3 convolutions + 2 BatchNorm passes + the fused add).  Blocks the
tensor-core epilogue cannot take (a single-channel side: the first encoder block 1->64 and the reconstruction block
64->1, < 1 % of the FLOPs) run through the module's own eval-mode forward.

`FoldedGenerator` is an inference engine bound to a generator's CURRENT weights and running statistics: call `refold()`
after they change.  `graphed(fn, example)` captures any of its methods in a CUDA graph for a fixed batch size.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import functional as VF
from . import modules as M
from ._lib import call, ptr, stream_ptr, vg_dtype


class _FoldedBlock:
    pass


class FoldedGenerator:
    def __init__(self, generator: M.UnsupervisedGeneratorNetwork, dtype=torch.bfloat16):
        self.G = generator
        self.dtype = dtype
        self.device = next(generator.parameters()).device
        _lib.ensure_device(self.device)
        self.refold()

    # ------------------------------------------------------------------------------------------
    def _affine(self, bn):
        c = bn.num_features
        sc = torch.empty(c, dtype=torch.float32, device=self.device)
        sh = torch.empty(c, dtype=torch.float32, device=self.device)
        call("vg_bn_eval_affine", ptr(bn.weight.detach()), ptr(bn.bias.detach()), ptr(bn.running_mean), ptr(bn.running_var),
             VF.BN_EPS, c, ptr(sc), ptr(sh), stream_ptr())
        return sc, sh

    def _pack(self, w, geom, c_in, c_out):
        d, _, _ = VF._conv_desc((1, c_in, 8, 8), c_out, geom, self.dtype, self.dtype)
        kn = torch.empty(w.numel(), dtype=self.dtype, device=self.device)
        nk = torch.empty(w.numel(), dtype=self.dtype, device=self.device)
        call("vg_conv_pack_weights", C.byref(d), ptr(w), None, ptr(kn), ptr(nk), stream_ptr())
        return kn, nk

    def _fold_conv(self, conv, bn, geom, c_in, c_out):
        """packs of conv with the eval-mode BatchNorm `bn` that follows it folded in, and the resulting bias."""
        w = conv.weight.detach().contiguous()
        wf = torch.empty_like(w)
        bias = torch.empty(c_out, dtype=torch.float32, device=self.device)
        inner = w.shape[2] * w.shape[3]
        call("vg_fold_bn_into_conv", ptr(w), ptr(conv.bias.detach()) if conv.bias is not None else None, ptr(bn.weight.detach()),
             ptr(bn.bias.detach()), ptr(bn.running_mean), ptr(bn.running_var), VF.BN_EPS, c_out, c_in, inner, int(geom.transposed),
             ptr(wf), ptr(bias), stream_ptr())
        kn, nk = self._pack(wf, geom, c_in, c_out)
        return kn, nk, bias

    def _fold_block(self, blk: M.ResBlockVAE):
        fb = _FoldedBlock()
        fb.module = blk
        c_in, c_out = blk.bn1.num_features if blk.res_mode == "pre-activation" else None, blk.bn2.num_features
        fb.c_out = c_out
        fb.tc = (blk.res_mode == "pre-activation" and self.dtype == torch.bfloat16 and c_in % 64 == 0 and c_out % 64 == 0)
        if not fb.tc:
            fb.c_in = c_in
            return fb
        fb.c_in = c_in
        fb.g1 = M._MODE_GEOM[blk.mode]
        fb.g2 = M._MODE_GEOM["level"]
        fb.slope = float(blk.activation_fun.negative_slope)
        fb.pre_scale, fb.pre_shift = self._affine(blk.bn1)
        fb.k1, fb.n1, fb.b1 = self._fold_conv(blk.conv1, blk.bn2, fb.g1, c_in, c_out)
        fb.ks, fb.ns, fb.bs = self._fold_conv(blk.shortcut[0], blk.shortcut[1], fb.g1, c_in, c_out)
        fb.k2, fb.n2 = self._pack(blk.conv2.weight.detach().contiguous(), fb.g2, c_out, c_out)
        return fb

    def refold(self):
        """(Re)build the folded weights from the generator's current parameters and running statistics."""
        G = self.G
        self.enc = [self._fold_block(b) for b in G.encoder.encoder]
        self.dec = [self._fold_block(b) for b in G.decoder.decoder]
        cp = G.code_processor
        fd = cp.mu.weight.shape[1]
        self.mu_geom = M._MODE_GEOM["level"]
        self.mu_packs = self._pack(cp.mu.weight.detach().contiguous(), self.mu_geom, fd, cp.mu.weight.shape[0])
        self.mu_bias = cp.mu.bias.detach()
        self._ident = {}

    # ------------------------------------------------------------------------------------------
    def _conv(self, x, kn, nk, geom, c_out, *, bias=None, act_slope=1.0, residual=None, post=None, out_dtype=None):
        out_dtype = out_dtype or x.dtype
        d, ho, wo = VF._conv_desc(x.shape, c_out, geom, x.dtype, out_dtype)
        y = VF.empty_act(x.shape[0], c_out, ho, wo, out_dtype, x.device)
        y2 = torch.empty_like(y) if post is not None else None
        ep = _lib.VgConvEpilogue(ptr(bias), None, None, 0, float(act_slope), ptr(residual), ptr(y2),
                                 ptr(post[0]) if post is not None else None, ptr(post[1]) if post is not None else None,
                                 float(post[2]) if post is not None else 1.0)
        call("vg_conv_forward_fused", C.byref(d), ptr(x), ptr(kn), ptr(nk), C.byref(ep), ptr(y), None, stream_ptr())
        return y, y2

    def _pre_act(self, x, fb):
        """a = lrelu(bn1(x)) with the eval-mode affine (only needed where no producer emitted it as its second output)."""
        c = x.shape[1]
        if c not in self._ident:
            self._ident[c] = torch.cat([torch.zeros(c, device=self.device), torch.ones(c, device=self.device)])
        d = VF._bn_desc(x, fb.slope, 0.0, 0, False)
        a = torch.empty_like(x)
        call("vg_bn_act_forward", ptr(x), ptr(self._ident[c]), ptr(fb.pre_scale), ptr(fb.pre_shift), C.byref(d), ptr(a), stream_ptr())
        return a

    # Measured on B200 (profiles/r2_bench_decode_sweep.json): the folded path wins while the decode is launch-bound (B = 1:
    # 7.9 k vs 6.0 k img/s, B = 4: 28.4 k vs 21.2 k), ties at B = 16 and LOSES beyond (B = 256: 2.86 vs 2.33 ms) - the
    # 64-channel convolutions are epilogue-bound, so moving the residual read and a second store into their epilogue costs
    # more than the streaming kernels (80-88 % of HBM rate) it removes.  Above FOLDED_MAX_PIXELS output pixels the
    # methods therefore run the module's own eval-mode path; EPILOGUE_TAIL_MAX_PIXELS picks, inside the folded path,
    # between the dual-output conv2 epilogue and one dual-output streaming pass (vg_add_dual_forward).
    FOLDED_MAX_PIXELS = int(__import__("os").environ.get("VG_FOLDED_MAX_PIXELS", 12 * 96 * 96))
    EPILOGUE_TAIL_MAX_PIXELS = 32 * 96 * 96

    def _add_dual(self, c2, sc, post):
        d = VF._bn_desc(c2, 1.0, 0.0, 0, False)
        out, a = torch.empty_like(c2), torch.empty_like(c2)
        call("vg_add_dual_forward", ptr(c2), ptr(sc), ptr(post[0]), ptr(post[1]), float(post[2]), C.byref(d), ptr(out), ptr(a), stream_ptr())
        return out, a

    def _run_blocks(self, blocks, x):
        a = None
        for i, fb in enumerate(blocks):
            if not fb.tc:
                assert not fb.module.training, "FoldedGenerator runs eval-mode semantics: call generator.eval() first"
                with M._scope():
                    with M._scope():
                        x = fb.module(x)
                a = None
                continue
            if a is None:
                a = self._pre_act(x, fb)
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            post = (nxt.pre_scale, nxt.pre_shift, nxt.slope) if (nxt is not None and nxt.tc) else None
            b, _ = self._conv(a, fb.k1, fb.n1, fb.g1, fb.c_out, bias=fb.b1, act_slope=fb.slope)
            sc, _ = self._conv(x, fb.ks, fb.ns, fb.g1, fb.c_out, bias=fb.bs)
            if b.shape[0] * b.shape[2] * b.shape[3] <= self.EPILOGUE_TAIL_MAX_PIXELS:
                x, a = self._conv(b, fb.k2, fb.n2, fb.g2, fb.c_out, residual=sc, post=post)
            else:
                c2, _ = self._conv(b, fb.k2, fb.n2, fb.g2, fb.c_out)
                if post is not None:
                    x, a = self._add_dual(c2, sc, post)
                else:
                    x, a = VF.AddFn.apply(c2, sc), None
        return x

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """UnsupervisedGeneratorNetwork.decode (README.md:661-664) in eval mode: latent (B, C, h, w) -> image, fp32 NCHW."""
        with VF.compute_dtype(self.dtype):
            if z.shape[0] * z.shape[2] * z.shape[3] * 16 > self.FOLDED_MAX_PIXELS:      # x16: two upsampling blocks
                return self.G.decode(z)
            x = VF.to_act(z, self.dtype)
            y = self._run_blocks(self.dec, x)
            return VF.from_act(y)

    @torch.no_grad()
    def encode(self, img: torch.Tensor) -> torch.Tensor:
        """UnsupervisedGeneratorNetwork.encode (README.md:655-659) in eval mode: image -> latent mean mu, fp32 NCHW."""
        with VF.compute_dtype(self.dtype):
            if img.shape[0] * img.shape[2] * img.shape[3] > self.FOLDED_MAX_PIXELS:
                return self.G.encode(img)
            h = self._run_blocks(self.enc, VF.to_act(img, self.dtype))
            mu, _ = self._conv(h, self.mu_packs[0], self.mu_packs[1], self.mu_geom, self.mu_bias.shape[0], bias=self.mu_bias,
                               out_dtype=torch.float32)
            return VF.from_act(mu)

    @torch.no_grad()
    def reconstruct(self, img: torch.Tensor) -> torch.Tensor:
        """The eval-mode forward of visualize_reconstructions (README.md:1223-1226): z = mu, x_hat = decode(z)."""
        with VF.compute_dtype(self.dtype):
            if img.shape[0] * img.shape[2] * img.shape[3] > self.FOLDED_MAX_PIXELS:
                return self.G(img)[0]
            h = self._run_blocks(self.enc, VF.to_act(img, self.dtype))
            mu, _ = self._conv(h, self.mu_packs[0], self.mu_packs[1], self.mu_geom, self.mu_bias.shape[0], bias=self.mu_bias,
                               out_dtype=torch.float32)
            z = VF.to_act(mu, self.dtype)
            return VF.from_act(self._run_blocks(self.dec, z))

    def graphed(self, fn, example: torch.Tensor, warmup: int = 2):
        """Capture `fn(example)` (one of decode / encode / reconstruct) in a CUDA graph for this batch shape; returns
        (static_input, static_output, replay) - copy new data into static_input, call replay(), read static_output."""
        static_in = example.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = fn(static_in)
        return static_in, static_out, graph.replay
