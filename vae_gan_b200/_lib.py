"""ctypes binding of libvaegan_sm100.so (include/vaegan_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
# VG_LIB: an alternative build of the same library (kernel A/B experiments); the default is the in-tree build
LIB_PATH = Path(os.environ["VG_LIB"]) if os.environ.get("VG_LIB") else _PKG / "lib" / "libvaegan_sm100.so"

VG_F32, VG_BF16 = 0, 1

c_vp, c_int, c_ll, c_ull, c_f, c_d = C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_float, C.c_double


class VgConvDesc(C.Structure):
    _fields_ = [(n, c_int) for n in ("n", "h_in", "w_in", "c_in", "h_out", "w_out", "c_out", "kh", "kw",
                                     "stride", "pad", "transposed", "act_dtype", "out_dtype")]


class VgBnDesc(C.Structure):
    _fields_ = [("rows", c_ll), ("c", c_int), ("hw", c_int), ("dtype", c_int), ("slope", c_f),
                ("drop_p", c_f), ("seed", c_ull), ("offset", c_ull), ("sample_offset", c_ll),
                ("training", c_int), ("step_ptr", c_vp)]


class VgBnChannel(C.Structure):
    _fields_ = [("gamma", c_vp), ("beta", c_vp), ("mean_rstd_in", c_vp), ("sums", c_vp), ("count", c_d),
                ("running_mean", c_vp), ("running_var", c_vp), ("mean_rstd_out", c_vp), ("eps", c_f), ("momentum", c_f)]


class VgPackItem(C.Structure):
    _fields_ = [("w", c_vp), ("pack_kn", c_vp), ("pack_nk", c_vp), ("n0", c_int), ("n1", c_int), ("taps", c_int), ("transposed", c_int)]


class VgSnItem(C.Structure):
    _fields_ = [("w", c_vp), ("u", c_vp), ("v", c_vp), ("u_out", c_vp), ("v_out", c_vp), ("sigma", c_vp), ("rows", c_int), ("cols", c_int)]


class VgConvEpilogue(C.Structure):
    _fields_ = [("bias", c_vp), ("colscale", c_vp), ("sigma", c_vp), ("sigma_group_n", c_int), ("act_slope", c_f),
                ("residual", c_vp), ("y2", c_vp), ("post_scale", c_vp), ("post_shift", c_vp), ("post_slope", c_f)]


class VgLossDesc(C.Structure):
    _fields_ = [("n_pix", c_ll), ("n_pix_global", c_ll), ("n_lat", c_ll), ("n_logits", c_int),
                ("n_logits_global", c_int), ("adv_mode", c_int), ("w_adv", c_f), ("w_recon", c_f),
                ("w_kl", c_f), ("xhat_dtype", c_int)]


class VgPeerDesc(C.Structure):
    _fields_ = [("peer_data", c_vp * 8), ("peer_flags", c_vp * 8), ("rank", c_int), ("world", c_int), ("n_slots", c_int)]


class VgOptDesc(C.Structure):
    _fields_ = [("kind", c_int), ("lr", c_f), ("beta1", c_f), ("beta2", c_f), ("alpha", c_f), ("eps", c_f),
                ("weight_decay", c_f), ("bias_corr1", c_f), ("bias_corr2", c_f), ("clamp", c_f),
                ("grad_scale", c_f)]


_PROTOS = {
    "vg_version": (c_int, []),
    "vg_init": (c_int, [c_int]),
    "vg_last_error": (C.c_char_p, []),
    "vg_launch_count": (c_ull, []),
    "vg_set_force_simt": (c_int, [c_int]),
    "vg_set_deterministic": (c_int, [c_int, c_vp, C.c_size_t, c_vp, c_int]),
    "vg_get_deterministic": (c_int, []),
    "vg_conv_pack_weights": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_conv_wgrad_sn_supported": (c_int, [C.POINTER(VgConvDesc)]),
    "vg_conv_wgrad_sn": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_conv_tune_set": (c_int, [C.POINTER(VgConvDesc), c_int, c_int, c_int]),
    "vg_conv_tune_clear": (c_int, []),
    "vg_conv_tune_record": (c_int, [c_int]),
    "vg_conv_tune_seen": (c_int, [c_vp, c_int, C.POINTER(c_int)]),
    "vg_conv_pack_weights_batched": (c_int, [c_vp, c_int, c_int, c_vp]),
    "vg_conv_forward_scaled": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "vg_conv_forward_fused": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, C.POINTER(VgConvEpilogue), c_vp, c_vp, c_vp]),
    "vg_fold_bn_into_conv": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "vg_bn_eval_affine": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f, c_int, c_vp, c_vp, c_vp]),
    "vg_conv_dgrad_scaled": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "vg_spectral_norm_sigma_batched": (c_int, [c_vp, c_int, c_int, c_f, c_vp, C.c_size_t, c_vp]),
    "vg_conv_forward": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_conv_dgrad": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_conv_wgrad": (c_int, [C.POINTER(VgConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_bn_stats": (c_int, [c_vp, C.POINTER(VgBnDesc), c_vp, c_vp]),
    "vg_bn_finalize": (c_int, [c_vp, c_d, c_int, c_f, c_f, c_vp, c_vp, c_vp, c_vp]),
    "vg_bn_eval_stats": (c_int, [c_vp, c_vp, c_int, c_f, c_vp, c_vp]),
    "vg_bn_act_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, C.POINTER(VgBnDesc), c_vp, c_vp]),
    "vg_bn_act_backward_reduce": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(VgBnDesc), c_vp, c_vp]),
    "vg_bn_act_backward_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_d, C.POINTER(VgBnDesc), c_vp, c_vp, c_vp, c_vp]),
    "vg_bn_param_grads": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp]),
    "vg_bn_add_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(VgBnDesc), c_vp, c_vp, c_vp]),
    "vg_bn_act_forward_fused": (c_int, [c_vp, C.POINTER(VgBnChannel), C.POINTER(VgBnDesc), c_vp, c_vp]),
    "vg_bn_add_forward_fused": (c_int, [c_vp, C.POINTER(VgBnChannel), c_vp, C.POINTER(VgBnChannel), C.POINTER(VgBnDesc), c_vp, c_vp, c_vp]),
    "vg_bn_act_backward_apply_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_d, C.POINTER(VgBnDesc), c_vp, c_vp, c_vp,
                                               c_vp, c_vp, c_f, c_vp]),
    "vg_bn_param_grads_scaled": (c_int, [c_vp, c_int, c_f, c_vp, c_vp, c_vp]),
    "vg_add_dual_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f, C.POINTER(VgBnDesc), c_vp, c_vp, c_vp]),
    "vg_scale": (c_int, [c_vp, c_vp, c_ll, c_int, c_vp, c_vp]),
    "vg_normalize_images": (c_int, [c_vp, c_int, c_int, c_ll, c_vp, c_vp, c_vp]),
    "vg_lrelu_forward": (c_int, [c_vp, c_ll, c_int, c_f, c_vp, c_vp]),
    "vg_lrelu_backward": (c_int, [c_vp, c_vp, c_ll, c_int, c_f, c_vp, c_vp]),
    "vg_add": (c_int, [c_vp, c_vp, c_ll, c_int, c_vp, c_vp]),
    "vg_dropout_mask": (c_int, [C.POINTER(VgBnDesc), c_vp, c_vp]),
    "vg_dropout2d_scale": (c_int, [c_vp, c_int, c_int, c_f, c_ull, c_ull, c_vp, c_ll, c_vp]),
    "vg_philox_normal": (c_int, [c_vp, c_ll, c_ull, c_ull, c_vp, c_ll, c_vp]),
    "vg_bn_act_double_backward_reduce": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(VgBnDesc), c_vp, c_vp, c_vp]),
    "vg_bn_act_double_backward_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_d, C.POINTER(VgBnDesc), c_vp,
                                                c_vp, c_vp, c_vp]),
    "vg_philox_uniform": (c_int, [c_vp, c_ll, c_ull, c_ull, c_vp, c_ll, c_vp]),
    "vg_counter_add": (c_int, [c_vp, c_ull, c_vp]),
    "vg_avgpool_flatten_forward": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "vg_avgpool_flatten_backward": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "vg_linear_forward": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_vp, c_vp]),
    "vg_linear_dgrad": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "vg_linear_wgrad": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "vg_spectral_norm_sigma": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_f, c_vp, c_vp, c_vp]),
    "vg_spectral_norm_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "vg_reparam_forward": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp]),
    "vg_reparam_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_vp, c_vp, c_vp]),
    "vg_generator_loss": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(VgLossDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "vg_discriminator_loss": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "vg_optimizer_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, C.POINTER(VgOptDesc), c_vp, c_vp]),
    "vg_enable_peer_access": (c_int, [c_int]),
    "vg_peer_alloc": (c_int, [C.c_size_t, C.POINTER(c_vp)]),
    "vg_peer_free": (c_int, [c_vp]),
    "vg_peer_get_handle": (c_int, [c_vp, c_vp]),
    "vg_peer_open_handle": (c_int, [c_vp, C.POINTER(c_vp)]),
    "vg_peer_close_handle": (c_int, [c_vp]),
    "vg_peer_allreduce_f64": (c_int, [c_vp, c_int, C.POINTER(VgPeerDesc), c_int, c_vp, c_vp]),
    "vg_cast": (c_int, [c_vp, c_int, c_vp, c_int, c_ll, c_vp]),
    "vg_nchw_to_nhwc": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "vg_nhwc_to_nchw": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "vg_fill_zero": (c_int, [c_vp, C.c_size_t, c_vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None
_inited_devices = set()


class VgError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree shared library (no compute).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VgError(
            f"{LIB_PATH} not found: build it with `python -m vae_gan_b200.build` (or "
            "`__graft_entry__.build()`).  There is no CPU / library fallback for this path.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ensure_device(device: torch.device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        lib = load()
        rc = lib.vg_init(idx)
        if rc != 0:
            raise VgError(f"vg_init({idx}) failed ({rc}): {lib.vg_last_error().decode()}")
        _inited_devices.add(idx)
        if len(_inited_devices) == 1:
            from . import tune       # the tile table measured on B200 for the BASELINE shapes (VG_TILE_TABLE=0: heuristics only)
            tune.apply_default()
        if os.environ.get("VG_DETERMINISTIC", "0") == "1" and _det_buffers is None:
            set_deterministic(True, torch.device("cuda", idx))


_det_buffers = None      # (scratch, locks): caller-owned memory of the deterministic mode, alive while it is on


def set_deterministic(on: bool, device=None, scratch_mb: int = 64, n_locks: int = 1 << 16):
    """Bit-reproducible reductions (include/vaegan_b200.h, vg_set_deterministic): every cross-block floating-point sum -
    BatchNorm statistics / backward sums, split-K convolutions and Linear layers, weight gradients, bias gradients,
    spectral-norm dots, loss scalars - is accumulated in a fixed order (per-block partial sums added in block order, or
    the splits of one output tile taking turns) instead of with atomics, so two runs from the same state give identical
    bits.  Costs a few percent (profiles/README.md).  Also enabled by VG_DETERMINISTIC=1 at the first use of a device."""
    global _det_buffers
    lib = load()
    if not on:
        rc = lib.vg_set_deterministic(0, None, 0, None, 0)
        _det_buffers = None
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ensure_device(dev)
        scratch = torch.empty(scratch_mb << 20, dtype=torch.uint8, device=dev)
        locks = torch.zeros(n_locks, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        rc = lib.vg_set_deterministic(1, scratch.data_ptr(), scratch.numel(), locks.data_ptr(), n_locks)
        if rc == 0:
            _det_buffers = (scratch, locks)
    if rc != 0:
        raise VgError(f"vg_set_deterministic failed ({rc}): {lib.vg_last_error().decode()}")


def is_deterministic() -> bool:
    return bool(load().vg_get_deterministic())


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise VgError(f"{name} failed ({rc}): {lib.vg_last_error().decode()}")


def launch_count() -> int:
    return int(load().vg_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def vg_dtype(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return VG_F32
    if dt == torch.bfloat16:
        return VG_BF16
    raise VgError(f"unsupported dtype {dt}")
