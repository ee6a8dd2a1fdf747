"""Data-parallel plumbing: NVLink peer-memory exchange buffers for the SyncBN statistics.

`PeerExchange` allocates, on every rank, a data buffer double[n_slots][world][2048] and a flag buffer,
shares them with the other ranks of the node through CUDA IPC (the mechanism torch.multiprocessing uses
for CUDA tensors) and hands the peer-mapped pointers to `vg_peer_allreduce_f64` (include/vaegan_b200.h):
a one-shot all-reduce of the 2C BatchNorm sums that costs one NVLink round trip instead of an NCCL
collective.  torch.distributed is used only for the rendezvous (handle exchange)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib

MAX_N = 2048


class PeerExchange:
    def __init__(self, pg, device: torch.device, n_slots: int = 512):
        self.pg = pg
        self.world = dist.get_world_size(pg)
        self.rank = dist.get_rank(pg)
        assert self.world <= 8, "one NVSwitch node (<= 8 GPUs)"
        self.device = device
        self.n_slots = n_slots
        self.slot = 0
        _lib.ensure_device(device)
        data_bytes = n_slots * self.world * MAX_N * 8
        flag_bytes = n_slots * self.world * 8
        self._data, self._flags = C.c_void_p(), C.c_void_p()
        _lib.call("vg_peer_alloc", data_bytes, C.byref(self._data))
        _lib.call("vg_peer_alloc", flag_bytes, C.byref(self._flags))
        hd, hf = C.create_string_buffer(64), C.create_string_buffer(64)
        _lib.call("vg_peer_get_handle", self._data, hd)
        _lib.call("vg_peer_get_handle", self._flags, hf)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (device.index, bytes(hd.raw), bytes(hf.raw)), group=pg)
        self._opened = []
        desc = _lib.VgPeerDesc()
        for r, (dev_idx, h_data, h_flags) in enumerate(gathered):
            if r == self.rank:
                dptr, fptr = self._data.value, self._flags.value
            else:
                _lib.call("vg_enable_peer_access", int(dev_idx))
                pd_, pf_ = C.c_void_p(), C.c_void_p()
                _lib.call("vg_peer_open_handle", C.create_string_buffer(h_data, 64), C.byref(pd_))
                _lib.call("vg_peer_open_handle", C.create_string_buffer(h_flags, 64), C.byref(pf_))
                self._opened += [pd_, pf_]
                dptr, fptr = pd_.value, pf_.value
            desc.peer_data[r] = dptr
            desc.peer_flags[r] = fptr
        desc.rank, desc.world, desc.n_slots = self.rank, self.world, n_slots
        self.desc = desc
        # the exchange's OWN epoch: a device counter that only ever grows (one increment per iteration, captured in
        # the CUDA graph), independent of the Philox step that tests and capture() reset / restore
        self.epoch = torch.zeros(1, dtype=torch.int64, device=device)
        self._closed = False
        dist.barrier(group=pg)

    def reset(self):
        """Called once per training iteration (all ranks walk the same slot sequence): next epoch, slot 0."""
        self.slot = 0
        _lib.call("vg_counter_add", self.epoch.data_ptr(), 1, _lib.stream_ptr())

    def allreduce_(self, vec: torch.Tensor, epoch_tensor: torch.Tensor = None):
        assert vec.dtype == torch.float64 and vec.is_contiguous() and vec.numel() <= MAX_N
        if self.slot >= self.n_slots:
            raise _lib.VgError(f"more than {self.n_slots} SyncBN exchanges in one iteration")
        _lib.call("vg_peer_allreduce_f64", vec.data_ptr(), vec.numel(), C.byref(self.desc), self.slot, self.epoch.data_ptr(),
                  _lib.stream_ptr())
        self.slot += 1

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: every rank must have stopped using them)."""
        if self._closed:
            return
        self._closed = True
        try:
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                _lib.call("vg_peer_close_handle", p)
            _lib.call("vg_peer_free", self._data)
            _lib.call("vg_peer_free", self._flags)
        except Exception:      # interpreter / context teardown: nothing left to release
            pass

    def __del__(self):
        # peers may still have our buffers mapped at interpreter exit; only unmap ours, the driver frees the rest
        if not getattr(self, "_closed", True):
            try:
                for p in self._opened:
                    _lib.call("vg_peer_close_handle", p)
            except Exception:
                pass
