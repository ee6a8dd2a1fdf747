"""Tile table for the tensor-core convolutions ("shapes outside the tuned set (feature sizes
8-64 step 8, depth 1-8, num_blocks up to 16): autotuned tile table").

The library picks an N tile (64 / 128 / 256) and a kernel form (one-shot CTAs, persistent CTAs, persistent CTA pairs) per
convolution from heuristics measured on the BASELINE shapes.  For any other model / batch,

    table = tune.autotune(lambda: trainer.step(batch))      # records the shapes the step launches, times the candidates
    tune.save(table, "my_table.json");  tune.apply(tune.load("my_table.json"))

measures every candidate (N tile x form) of every distinct (layer shape, batch, direction) the step runs - on synthetic
tensors of that shape, CUDA events on the current stream, median of `iters` after a warm-up - and pins the fastest through
`vg_conv_tune_set` when it beats the heuristic by more than `min_gain`.  `vae_gan_b200/tile_table.json` is the table
measured on B200 for the BASELINE configurations; it is applied at import (VG_TILE_TABLE=0 disables it, VG_TILE_TABLE=path
loads another one).  Entries are keyed by the exact shape INCLUDING the batch, so a table never changes a shape it was not
measured on.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from pathlib import Path
from typing import Callable, Dict, List, Tuple

import torch

from . import _lib

Key = Tuple[int, ...]        # (dgrad, n, h_in, w_in, c_in, c_out, k, stride, pad, transposed)
FORMS = {0: "heuristic", 1: "one-shot", 2: "persistent", 3: "cta-pair"}
DEFAULT_TABLE = Path(__file__).resolve().parent / "tile_table.json"


def _desc(key: Key) -> _lib.VgConvDesc:
    dgrad, n, h, w, cin, cout, k, stride, pad, tr = key
    if tr:
        ho, wo = (h - 1) * stride - 2 * pad + k, (w - 1) * stride - 2 * pad + k
    else:
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    return _lib.VgConvDesc(n, h, w, cin, ho, wo, cout, k, k, stride, pad, tr, _lib.VG_BF16, _lib.VG_BF16)


def set_entry(key: Key, bn: int, form: int):
    _lib.call("vg_conv_tune_set", C.byref(_desc(key)), int(key[0]), int(bn), int(form))


def clear():
    _lib.call("vg_conv_tune_clear")


def record(fn: Callable[[], None]) -> List[Key]:
    """Run `fn` once and return the distinct (shape, direction) keys the tensor-core convolution path was called with."""
    _lib.call("vg_conv_tune_record", 1)
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        _lib.call("vg_conv_tune_record", 0)
    cnt = C.c_int(0)
    _lib.call("vg_conv_tune_seen", None, 0, C.byref(cnt))
    buf = (C.c_int * (10 * max(1, cnt.value)))()
    _lib.call("vg_conv_tune_seen", C.cast(buf, C.c_void_p), cnt.value, C.byref(cnt))
    return [tuple(buf[i * 10 + j] for j in range(10)) for i in range(cnt.value)]


def _time_key(key: Key, iters: int) -> float:
    """Median time (ms) of the convolution `key` with whatever table entry is currently set for it."""
    dgrad, n, h, w, cin, cout, k, stride, pad, tr = key
    dev = torch.device("cuda", torch.cuda.current_device())
    d = _desc(key)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, h, w, cin, generator=g, device=dev, dtype=torch.bfloat16)
    y = torch.empty(n, d.h_out, d.w_out, cout, device=dev, dtype=torch.bfloat16)
    wt = torch.randn(k * k * cin * cout, generator=g, device=dev) / (cin * k * k) ** 0.5
    pk = torch.empty(wt.numel(), dtype=torch.bfloat16, device=dev)
    pn = torch.empty_like(pk)
    s = _lib.stream_ptr()
    _lib.call("vg_conv_pack_weights", C.byref(d), wt.data_ptr(), None, pk.data_ptr(), pn.data_ptr(), s)
    if dgrad:
        dy = torch.randn(y.shape, generator=g, device=dev, dtype=torch.bfloat16)
        dx = torch.empty_like(x)
        run = lambda: _lib.call("vg_conv_dgrad", C.byref(d), dy.data_ptr(), pk.data_ptr(), pn.data_ptr(), dx.data_ptr(), s)
    else:
        run = lambda: _lib.call("vg_conv_forward", C.byref(d), x.data_ptr(), pk.data_ptr(), pn.data_ptr(), None, None, y.data_ptr(), None, s)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        run()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def autotune(fn: Callable[[], None] = None, keys: List[Key] = None, iters: int = 15, min_gain: float = 0.03, verbose: bool = False) -> Dict[Key, dict]:
    """Measure the candidates of every key (recorded from `fn`, or given) and pin the winners.  Returns
    {key: {"bn", "form", "ms", "heuristic_ms"}} for the keys where a candidate beat the heuristic by > min_gain."""
    if keys is None:
        keys = record(fn)
    table: Dict[Key, dict] = {}
    for key in keys:
        dgrad, n, h, w, cin, cout, k, stride, pad, tr = key
        n_out = cin if dgrad else cout
        if n_out % 64 != 0 or (h == 1 and w == 1):
            continue                      # single-channel outputs have one kernel; Linear layers ([B,1,1,C]) run split-K with fp32 outputs
        set_entry(key, 0, 0)
        base = _time_key(key, iters)
        best = (base, 0, 0)
        for bn in (64, 128, 256):
            if n_out % bn != 0:
                continue
            for form in (1, 2, 3):
                set_entry(key, bn, form)
                try:
                    t = _time_key(key, iters)
                except _lib.VgError:
                    continue
                if t < best[0]:
                    best = (t, bn, form)
        set_entry(key, 0, 0)
        if best[1] and best[0] < base * (1.0 - min_gain):
            set_entry(key, best[1], best[2])
            table[key] = {"bn": best[1], "form": best[2], "ms": round(best[0], 5), "heuristic_ms": round(base, 5)}
        if verbose:
            print(f"{'dgrad' if dgrad else 'fwd  '} n={n} {cin}->{cout} @{h}x{w} k{k} s{stride}{' T' if tr else ''}: heuristic {base * 1e3:.1f} us, "
                  f"best {best[0] * 1e3:.1f} us (bn {best[1] or '-'}, {FORMS[best[2]]})")
    return table


def save(table: Dict[Key, dict], path, meta: dict = None):
    rows = [{"key": list(k), **v} for k, v in sorted(table.items())]
    Path(path).write_text(json.dumps({"format": 1, "key": ["dgrad", "n", "h_in", "w_in", "c_in", "c_out", "k", "stride", "pad", "transposed"],
                                      "forms": FORMS, "meta": meta or {}, "entries": rows}, indent=1))


def load(path) -> Dict[Key, dict]:
    doc = json.loads(Path(path).read_text())
    return {tuple(r["key"]): {k: v for k, v in r.items() if k != "key"} for r in doc["entries"]}


def apply(table: Dict[Key, dict]):
    for key, v in table.items():
        set_entry(key, v["bn"], v["form"])


def apply_default():
    """Called once when the library is first used on a device: the committed B200 table (or VG_TILE_TABLE)."""
    sel = os.environ.get("VG_TILE_TABLE", "")
    if sel == "0":
        return 0
    path = Path(sel) if sel else DEFAULT_TABLE
    if not path.exists():
        return 0
    table = load(path)
    apply(table)
    return len(table)
