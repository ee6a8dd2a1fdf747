"""Loader for the REAL reference (Don-Yin/VAE-GAN notebook) -- TEST INFRASTRUCTURE ONLY.

This file only works inside the build container, where the read-only reference
tree is mounted at /root/reference.  It `exec`s the notebook's model/training
code cells (nothing is copied into this repo) and returns their namespace.
It is used by `oracle/make_golden.py` (to generate tests/golden/*.pt) and by the
CPU tests that pin `oracle/vaegan_oracle.py` against the executed reference.

The GPU box has no /root/reference: nothing on the product path, in `-m gpu`
tests, in `smoke()` or in `bench.py` may import this module.

Cells used (0-based index among code cells; README.md line ranges for citation):
  3: ResBlockVAE / Encoder / Decoder                (README.md:119-295)
  4: ResBlockDiscriminator / Discriminator          (README.md:350-499)
  5: SpatialVAECodeProcessor / UnsupervisedGeneratorNetwork (README.md:522-668)
  6 (up to `generator = experiment(`): init_weights, compute_gradient_penalty,
     train_network_wgan, experiment                 (README.md:690-935)
"""
from __future__ import annotations

import json
import os

REFERENCE_ROOT = os.environ.get("VAEGAN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gan.ipynb"))


def load_reference_namespace() -> dict:
    """Exec the reference notebook's model + training cells; return the namespace."""
    if not reference_available():
        raise FileNotFoundError(f"reference notebook not found under {REFERENCE_ROOT}")
    with open(os.path.join(REFERENCE_ROOT, "gan.ipynb")) as f:
        nb = json.load(f)
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    ns: dict = {}
    exec("import os, torch, numpy as np\nimport torch.nn as nn\nfrom pathlib import Path\n", ns)
    exec(cells[3], ns)
    exec(cells[4], ns)
    exec(cells[5], ns)
    src = cells[6]
    ns["dataset_loader"] = None
    exec(src[: src.index("generator = experiment(")], ns)
    return ns


def build_reference_models(ns, *, depth=2, length=1, feature_size=64, image_size=96,
                           disc_params=None, seed=0):
    """Construct G and D exactly as `experiment()` does (README.md:882-907), plus the
    non-256 input adapter for D.linear_1 (the reference hard-codes 256x256)."""
    import torch
    import torch.nn as nn

    if disc_params is None:
        disc_params = dict(num_stride_conv1=1, num_features_conv1=feature_size,
                           num_blocks=[1, 1, 1], num_strides_res=[1, 2, 2],
                           num_features_res=[2 * feature_size, 4 * feature_size, 8 * feature_size])
    torch.manual_seed(seed)
    feature_depth = feature_size * (2 ** depth)
    G = ns["UnsupervisedGeneratorNetwork"](
        encoder=ns["Encoder"](in_channels=1, depth=depth, length=length, feature_size=feature_size),
        decoder=ns["Decoder"](in_channels=feature_depth, depth=depth, length=length,
                              reconstruction_channels=1),
        code_processor=ns["SpatialVAECodeProcessor"](feature_depth=feature_depth, is_training=True),
        is_vae=True,
    )
    D = ns["Discriminator"](block=ns["ResBlockDiscriminator"], **disc_params)
    if image_size != 256:
        import numpy as np
        side = image_size // disc_params["num_stride_conv1"] // 4 // int(np.prod(disc_params["num_strides_res"]))
        D.linear_len = side * side * disc_params["num_features_res"][-1]
        D.linear_1 = nn.Linear(int(D.linear_len), 1024)
    G.apply(ns["init_weights"])
    D.apply(ns["init_weights"])
    return G, D, disc_params


class MaskFeed:
    """Replacement for nn.Dropout / nn.Dropout2d inside reference blocks: multiplies by a
    supplied keep-mask (already scaled by 1/(1-p)) in train mode, identity in eval."""

    def __new__(cls, scaled_mask):
        import torch.nn as nn

        class _MaskFeed(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, x):
                if not self.training or self.m is None:
                    return x
                return x * self.m.to(x.dtype)

        return _MaskFeed(scaled_mask)
