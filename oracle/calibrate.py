"""Seeded, BN-calibrated synthetic checkpoint for the parity tests.  TEST INFRASTRUCTURE ONLY.

With default-initialised weights and raw 0..255 inputs the reference's activations reach
1e6-1e8, which makes a relative tolerance meaningless.  This takes the
product's `synthetic_state_dict(seed)` and runs ONE oracle pass over seeded inputs with a
batch-norm hook that sets every BN's running_mean / running_var to the statistics of its own
input, so each BN output is ~N(beta, gamma^2) and every activation on the path is O(1).
"""
from __future__ import annotations

import functools
import torch

from siammask_b200.checkpoint import synthetic_state_dict
from .siammask_oracle import Oracle


def _hook(prefixes):
    def hook(key, x, sd):
        if not key.startswith(prefixes):
            return
        sd[key + ".running_mean"] = x.mean(dim=(0, 2, 3)).clone()
        sd[key + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-4).clone()
    return hook


def synthetic_inputs(seed: int, batch: int, search: int = 255):
    """z f32[B,3,127,127], x f32[B,3,S,S], raw pixel range (tools/test.py:61-64)."""
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(batch, 3, 127, 127, generator=g) * 255.0
    x = torch.rand(batch, 3, search, search, generator=g) * 255.0
    return z, x


@functools.lru_cache(maxsize=4)
def calibrated_state_dict(seed: int = 0, log2_scale: int = 0):
    """log2_scale != 0: every BN's gamma / beta is multiplied by 2^log2_scale BEFORE the statistics pass, so every
    activation of the network is ~2^log2_scale times larger (smaller) than in the O(1) fixture — the dynamic-range
    fixture for the engine's fp16 split activation format."""
    sd = synthetic_state_dict(seed)
    if log2_scale != 0:
        f = 2.0 ** log2_scale
        for k in [k for k in sd if k.endswith(".running_mean")]:
            base = k[:-len(".running_mean")]
            sd[base + ".weight"] = sd[base + ".weight"] * f
            sd[base + ".bias"] = sd[base + ".bias"] * f
    z, x = synthetic_inputs(seed + 1000, 2)
    with torch.no_grad():
        # 1) backbone + ResDownS statistics from the search crops
        o = Oracle(sd, bn_hook=_hook(("features.",)))
        o.features_and_resdown = o.resdown(o.features(x)[-1])
        # 2) template / search features with the calibrated backbone, then the head BNs
        o = Oracle(sd)
        o.template(z)
        o.bn_hook = _hook(("rpn_model.", "mask_model."))
        xf = o.resdown(o.features(x)[-1])
        for p in ("rpn_model.cls.", "rpn_model.loc.", "mask_model.mask."):
            o.head(o.forward_corr(o.zf, xf, p), p)
    return sd
