"""Per-tracker constants the on-device score/box selection consumes (C ABI `sm_select` / `sm_step`): the anchor table
and the cosine window, in the element order the network's cls/loc outputs use — (anchor, y, x), flattened.

What they must equal (checked in tests against the reference-loop restatement in `oracle/ref_loop.py`):
  * anchor table: `generate_anchor`, tools/test.py:113-129 on top of `Anchors.generate_anchors`, utils/anchors.py:26-48
    (anchor_density 1) — centre (cx, cy) on a stride grid centred at 0, size (w, h) per ratio/scale, float32;
  * window: tiled outer product of two Hanning windows, tools/test.py:157-161.
"""
from __future__ import annotations

import math

import numpy as np


def anchor_sizes(cfg: dict) -> np.ndarray:
    """(A, 2) float32 (w, h) of the A = len(ratios) * len(scales) base anchors, ratio-major."""
    stride, rd = cfg.get("stride", 8), cfg.get("round_dight", 0)
    area = float(stride * stride)
    sizes = []
    for ratio in cfg["ratios"]:
        w0 = math.sqrt(area / ratio)
        if rd > 0:
            w0 = round(w0, rd)
            h0 = round(w0 * ratio, rd)
        else:
            w0 = int(w0)
            h0 = int(w0 * ratio)
        sizes += [(w0 * sc, h0 * sc) for sc in cfg["scales"]]
    wh = np.asarray(sizes, dtype=np.float32)
    # the reference stores corners (-w/2, -h/2, w/2, h/2) in float32 and takes size = x2 - x1: same rounding here
    half = (wh * np.float32(0.5)).astype(np.float32)
    return (half - (-half)).astype(np.float32)


def generate_anchor(cfg: dict, score_size: int) -> np.ndarray:
    """(A * score_size**2, 4) float32 rows (cx, cy, w, h), ordered (anchor, y, x)."""
    wh = anchor_sizes(cfg)
    A, R, stride = wh.shape[0], int(score_size), cfg.get("stride", 8)
    centre = (np.arange(R, dtype=np.int64) - R // 2) * stride          # grid centred on the search crop
    out = np.empty((A, R, R, 4), dtype=np.float32)
    out[..., 0] = centre[None, None, :]
    out[..., 1] = centre[None, :, None]
    out[..., 2] = wh[:, 0, None, None]
    out[..., 3] = wh[:, 1, None, None]
    return out.reshape(A * R * R, 4)


def cosine_window(score_size: int, anchor_num: int, windowing: str = "cosine") -> np.ndarray:
    """(A * score_size**2,) float64 window, tiled over the anchors."""
    if windowing == "cosine":
        w = np.outer(np.hanning(score_size), np.hanning(score_size))
    else:
        w = np.ones((score_size, score_size))
    return np.tile(w.flatten(), anchor_num)
