"""Checkpoint ingest for the SiamMask hot path.

* `load_checkpoint(path)` mirrors the reference loader's key handling
  (utils/load_helper.py:30-54): accepts a bare state dict or {'state_dict': ...},
  strips a leading 'module.' and, when nothing matches, retries with a 'features.'
  prefix (a backbone-only checkpoint).
* `synthetic_state_dict(seed)` builds a random-init checkpoint with exactly the
  reference's 356 keys/shapes — there is no network and no
  pretrained .pth on the box, so benchmarks and tests run on seeded weights.
  Values are deterministic for a given seed and torch build (CPU generator).
"""
from __future__ import annotations

import math
import torch

# (layer, planes, blocks) — experiments/siammask_sharp/resnet.py:159-165, resnet50 = [3, 4, 6, 3]
_LAYERS = (("layer1", 64, 3), ("layer2", 128, 4), ("layer3", 256, 6))
_DS_KERNEL = {"layer1": 1, "layer2": 3, "layer3": 3}   # resnet.py:184-215
REFINE_SHAPES = {  # custom.py:102-124  (Cout, Cin)
    "v0": [(16, 64), (4, 16)], "v1": [(64, 256), (16, 64)], "v2": [(128, 512), (32, 128)],
    "h2": [(32, 32), (32, 32)], "h1": [(16, 16), (16, 16)], "h0": [(4, 4), (4, 4)],
    "post0": (16, 32), "post1": (4, 16), "post2": (1, 4),
}


def expected_keys(mask: bool = True, refine: bool = True) -> dict[str, tuple[int, ...]]:
    """name -> shape for every tensor the hot path consumes (num_batches_tracked excluded)."""
    keys: dict[str, tuple[int, ...]] = {}

    def bn(p, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            keys[f"{p}.{s}"] = (c,)

    F = "features.features."
    keys[F + "conv1.weight"] = (64, 3, 7, 7)
    bn(F + "bn1", 64)
    inplanes = 64
    for name, planes, blocks in _LAYERS:
        for i in range(blocks):
            p = f"{F}{name}.{i}."
            keys[p + "conv1.weight"] = (planes, inplanes, 1, 1); bn(p + "bn1", planes)
            keys[p + "conv2.weight"] = (planes, planes, 3, 3); bn(p + "bn2", planes)
            keys[p + "conv3.weight"] = (planes * 4, planes, 1, 1); bn(p + "bn3", planes * 4)
            if i == 0:
                k = _DS_KERNEL[name]
                keys[p + "downsample.0.weight"] = (planes * 4, inplanes, k, k)
                bn(p + "downsample.1", planes * 4)
            inplanes = planes * 4
    keys["features.downsample.downsample.0.weight"] = (256, 1024, 1, 1)
    bn("features.downsample.downsample.1", 256)
    heads = [("rpn_model.cls.", 10), ("rpn_model.loc.", 20)]
    if mask:
        heads.append(("mask_model.mask.", 63 * 63))
    for p, nout in heads:
        for br in ("conv_kernel", "conv_search"):
            keys[f"{p}{br}.0.weight"] = (256, 256, 3, 3); bn(f"{p}{br}.1", 256)
        keys[p + "head.0.weight"] = (256, 256, 1, 1); bn(p + "head.1", 256)
        keys[p + "head.3.weight"] = (nout, 256, 1, 1)
        keys[p + "head.3.bias"] = (nout,)
    if refine:
        R = "refine_model."
        for n, shp in REFINE_SHAPES.items():
            if isinstance(shp, list):
                for idx, (co, ci) in zip((0, 2), shp):
                    keys[f"{R}{n}.{idx}.weight"] = (co, ci, 3, 3)
                    keys[f"{R}{n}.{idx}.bias"] = (co,)
            else:
                keys[f"{R}{n}.weight"] = (shp[0], shp[1], 3, 3)
                keys[f"{R}{n}.bias"] = (shp[0],)
        keys[R + "deconv.weight"] = (256, 32, 15, 15)
        keys[R + "deconv.bias"] = (32,)
    return keys


def synthetic_state_dict(seed: int = 0, mask: bool = True, refine: bool = True) -> dict[str, torch.Tensor]:
    """Seeded random-init weights with the reference's key layout.

    Conv weights ~ N(0, 2/fan_in); BN gamma ~ U[0.5,1.5] (bn3 / downsample BN: U[0.2,0.4] so the
    residual stream stays O(1) without a calibration pass), beta ~ U[-0.5,0.5], running_mean 0,
    running_var 1.  `oracle/calibrate.py` refines the running statistics for the parity tests."""
    g = torch.Generator().manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    for k, shp in expected_keys(mask, refine).items():
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 1 and k.endswith(".weight"):      # BN gamma
            small = ".bn3." in k or "downsample.1." in k
            lo, hi = (0.2, 0.4) if small else (0.5, 1.5)
            sd[k] = torch.rand(shp, generator=g) * (hi - lo) + lo
        elif len(shp) == 1:                                 # BN beta / conv bias
            amp = 0.1 if k.startswith("refine_model") else 0.5
            sd[k] = (torch.rand(shp, generator=g) - 0.5) * 2 * amp
        elif k == "refine_model.deconv.weight":             # ConvTranspose2d: (Cin, Cout, 15, 15)
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(1.0 / shp[0])
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in)
    # the raw-pixel stem sees inputs in [0,255] (tools/test.py:61-64): bring bn1's input to O(1)
    sd["features.features.conv1.weight"] /= 74.0
    return sd


def normalize_keys(obj) -> dict[str, torch.Tensor]:
    """utils/load_helper.py:38-41 — unwrap {'state_dict':…}, strip 'module.'."""
    sd = obj["state_dict"] if isinstance(obj, dict) and "state_dict" in obj else obj
    return {(k.split("module.", 1)[-1] if k.startswith("module.") else k): v for k, v in sd.items()}


def load_checkpoint(path: str) -> dict[str, torch.Tensor]:
    sd = normalize_keys(torch.load(path, map_location="cpu"))
    want = expected_keys(True, True)
    if not (set(sd) & set(want)):                            # load_helper.py:43-52
        sd = {"features." + k: v for k, v in sd.items()}
        assert set(sd) & set(want), "load NONE from pretrained checkpoint"
    return sd
