"""siammask_b200 — B200-native (sm_100a) implementation of SiamMask's per-frame inference hot path.

Public surface mirrors the reference's model object for that path:
    Custom.template / track / track_mask / track_refine   (experiments/siammask_sharp/custom.py:173-190)
    conv2d_dw_group                                       (models/rpn.py:32-38)
All compute lives in libsiammask_b200.so (C ABI: include/siammask_b200.h)."""
from .custom import Custom, DEFAULT_ANCHORS
from .ops import conv2d_dw_group, xcorr_depthwise, conv2d, crop_resize, warp_affine
from .checkpoint import synthetic_state_dict, load_checkpoint, expected_keys

__all__ = ["Custom", "DEFAULT_ANCHORS", "conv2d_dw_group", "xcorr_depthwise", "conv2d", "crop_resize", "warp_affine",
           "synthetic_state_dict", "load_checkpoint", "expected_keys"]
