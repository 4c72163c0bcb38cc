"""Standalone operators of the hot path, bound to the C ABI (CUDA tensors in, CUDA tensors out).

`conv2d_dw_group` keeps the reference's name and argument order (models/rpn.py:32-38)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def conv2d_dw_group(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """Depthwise cross-correlation: x f32[B,C,H,W], kernel f32[B,C,kh,kw] -> f32[B,C,H-kh+1,W-kw+1].
    Like the reference it requires paired batches (kernel.shape[:2] == x.shape[:2])."""
    if not (x.is_cuda and kernel.is_cuda):
        raise RuntimeError("siammask_b200 operators run on CUDA tensors only; there is no CPU path")
    if x.shape[:2] != kernel.shape[:2]:
        raise RuntimeError(f"paired batch required: x {tuple(x.shape)} vs kernel {tuple(kernel.shape)}")
    lib = _lib.load()
    x = x.to(torch.float32).contiguous()
    kernel = kernel.to(torch.float32).contiguous()
    B, Cn, H, W = x.shape
    kh, kw = kernel.shape[2:]
    out = torch.empty(B, Cn, H - kh + 1, W - kw + 1, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.sm_xcorr_depthwise(x.data_ptr(), kernel.data_ptr(), out.data_ptr(), B, Cn, H, W, kh, kw,
                                          _stream(x.device)))
    return out


xcorr_depthwise = conv2d_dw_group


def conv2d(x, weight, scale=None, shift=None, stride=1, padding=0, dilation=1, relu=False, backend="tensor",
           precision="exact"):
    """F.conv2d(x, weight) * scale[c] + shift[c] (+ReLU) through the engine's convolution kernels.
    x f32[B,Cin,H,W] NCHW, weight f32[Cout,Cin,KH,KW]; returns f32 NCHW."""
    if not x.is_cuda:
        raise RuntimeError("siammask_b200 operators run on CUDA tensors only; there is no CPU path")
    lib = _lib.load()
    dev = x.device
    x = x.to(torch.float32).contiguous()
    weight = weight.to(dev, torch.float32).contiguous()
    B, Cin, H, W = x.shape
    Cout, _, KH, KW = weight.shape
    Ho = (H + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wo = (W + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    out = torch.empty(B, Cout, Ho, Wo, device=dev, dtype=torch.float32)
    sc = scale.to(dev, torch.float32).contiguous() if scale is not None else None
    sh = shift.to(dev, torch.float32).contiguous() if shift is not None else None
    be = {"tensor": _lib.SM_BACKEND_TENSOR, "simt": _lib.SM_BACKEND_SIMT}[backend]
    pr = {"exact": _lib.SM_PRECISION_EXACT, "fast": _lib.SM_PRECISION_FAST}[precision]
    with torch.cuda.device(dev):
        _lib.check(lib.sm_conv2d(x.data_ptr(), weight.data_ptr(), sc.data_ptr() if sc is not None else None,
                                 sh.data_ptr() if sh is not None else None, out.data_ptr(), B, Cin, H, W, Cout, KH, KW,
                                 stride, padding, dilation, int(relu), be, pr, _stream(dev)))
    return out


def crop_resize(frames: torch.Tensor, boxes, model_size: int) -> torch.Tensor:
    """Device form of get_subwindow_tracking (tools/test.py:67-110).  frames: uint8 CUDA tensor [H,W,3] (shared by all
    boxes) or [B,H,W,3]; boxes: int [B,6] = (context_xmin, context_ymin, original_sz, avg0, avg1, avg2) in frame
    coordinates before padding.  Returns f32 [B,3,model,model], bit-identical to the cv2 path."""
    if not frames.is_cuda or frames.dtype != torch.uint8:
        raise RuntimeError("crop_resize expects uint8 CUDA frames; there is no CPU path")
    lib = _lib.load()
    dev = frames.device
    frames = frames.contiguous()
    bx = torch.as_tensor(boxes, dtype=torch.int32).reshape(-1, 6)
    B = bx.shape[0]
    full = torch.zeros(B, 8, dtype=torch.int32)
    full[:, :6] = bx
    full = full.to(dev)
    if frames.dim() == 3:
        H, W, stride = frames.shape[0], frames.shape[1], 0
    else:
        if frames.shape[0] != B:
            raise ValueError("one frame per box expected")
        H, W, stride = frames.shape[1], frames.shape[2], frames.shape[1] * frames.shape[2] * 3
    out = torch.empty(B, 3, model_size, model_size, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.sm_crop_resize(frames.data_ptr(), stride, H, W, full.data_ptr(), B, model_size, out.data_ptr(),
                                      _stream(dev)))
    return out


def warp_affine(src: torch.Tensor, maps, dsize, border_value: float = -1.0) -> torch.Tensor:
    """cv2.warpAffine(src, M, dsize, INTER_LINEAR, BORDER_CONSTANT, border_value) on the device (crop_back,
    tools/test.py:263-275).  src f32 CUDA [h,w] or [B,h,w]; maps: forward 2x3 map(s); dsize = (width, height)."""
    if not src.is_cuda:
        raise RuntimeError("warp_affine expects a CUDA tensor; there is no CPU path")
    lib = _lib.load()
    dev = src.device
    squeeze = src.dim() == 2
    s3 = (src.unsqueeze(0) if squeeze else src).to(torch.float32).contiguous()
    B, sh, sw = s3.shape
    m = torch.as_tensor(maps, dtype=torch.float64).reshape(-1, 6)
    if m.shape[0] != B:
        raise ValueError("one 2x3 map per image expected")
    m = m.to(dev).contiguous()
    dw, dh = int(dsize[0]), int(dsize[1])
    out = torch.empty(B, dh, dw, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.sm_warp_affine(s3.data_ptr(), sh, sw, m.data_ptr(), out.data_ptr(), dh, dw, float(border_value), B,
                                      _stream(dev)))
    return out[0] if squeeze else out
