"""Batched, device-resident tracker loop around the hot path: N concurrent tracker streams advance
frame to frame without the host touching per-stream state.

What the reference does per frame on the host in numpy, for ONE stream (tools/test.py:172-315) — search-window
arithmetic, crop + resize, score/box post-processing + argmax, learning-rate update and clamping of the target state,
mask paste-back — runs here as a fixed sequence of kernels over all streams (C ABI in include/siammask_b200.h):

    sm_tracker_prepare   state -> crop boxes, target size in the crop, scale          (tools/test.py:180-198, 71-76)
    sm_crop_resize       uint8 frames -> f32 [N,3,S,S] search crops (cv2-exact)        (:67-110)
    sm_step              track_mask -> select -> track_refine                          (:201-261)
    sm_tracker_update    winner box + score -> new state (lr, clamps), paste-back map  (:239-249, 263-282, 305-315)
    sm_warp_affine       127x127 sigmoid mask -> frame, threshold                      (:263-284)

The state (target_pos, target_sz, float64) lives on the device; a frame costs one small D2H copy only if the caller
asks for the numbers (`TrackResult.cpu()`).  Contour extraction / minAreaRect (:285-303) is not part of this module.
The arithmetic is pinned by `tests/test_batch_tracker.py` to the reference loop's golden trajectory and to
single-stream runs of the host restatement in `oracle/ref_loop.py`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .anchors import cosine_window, generate_anchor


@dataclass
class TrackerParams:
    """Hyper-parameters of the loop: defaults of utils/tracker_config.py:10-21 overlaid with the `hp` block of
    experiments/siammask_sharp/config_davis.json."""
    instance_size: int = 255
    exemplar_size: int = 127
    total_stride: int = 8
    base_size: int = 8
    out_size: int = 127              # 127 with the refine module, 63 for the plain mask head
    context_amount: float = 0.5
    penalty_k: float = 0.04
    window_influence: float = 0.4
    lr: float = 1.0
    seg_thr: float = 0.35
    windowing: str = "cosine"

    @property
    def score_size(self) -> int:     # utils/tracker_config.py:46
        return (self.instance_size - self.exemplar_size) // self.total_stride + 1 + self.base_size

    def c_struct(self) -> _lib.SmTrackerHp:
        return _lib.SmTrackerHp(self.context_amount, self.penalty_k, self.window_influence, self.lr, self.exemplar_size,
                                self.instance_size, self.total_stride, self.base_size, self.out_size, 0)


@dataclass
class TrackResult:
    """Per-frame outputs, all on the device.  state f64 [N,8] = x, y, w, h (new target_pos / target_sz), score,
    penalty, lr, best index; mask: bool [N,H,W] frame-sized masks (or None)."""
    state: torch.Tensor
    mask: torch.Tensor | None = None
    extras: dict = field(default_factory=dict)

    def cpu(self):
        s = self.state.cpu().numpy()
        return {"target_pos": s[:, 0:2].copy(), "target_sz": s[:, 2:4].copy(), "score": s[:, 4].copy(),
                "best_id": s[:, 7].astype(np.int64)}


class BatchTracker:
    """N tracker streams on one engine.  `net` is a `siammask_b200.Custom` on a CUDA device with
    max_batch >= N and num_slots >= slot0 + N."""

    def __init__(self, net, params: TrackerParams | None = None, slot0: int = 0):
        self.net = net
        self.p = params or TrackerParams(instance_size=net.search_size)
        if self.p.instance_size != net.search_size:
            raise ValueError("tracker instance_size must equal the engine's search_size")
        self.slot0 = int(slot0)
        self.dev = net._device
        self.lib = _lib.load()
        R, A = self.p.score_size, net.anchor_num
        self.anchors = torch.from_numpy(generate_anchor(net.anchors, R)).to(self.dev)
        self.window = torch.from_numpy(cosine_window(R, A, self.p.windowing).astype(np.float32)).to(self.dev)
        self.hp = self.p.c_struct()
        self.N = 0

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def _frames(self, frames) -> torch.Tensor:
        """uint8 [N,H,W,3] on the device (a single [H,W,3] frame is shared by all streams)."""
        if isinstance(frames, (list, tuple)):
            frames = np.stack([np.asarray(f) for f in frames], 0)
        t = torch.as_tensor(frames)
        if t.dtype != torch.uint8:
            raise ValueError("frames must be uint8 HWC (BGR as cv2.imread returns them)")
        return t.to(self.dev).contiguous()

    def _crop(self, frames: torch.Tensor, boxes: torch.Tensor, size: int) -> torch.Tensor:
        N = boxes.shape[0]
        if frames.dim() == 3:
            H, W, stride = frames.shape[0], frames.shape[1], 0
        else:
            if frames.shape[0] != N:
                raise ValueError("one frame per stream expected")
            H, W, stride = frames.shape[1], frames.shape[2], frames.shape[1] * frames.shape[2] * 3
        out = torch.empty(N, 3, size, size, device=self.dev, dtype=torch.float32)
        _lib.check(self.lib.sm_crop_resize(frames.data_ptr(), stride, H, W, boxes.data_ptr(), N, size, out.data_ptr(),
                                           self._stream()))
        return out

    # ------------------------------------------------------------------ siamese_init (tools/test.py:132-169)
    @torch.no_grad()
    def init(self, frames, boxes_xywh):
        """frames: uint8 [N,H,W,3] (or one shared [H,W,3]); boxes_xywh: [N,4] top-left x, y, w, h of the targets."""
        with torch.cuda.device(self.dev):
            fr = self._frames(frames)
            bx = torch.as_tensor(np.asarray(boxes_xywh, dtype=np.float64)).reshape(-1, 4).to(self.dev)
            N = bx.shape[0]
            if N > self.net.max_batch or self.slot0 + N > self.net.num_slots:
                raise ValueError("more streams than the engine was built for")
            self.N = N
            H, W = (fr.shape[0], fr.shape[1]) if fr.dim() == 3 else (fr.shape[1], fr.shape[2])
            self.im_w, self.im_h = int(W), int(H)
            self.imsize = torch.tensor([[W, H]] * N, dtype=torch.int32, device=self.dev)
            # target_pos = box centre, target_sz = (w, h)  (tools/test.py:338-339 / demo.py)
            self.state = torch.stack([bx[:, 0] + bx[:, 2] / 2, bx[:, 1] + bx[:, 3] / 2, bx[:, 2], bx[:, 3]], 1).contiguous()
            # avg_chans = np.mean(im, axis=(0, 1)); written into a uint8 image it truncates (:146, :89-100).
            # Sums of < 2^53 integers are exact in float64, so sum / n equals numpy's mean bit for bit.
            f4 = fr if fr.dim() == 4 else fr.unsqueeze(0).expand(N, -1, -1, -1)
            mean = f4.to(torch.float64).sum(dim=(1, 2)) / float(H * W)
            self.avg = mean.to(torch.uint8).to(torch.int32).contiguous()
            # template window (:149-155): s_z = round(sqrt(wc_z * hc_z)), crop around target_pos, resize to 127
            sw, sh = self.state[:, 2], self.state[:, 3]
            wc_z = sw + self.p.context_amount * (sw + sh)
            hc_z = sh + self.p.context_amount * (sw + sh)
            s_z = torch.round(torch.sqrt(wc_z * hc_z))                  # half-to-even, like Python's round()
            c = (s_z + 1) / 2
            zb = torch.zeros(N, 8, dtype=torch.int32, device=self.dev)
            zb[:, 0] = torch.round(self.state[:, 0] - c).to(torch.int32)
            zb[:, 1] = torch.round(self.state[:, 1] - c).to(torch.int32)
            zb[:, 2] = s_z.to(torch.int32)
            zb[:, 3:6] = self.avg
            z = self._crop(fr, zb, self.p.exemplar_size)
            self.net.template(z, slot0=self.slot0)
            self.boxes = torch.zeros(N, 8, dtype=torch.int32, device=self.dev)
            self.tsz = torch.zeros(N, 2, dtype=torch.float64, device=self.dev)
            self.aux = torch.zeros(N, 4, dtype=torch.float64, device=self.dev)
            self.maps = torch.zeros(N, 6, dtype=torch.float64, device=self.dev)
        return self

    # ------------------------------------------------------------------ siamese_track (tools/test.py:172-315)
    @torch.no_grad()
    def track(self, frames, mask: bool = True, refine: bool = True) -> TrackResult:
        """Advance all N streams by one frame.  mask=True pastes the (refined) mask back into the frame and thresholds it
        at seg_thr; refine=False uses the 63x63 mask head column instead of the refine module (tools/test.py:256-260)."""
        if self.N == 0:
            raise RuntimeError("call init() first")
        p, N = self.p, self.N
        with torch.cuda.device(self.dev):
            fr = self._frames(frames)
            st = self._stream()
            _lib.check(self.lib.sm_tracker_prepare(N, self.state.data_ptr(), self.avg.data_ptr(), C.byref(self.hp),
                                                   self.boxes.data_ptr(), self.tsz.data_ptr(), self.aux.data_ptr(), st))
            x = self._crop(fr, self.boxes, p.instance_size)
            use_refine = mask and refine
            use_head = mask and not refine
            out = self.net.step(x, self.anchors, self.window, self.tsz, p.penalty_k, p.window_influence, slot0=self.slot0,
                                refine=use_refine, mask_head=use_head, mask_col=use_head)
            res = torch.empty(N, 8, dtype=torch.float64, device=self.dev)
            _lib.check(self.lib.sm_tracker_update(N, self.state.data_ptr(), out["records"].data_ptr(), self.aux.data_ptr(),
                                                  self.imsize.data_ptr(), C.byref(self.hp), self.net.anchor_num,
                                                  p.score_size, self.maps.data_ptr() if mask else None, res.data_ptr(), st))
            mask_out = None
            if mask:
                logits = out["refine"] if use_refine else out["mask_col"]
                side = 127 if use_refine else 63
                if side != p.out_size:
                    raise ValueError(f"out_size {p.out_size} does not match the mask source ({side})")
                m = logits.sigmoid().view(N, side, side).contiguous()
                W, H = self.im_w, self.im_h
                pasted = torch.empty(N, H, W, device=self.dev, dtype=torch.float32)
                _lib.check(self.lib.sm_warp_affine(m.data_ptr(), side, side, self.maps.data_ptr(), pasted.data_ptr(), H, W,
                                                   C.c_float(-1.0), N, st))
                mask_out = pasted > p.seg_thr
            return TrackResult(state=res, mask=mask_out, extras={"records": out["records"], "pos": out["pos"], "x_crop": x})
