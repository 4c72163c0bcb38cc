"""SURVEY §8f row 2 — get_subwindow_tracking (tools/test.py:67-110): the numpy restatement of OpenCV's 8-bit
INTER_LINEAR resize against cv2 itself (CPU), and the CUDA crop+resize kernel against the cv2-based host path,
bit for bit, including windows that leave the frame (average-colour padding) and the no-resize case."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from oracle.cv_resize import resize_linear_u8
from oracle import ref_loop as tracker


@pytest.mark.parametrize("src,dst", [(180, 255), (300, 255), (90, 127), (411, 255), (64, 255), (700, 383), (127, 127)])
def test_resize_restatement_is_bit_exact(src, dst):
    rng = np.random.RandomState(src * 1000 + dst)
    img = rng.randint(0, 256, (src, src, 3)).astype(np.uint8)
    assert np.array_equal(cv2.resize(img, (dst, dst)), resize_linear_u8(img, (dst, dst)))


def test_subwindow_box_matches_reference_arithmetic():
    # Python-3 round() is round-half-even, exactly what tools/test.py:72,74 uses
    assert tracker.subwindow_box([100.5, 50.0], 128, [10.9, 20.2, 255.0]) == [36, -14, 128, 10, 20, 255]
    assert tracker.subwindow_box([101.5, 51.5], 127, [0.99, 1.0, 2.5]) == [38, -12, 127, 0, 1, 2]


@pytest.mark.gpu
def test_device_crop_matches_cv2_path():
    rng = np.random.RandomState(3)
    frame = rng.randint(0, 256, (240, 320, 3)).astype(np.uint8)
    fdev = torch.from_numpy(frame).cuda()
    avg = np.mean(frame, axis=(0, 1))
    cases = [((160.0, 120.0), 127, 150), ((10.2, 7.7), 255, 301), ((315.5, 236.5), 255, 97), ((160.3, 119.6), 127, 127),
             ((-20.0, 300.0), 255, 211), ((100.0, 100.0), 255, 640)]
    for pos, model, orig in cases:
        ref = tracker.get_subwindow_tracking(frame, pos, model, orig, avg)
        got = tracker.get_subwindow_tracking(fdev, pos, model, orig, avg)
        assert got.is_cuda and got.shape == (3, model, model)
        assert torch.equal(got.cpu(), ref), f"crop {pos} {orig}->{model}: max diff {(got.cpu() - ref).abs().max()}"
    # batched: several boxes on one frame
    from siammask_b200.ops import crop_resize
    boxes = [tracker.subwindow_box(p, o, avg) for p, m, o in cases if m == 255]
    out = crop_resize(fdev, boxes, 255)
    k = 0
    for p, m, o in cases:
        if m == 255:
            assert torch.equal(out[k].cpu(), tracker.get_subwindow_tracking(frame, p, 255, o, avg))
            k += 1


@pytest.mark.gpu
def test_loop_with_device_frames_equals_host_frames(calib_sd):
    """The tracker loop fed with GPU-resident frames (device crop) follows the same trajectory as with numpy frames."""
    import siammask_b200 as smb
    from oracle.synthetic_video import make_frames
    frames, boxes = make_frames()
    x, y, w, h = boxes[0]
    hp = {"instance_size": 255, "base_size": 8, "out_size": 127, "seg_thr": 0.35, "penalty_k": 0.04,
          "window_influence": 0.4, "lr": 1.0}
    traj = []
    for on_dev in (False, True):
        m = smb.Custom(anchors=smb.DEFAULT_ANCHORS).load_state_dict(calib_sd).eval().to("cuda")
        fs = [torch.from_numpy(f).cuda() for f in frames] if on_dev else frames
        st = tracker.siamese_init(fs[0], np.array([x + w / 2, y + h / 2]), np.array([w, h]), m, hp, device="cuda")
        if on_dev:      # paste-back etc. still run on the host with cv2: give the loop the numpy frame size info only
            st["im_h"], st["im_w"] = frames[0].shape[0], frames[0].shape[1]
        pos = []
        for f in fs[1:]:
            st = tracker.siamese_track(st, f, mask_enable=True, refine_enable=True, device="cuda")
            pos.append(st["target_pos"].copy())
        traj.append(np.asarray(pos))
    np.testing.assert_array_equal(traj[0], traj[1])


def _random_maps(rng, n):
    maps = []
    for _ in range(n):
        a, b = 0.3 + rng.rand() * 2.5, 0.3 + rng.rand() * 2.5
        maps.append(np.array([[a, 0, -rng.rand() * 120 + 20], [0, b, -rng.rand() * 90 + 15]], dtype=float))
    return maps


def test_warp_restatement_is_bit_exact():
    from oracle.cv_warp import warp_affine_f32
    rng = np.random.RandomState(11)
    for M in _random_maps(rng, 4):
        src = rng.rand(127, 127).astype(np.float32)
        ref = cv2.warpAffine(src, M, (320, 240), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=-1)
        assert np.array_equal(ref, warp_affine_f32(src, M, (320, 240)))


@pytest.mark.gpu
def test_device_paste_back_matches_cv2():
    """crop_back (tools/test.py:263-282) on the device == cv2.warpAffine, bit for bit."""
    from siammask_b200.ops import warp_affine
    rng = np.random.RandomState(12)
    maps = _random_maps(rng, 5)
    srcs = rng.rand(5, 127, 127).astype(np.float32)
    out = warp_affine(torch.from_numpy(srcs).cuda(), np.stack(maps), (854, 480)).cpu().numpy()
    for i, M in enumerate(maps):
        ref = cv2.warpAffine(srcs[i], M, (854, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=-1)
        assert np.array_equal(out[i], ref), f"map {i}: max diff {np.abs(out[i] - ref).max()}"


@pytest.mark.gpu
def test_loop_with_device_paste_back(calib_sd):
    """Whole loop with device crop + device select + device paste-back == the same loop with host cv2 steps."""
    import siammask_b200 as smb
    from oracle.synthetic_video import make_frames
    frames, boxes = make_frames()
    x, y, w, h = boxes[0]
    hp = {"instance_size": 255, "base_size": 8, "out_size": 127, "seg_thr": 0.35, "penalty_k": 0.04,
          "window_influence": 0.4, "lr": 1.0}
    res = []
    for dev_path in (False, True):
        m = smb.Custom(anchors=smb.DEFAULT_ANCHORS).load_state_dict(calib_sd).eval().to("cuda")
        fs = [torch.from_numpy(f).cuda() for f in frames] if dev_path else frames
        st = tracker.siamese_init(fs[0], np.array([x + w / 2, y + h / 2]), np.array([w, h]), m, hp, device="cuda")
        rec = []
        for f in fs[1:]:
            st = tracker.siamese_track(st, f, mask_enable=True, refine_enable=True, device="cuda", device_paste=dev_path)
            mk = st["mask"]
            mk = mk.cpu().numpy() if isinstance(mk, torch.Tensor) else mk
            rec.append((st["target_pos"].copy(), mk.copy(), np.asarray(st["ploygon"]).copy()))
        res.append(rec)
    for (p0, m0, g0), (p1, m1, g1) in zip(*res):
        np.testing.assert_array_equal(p0, p1)
        np.testing.assert_array_equal(m0, m1)
        np.testing.assert_allclose(g0, g1, atol=1e-4)
