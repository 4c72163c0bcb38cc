"""Dynamic range of the fp16 split activation format (VERDICT r01 weak 10 / ADVICE): checkpoints whose activations
sit 2^10 above or below the O(1) fixture.  Uncalibrated, the engine must SAY so (overflow flag) or lose precision;
after `calibrate()` (static per-tensor power-of-two scales, re-packed weights) parity is back at 1e-3."""
import numpy as np
import pytest
import torch

from conftest import assert_close, max_rel
import siammask_b200 as smb
from oracle.calibrate import calibrated_state_dict, synthetic_inputs
from oracle.siammask_oracle import Oracle

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _engine(sd, **kw):
    m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, **kw)
    m.load_state_dict(sd)
    return m.eval().to("cuda")


def _run(m, z, x, pos):
    m.template(z.cuda())
    cls, loc, mask = m.track_mask(x.cuda())
    return cls, loc, mask, m.track_refine(pos)


@pytest.mark.parametrize("k", [10, -10])
def test_scaled_activations_need_and_get_calibration(k):
    sd = calibrated_state_dict(0, k)
    z, x = synthetic_inputs(81, 1)
    zc, xc = synthetic_inputs(82, 2)                 # calibration sample: different frames
    o = Oracle(sd)
    o.template(z)
    ocls, oloc, omask = o.track_mask(x)
    oref = o.track_refine((11, 13))
    m = _engine(sd, max_batch=2)
    cls, loc, mask, ref = _run(m, z, x, (11, 13))
    flagged = m.status() & 1
    worst = max(max_rel(cls, ocls), max_rel(loc, oloc), max_rel(ref, oref)) if torch.isfinite(cls).all() else float("inf")
    print(f"[range] 2^{k}: uncalibrated overflow flag {flagged}, worst error {worst:.2e}")
    if k > 0:
        assert flagged, "activations around 2^10..2^14 must trip the overflow flag without calibration"
    m.calibrate(zc, xc)
    cls, loc, mask, ref = _run(m, z, x, (11, 13))
    assert m.status() == 0
    assert_close(cls, ocls, TOL, f"cls, activations x 2^{k}, calibrated")
    assert_close(loc, oloc, TOL, f"loc, activations x 2^{k}, calibrated")
    assert_close(mask, omask, TOL, f"mask head, activations x 2^{k}, calibrated")
    assert_close(ref, oref, TOL, f"refine, activations x 2^{k}, calibrated")
    for i, name in enumerate(("p0", "p1", "p2", "p3")):
        assert_close(m.export(name), o.feature[i], TOL, f"{name}, activations x 2^{k}, calibrated")
    assert_close(m.export("corr_mask"), o.corr_feature, TOL, "corr feature (exported in true units)")


def test_calibration_keeps_parity_and_travels_with_the_weights(calib_sd, tmp_path):
    """On the O(1) fixture calibration must not hurt, a second calibrate() is a fixed point, and the scales are part
    of the packed-weight arena (save_packed / load_packed, i.e. also of the NCCL weight broadcast)."""
    z, x = synthetic_inputs(83, 2)
    o = Oracle(calib_sd)
    o.template(z)
    ocls, oloc, _ = o.track_mask(x, with_mask_head=False)
    pos = np.array([[5, 7], [20, 2]])
    oref = o.track_refine(pos)
    m = _engine(calib_sd, max_batch=2)
    m.calibrate(*synthetic_inputs(84, 2))
    m.template(z.cuda())
    cls, loc, _ = m.track_mask(x.cuda(), mask_head=False)
    ref = m.track_refine(pos)
    assert_close(cls, ocls, TOL, "cls after calibrate")
    assert_close(loc, oloc, TOL, "loc after calibrate")
    assert_close(ref, oref, TOL, "refine after calibrate")
    m.calibrate(*synthetic_inputs(84, 2))           # same sample again: nothing moves
    m.template(z.cuda())
    cls2, loc2, _ = m.track_mask(x.cuda(), mask_head=False)
    assert torch.equal(cls2, cls) and torch.equal(loc2, loc)
    path = str(tmp_path / "calibrated.smb")
    m.save_packed(path)
    b = smb.Custom(anchors=smb.DEFAULT_ANCHORS, max_batch=2).eval().to("cuda")
    b.load_packed(path)
    b.template(z.cuda())
    cls3, loc3, _ = b.track_mask(x.cuda(), mask_head=False)
    assert torch.equal(cls3, cls) and torch.equal(loc3, loc)
    assert torch.equal(b.track_refine(pos), ref)
    assert b.status() == 0
