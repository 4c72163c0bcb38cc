import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import siammask_b200 as smb
from siammask_b200 import anchors as tracker
dev = torch.device('cuda', 0)
B, R = 64, 25
for graphs in (False, True):
    m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, max_batch=B, num_slots=B, precision='exact', graphs=graphs).load_state_dict(smb.synthetic_state_dict(0)).eval().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    z = torch.rand(B, 3, 127, 127, device=dev, generator=gen) * 255
    xs = [torch.rand(B, 3, 255, 255, device=dev, generator=gen) * 255 for _ in range(4)]
    anchors_dev = torch.from_numpy(tracker.generate_anchor(smb.DEFAULT_ANCHORS, R)).to(dev)
    window_dev = torch.from_numpy(np.tile(np.outer(np.hanning(R), np.hanning(R)).flatten(), 5).astype(np.float32)).to(dev)
    tsz_dev = torch.rand(B, 2, device=dev, generator=gen) * 60 + 30
    m.template(z)
    def step(i):
        cls, loc, mask = m.track_mask(xs[i % 4], mask_head=True)
        best, sp, rec = m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)
        return m.track_refine(sp)
    for i in range(6): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30): step(i)
    e1.record(); torch.cuda.synchronize()
    print(f"graphs={graphs}: {e0.elapsed_time(e1)/30:.3f} ms/step = {B*30/(e0.elapsed_time(e1)*1e-3):.0f} FPS", flush=True)
    del m
