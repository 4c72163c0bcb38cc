import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import siammask_b200 as smb
from siammask_b200 import tracker
dev = torch.device('cuda', 0)
for prec in ('exact', 'fast'):
  for B in (1, 8):
   for graphs in (False, True):
    m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, max_batch=B, num_slots=B, precision=prec, graphs=graphs).load_state_dict(smb.synthetic_state_dict(0)).eval().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    z = torch.rand(B, 3, 127, 127, device=dev, generator=gen) * 255
    x = torch.rand(B, 3, 255, 255, device=dev, generator=gen) * 255
    R = 25
    anchors_dev = torch.from_numpy(tracker.generate_anchor(smb.DEFAULT_ANCHORS, R)).to(dev)
    window_dev = torch.from_numpy(np.tile(np.outer(np.hanning(R), np.hanning(R)).flatten(), 5).astype(np.float32)).to(dev)
    tsz_dev = torch.rand(B, 2, device=dev, generator=gen) * 60 + 30
    m.template(z)
    def step():
        cls, loc, _ = m.track_mask(x, mask_head=False)
        best, sp, rec = m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)
        return m.track_refine(sp)
    for _ in range(5): step()
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    # latency with a sync every frame (the tracker loop reads results on the host each frame)
    t0 = time.perf_counter()
    for _ in range(n):
        step().cpu()
    dl = (time.perf_counter() - t0) / n
    print(f"{prec:6s} graphs={graphs!s:5s} B={B:3d}: pipelined {1e3*dt:7.3f} ms/step = {B/dt:9.1f} FPS | synced {1e3*dl:7.3f} ms/step = {B/dl:9.1f} FPS", flush=True)
    del m
