"""Small-batch latency of one whole frame (sm_step: track_mask + selection + refine), eager vs CUDA-graph replay,
plus the host time of the call itself (enqueue only).  Used for profiles/r02_other_configs.md."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import siammask_b200 as smb
from siammask_b200 import anchors as anc
dev = torch.device('cuda', 0)
R = 25
anchors_dev = torch.from_numpy(anc.generate_anchor(smb.DEFAULT_ANCHORS, R)).to(dev)
window_dev = torch.from_numpy(anc.cosine_window(R, 5).astype(np.float32)).to(dev)
ONLY = os.environ.get("SMB200_LAT_ONLY", "")          # "exact1": exact mode, B=1 only (short GPU calls)
for prec in (('exact',) if ONLY == "exact1" else ('exact', 'fast')):
  for B in ((1,) if ONLY == "exact1" else (1, 8)):
   for graphs in (False, True):
    m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, max_batch=B, num_slots=B, precision=prec, graphs=graphs).load_state_dict(smb.synthetic_state_dict(0)).eval().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    z = torch.rand(B, 3, 127, 127, device=dev, generator=gen) * 255
    x = torch.rand(B, 3, 255, 255, device=dev, generator=gen) * 255
    tsz = (torch.rand(B, 2, device=dev, generator=gen) * 60 + 30).double()
    m.template(z)
    def step():
        return m.step(x, anchors_dev, window_dev, tsz, 0.04, 0.4, refine=True, mask_head=False)["refine"]
    for _ in range(5): step()
    torch.cuda.synchronize()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n): step()
    t_host = (time.perf_counter() - t0) / n          # enqueue cost (the GPU runs behind)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        step().cpu()                                 # a tracker loop reads results on the host each frame
    dl = (time.perf_counter() - t0) / n
    print(f"{prec:6s} graphs={graphs!s:5s} B={B:3d}: pipelined {1e3*dt:7.3f} ms/step = {B/dt:9.1f} FPS | synced {1e3*dl:7.3f} ms/step = {B/dl:9.1f} FPS | host enqueue {1e3*t_host:6.3f} ms", flush=True)
    del m
