import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import siammask_b200 as smb
from siammask_b200 import anchors as tracker
prec = sys.argv[1] if len(sys.argv) > 1 else 'fast'
dev = torch.device('cuda', 0)
B, S, R = 64, 255, 25
m = smb.Custom(anchors=smb.DEFAULT_ANCHORS, max_batch=B, num_slots=B, precision=prec).load_state_dict(smb.synthetic_state_dict(0)).eval().to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
z = torch.rand(B, 3, 127, 127, device=dev, generator=gen) * 255
xs = [torch.rand(B, 3, S, S, device=dev, generator=gen) * 255 for _ in range(4)]
pos = torch.randint(0, R, (B, 2), device=dev, generator=gen, dtype=torch.int32)
anchors_dev = torch.from_numpy(tracker.generate_anchor(smb.DEFAULT_ANCHORS, R)).to(dev)
window_dev = torch.from_numpy(np.tile(np.outer(np.hanning(R), np.hanning(R)).flatten(), 5).astype(np.float32)).to(dev)
tsz_dev = torch.rand(B, 2, device=dev, generator=gen) * 60 + 30
m.template(z)
def timeit(name, fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n:8.3f} ms/step  (host wall {1e3*(time.perf_counter()-t0)/n:8.3f})", flush=True)
def a(i): m.track_mask(xs[i % 4], mask_head=True)
def b(i): m.track_mask(xs[i % 4], mask_head=False)
def c(i):
    cls, loc, mask = m.track_mask(xs[i % 4], mask_head=True); m.track_refine(pos)
def d(i):
    cls, loc, mask = m.track_mask(xs[i % 4], mask_head=True); m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)
def e(i):
    cls, loc, mask = m.track_mask(xs[i % 4], mask_head=True); best, sp, rec = m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4); m.track_refine(sp)
def f(i):
    cls, loc, mask = m.track_mask(xs[i % 4], mask_head=False); best, sp, rec = m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4); m.track_refine(sp)
def g(i):
    cls, loc, mask = m.track_mask(xs[i % 4], mask_head=False); m.select(cls, loc, anchors_dev, window_dev, tsz_dev, 0.04, 0.4)
for name, fn in [('track_mask+head', a), ('track_mask nohead', b), ('head+refine(randpos)', c), ('head+select', d), ('head+select+refine', e), ('nohead+select+refine', f), ('nohead+select', g), ('track_mask+head again', a)]:
    timeit(name, fn)
print(torch.cuda.memory_stats()['num_alloc_retries'], torch.cuda.memory_stats()['num_device_alloc'])
