import sys, torch
sys.path.insert(0, '.')
import siammask_b200 as smb
import torch.nn.functional as F
dev='cuda'
for (B,C,H,k) in [(192,256,29,5),(96,256,45,5)]:
    x=torch.randn(B,C,H,H,device=dev); kk=torch.randn(B,C,k,k,device=dev)
    out=smb.conv2d_dw_group(x,kk)
    ref=F.conv2d(x[:2].reshape(1,2*C,H,H), kk[:2].reshape(2*C,1,k,k), groups=2*C).view(2,C,H-k+1,H-k+1)
    err=float((out[:2]-ref).abs().max()/ref.abs().max())
    for _ in range(3): smb.conv2d_dw_group(x,kk)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): smb.conv2d_dw_group(x,kk)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/20
    by=B*C*(H*H+k*k+(H-k+1)**2)*4
    print(f"xcorr planes={B*C} {H}x{H}: {ms:.4f} ms  {by/ms/1e6:.0f} GB/s  err {err:.2e}", flush=True)
