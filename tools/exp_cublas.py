import torch, time
dev='cuda'
def bench(M,N,K,dt=torch.float16,n=30):
    a=torch.randn(M,K,device=dev,dtype=dt); b=torch.randn(N,K,device=dev,dtype=dt)
    for _ in range(5): c=a@b.t()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): c=a@b.t()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/n
    print(f"cuBLAS {str(dt):14s} M={M} N={N} K={K}: {ms:.4f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
for (M,N,K) in [(61504,1024,4608),(61504,1024,4864),(61504,256,2304),(61504,1024,256),(61504,256,1024),(254016,64,576),(8192,8192,8192)]:
    bench(M,N,K)
bench(61504,1024,4608,torch.bfloat16)
# 3 back-to-back GEMMs as a stand-in for a 3-pass split product
a=torch.randn(61504,4608,device=dev,dtype=torch.float16); b=torch.randn(1024,4608,device=dev,dtype=torch.float16)
torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for _ in range(3): c=a@b.t()
torch.cuda.synchronize(); e0.record()
for _ in range(30): c=a@b.t()
e1.record(); torch.cuda.synchronize(); print('sustained 30x big:', e0.elapsed_time(e1)/30)
