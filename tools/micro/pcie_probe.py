import torch, time
n=411525120
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(200540160,dtype=torch.uint8).pin_memory(); d2=torch.empty(200540160,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for rep in range(2):
    torch.cuda.synchronize(); t=time.time()
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); a=time.time()-t
    torch.cuda.synchronize(); t=time.time()
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    torch.cuda.synchronize(); b=time.time()-t
    torch.cuda.synchronize(); t=time.time()
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    torch.cuda.synchronize(); c=time.time()-t
    print(f"H2D 411MB {a*1e3:.2f} ms ({n/a/1e9:.1f} GB/s)  D2H 200MB {b*1e3:.2f} ms ({200540160/b/1e9:.1f} GB/s)  both {c*1e3:.2f} ms")
