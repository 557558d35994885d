import torch, time
a=torch.empty(256<<20,dtype=torch.uint8).pin_memory(); b=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
c=torch.empty(64<<20,dtype=torch.uint8).pin_memory(); d=torch.empty(64<<20,dtype=torch.uint8,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for _ in range(2): b.copy_(a,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for _ in range(10): b.copy_(a,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t; print("H2D GB/s", 10*a.numel()/dt/1e9)
t=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): b.copy_(a,non_blocking=True)
    with torch.cuda.stream(s2): c.copy_(d,non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t; print("duplex H2D GB/s", 10*a.numel()/dt/1e9, "D2H", 10*c.numel()/dt/1e9)
