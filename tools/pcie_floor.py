import torch, time
dev=torch.device('cuda:0')
n=2147483648
d=torch.empty(n,dtype=torch.uint8,device=dev); h=torch.empty(n,dtype=torch.uint8).pin_memory()
d2=torch.empty(910446000,dtype=torch.uint8,device=dev); h2=torch.empty(910446000,dtype=torch.uint8).pin_memory()
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
for name,fn in (("D2H only",lambda: h.copy_(d,non_blocking=True)),("H2D only",lambda: d2.copy_(h2,non_blocking=True))):
    for _ in range(2):
        torch.cuda.synchronize(); t=time.perf_counter(); fn(); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, dt*1e3,'ms')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.cuda.stream(s1): h.copy_(d,non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
print("both directions concurrently", dt*1e3,'ms ->', 2.147483648/dt,'GB/s decoded-equivalent')
