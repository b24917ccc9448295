import torch
dev=torch.device("cuda",0)
B,C=8,64
buf=torch.empty(B,C,16,200,200,device=dev)
src=torch.randn_like(buf)
def timeit(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
gb=buf.numel()*4/1e9
for v in (0.0,1.2345):
    ms=timeit(lambda: buf.fill_(v)); print("fill",v,ms*1e3,"us",gb/ms*1e3,"GB/s")
# write of non-constant data with negligible read: expand a small tensor
small=torch.randn(1,1,16,200,200,device=dev)
ms=timeit(lambda: buf.copy_(small.expand_as(buf))); print("bcast copy (2.5MB src)",ms*1e3,"us",gb/ms*1e3,"GB/s")
ms=timeit(lambda: buf.copy_(src)); print("copy",ms*1e3,"us",2*gb/ms*1e3,"GB/s r+w")
