import torch,time
dev=torch.device('cuda')
n=int(6e9)//8
x=torch.empty(n,dtype=torch.float64,device=dev)
for name,fn in (('fill_',lambda: x.fill_(1.5)),('zero_',lambda: x.zero_())):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print(name, ms, 'ms', n*8/ms/1e6, 'GB/s')
y=torch.empty_like(x)
for _ in range(3): y.copy_(x)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('copy', ms, 'ms', 2*n*8/ms/1e6, 'GB/s')
# 14 separate write streams like kernel A (11 f64 + i32 + u8): torch multi-tensor fill
m=69_000_000
ts=[torch.empty(m,dtype=torch.float64,device=dev) for _ in range(10)]+[torch.empty(m,dtype=torch.int64,device=dev),torch.empty(m,dtype=torch.int32,device=dev),torch.empty(m,dtype=torch.uint8,device=dev)]
for _ in range(3):
    for t in ts: t.fill_(1)
e0.record()
for _ in range(10):
    for t in ts: t.fill_(1)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('13 fills', ms,'ms', sum(t.numel()*t.element_size() for t in ts)/ms/1e6,'GB/s')
