import torch, numpy as np
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for mb in (56, 125, 500, 2000):
    n=mb*1000*1000//8
    x=torch.randn(n,dtype=torch.float64,device='cuda'); y=torch.empty_like(x)
    ms=t(lambda: y.copy_(x)); print(f'{mb} MB copy (read+write {2*mb} MB): {ms*1e3:.1f} us  {2*mb/ms/1e3:.2f} TB/s')
    ms=t(lambda: x.mul_(1.0000001)); print(f'{mb} MB in-place mul: {ms*1e3:.1f} us  {2*mb/ms/1e3:.2f} TB/s')
