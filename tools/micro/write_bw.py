import torch
x=torch.empty(850_000_000, dtype=torch.float64, device='cuda')
def t(fn,reps=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    best=1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); best=min(best,e0.elapsed_time(e1))
    return best
ms=t(lambda: x.zero_()); print('zero_ 6.8GB: %.3f ms %.0f GB/s'%(ms, x.numel()*8/ms/1e6))
ms=t(lambda: x.fill_(1.5)); print('fill_ 6.8GB: %.3f ms %.0f GB/s'%(ms, x.numel()*8/ms/1e6))
y=torch.empty_like(x[:425_000_000]); z=x[:425_000_000]
ms=t(lambda: y.copy_(z)); print('copy 3.4GB->3.4GB: %.3f ms %.0f GB/s'%(ms, 2*z.numel()*8/ms/1e6))
ms=t(lambda: torch.sum(x)); print('sum read 6.8GB: %.3f ms %.0f GB/s'%(ms, x.numel()*8/ms/1e6))
