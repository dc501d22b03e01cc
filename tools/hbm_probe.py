import torch, time
x = torch.empty(1<<30, dtype=torch.float64, device='cuda')
y = torch.empty(1<<30, dtype=torch.float64, device='cuda')
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a,b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
ms = t(lambda: x.fill_(1.5)); print("fill 8GB: %.3f ms -> %.0f GB/s write" % (ms, 8.59e9/ms/1e6))
ms = t(lambda: y.copy_(x)); print("copy 8GB: %.3f ms -> %.0f GB/s r+w" % (ms, 2*8.59e9/ms/1e6))
ms = t(lambda: x.sum()); print("sum 8GB: %.3f ms -> %.0f GB/s read" % (ms, 8.59e9/ms/1e6))
