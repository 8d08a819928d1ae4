import torch
x = torch.empty(1<<28, dtype=torch.int32, device='cuda')  # 1 GiB
y = torch.empty(1<<28, dtype=torch.int32, device='cuda')
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
ms=t(lambda: x.fill_(7)); print(f"fill_ 1 GiB: {ms:.4f} ms = {(1<<30)/ms/1e6:.0f} GB/s write")
ms=t(lambda: x.zero_()); print(f"zero_ 1 GiB: {ms:.4f} ms = {(1<<30)/ms/1e6:.0f} GB/s write")
ms=t(lambda: y.copy_(x)); print(f"copy 1 GiB: {ms:.4f} ms = {2*(1<<30)/ms/1e6:.0f} GB/s r+w")
ms=t(lambda: x.sum()); print(f"sum (read) 1 GiB: {ms:.4f} ms = {(1<<30)/ms/1e6:.0f} GB/s read")
