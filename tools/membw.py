"""GPU microbenchmark: write-only and copy bandwidth references for the grid-fill roofline."""
import torch, time
x = torch.empty(8 * 128 * 10 * 352 * 400, dtype=torch.float32, device='cuda')
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gb = x.numel() * 4 / 1e9
ms = t(lambda: x.zero_()); print(f'zero_ (memset kernel) {gb:.2f} GB: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s write-only')
ms = t(lambda: x.fill_(1.5)); print(f'fill_: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s write-only')
ms = t(lambda: torch.cuda.memset if False else x.zero_()); 
ms = t(lambda: y.copy_(x)); print(f'copy_: {ms:.3f} ms  {2*gb/ms*1e3:.0f} GB/s read+write')
ms = t(lambda: x.sum()); print(f'sum (read-only): {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s read-only')
