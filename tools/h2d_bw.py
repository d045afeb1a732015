"""Dev tool (GPU): pinned-host -> device copy bandwidth of this box (the floor of the e2e leg: 391 MB per step)."""
import torch, time
x = torch.empty(376 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f'H2D {x.numel() * 4 / 1e6:.0f} MB in {ms:.3f} ms = {x.numel() * 4 / ms / 1e6:.1f} GB/s')
