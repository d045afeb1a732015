"""Dev tool (GPU): H2D time of the e2e leg's copy pattern alone (per-sub-batch slices of the pinned batch tensors)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvxnet_makise_b200 import synth
B = 8
maps_h = [torch.randn((B, 256, h, w)).pin_memory() for (h, w) in synth.fpn_shapes()]
pts_h = torch.randn((B * 120000, 4)).pin_memory()
for sizes in ([8], [2, 2, 2, 2], [2, 2, 2, 1, 1], [1] * 8):
    bounds = [0]
    for c in sizes: bounds.append(bounds[-1] + c)
    dm = [[torch.empty((b1 - b0,) + tuple(m.shape[1:]), device='cuda') for m in maps_h] for b0, b1 in zip(bounds[:-1], bounds[1:])]
    dp = [torch.empty(((b1 - b0) * 120000, 4), device='cuda') for b0, b1 in zip(bounds[:-1], bounds[1:])]
    def run():
        for i, (b0, b1) in enumerate(zip(bounds[:-1], bounds[1:])):
            dp[i].copy_(pts_h[b0 * 120000:b1 * 120000], non_blocking=True)
            for d, h in zip(dm[i], maps_h):
                d.copy_(h[b0:b1], non_blocking=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(sizes, f'{e0.elapsed_time(e1) / 10:.3f} ms per step')
