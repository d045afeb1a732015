"""Dev tool (GPU): pinned-host -> device copy bandwidth of this box (the floor of the e2e leg: 391 MB per step).
Alone: `python tools/h2d_bw.py`. All GPUs at once (what the 8-GPU e2e leg does every step):
`python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_bw.py` - every rank copies concurrently
between two barriers and rank 0 prints the per-GPU and the aggregate rate."""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
x = torch.empty(376 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gbs = x.numel() * 4 / ms / 1e6
if world > 1:
    t = torch.tensor([gbs], device='cuda', dtype=torch.float64)
    g = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(g, t)
    if local == 0:
        r = [round(float(v.item()), 1) for v in g]
        print(f'H2D {x.numel() * 4 / 1e6:.0f} MB per copy, {world} GPUs at once: per GPU {r} GB/s, aggregate {sum(r):.1f} GB/s')
    dist.destroy_process_group()
else:
    print(f'H2D {x.numel() * 4 / 1e6:.0f} MB in {ms:.3f} ms = {gbs:.1f} GB/s')
