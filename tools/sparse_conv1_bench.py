"""Dev tool (GPU): cost of the sparse CML.conv1 hand-off at the headline size, next to the dense route
(grid fill + torch/cuDNN Conv3d + BatchNorm3d on the dense grid, what the reference does after the path)."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvxnet_makise_b200 import synth
from mvxnet_makise_b200.pipeline import PointPath
from mvxnet_makise_b200.modules import pack_calib

B, P = 8, 120_000
dev = torch.device('cuda')
frames = [synth.make_points(f, P) for f in range(B)]
offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in frames])]).tolist()
points = torch.from_numpy(np.concatenate(frames, 0)).to(dev)
calib = torch.stack([pack_calib(synth.kitti_calib()) for _ in range(B)]).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
maps = [torch.randn((B, 256, h, w), generator=g, device=dev) for (h, w) in synth.fpn_shapes()]
w = torch.randn((64, 128, 3, 3, 3), generator=g, device=dev) / (128 * 27) ** 0.5
b = torch.randn(64, generator=g, device=dev) * 0.1
path = PointPath(synth.make_weights(0), synth.KITTI_GRID, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_path_grid = timed(lambda: path.forward_device(points, offsets, calib, maps, True))
t_path_nogrid = timed(lambda: path.forward_device(points, offsets, calib, maps, False))
t_sparse = timed(lambda: (path.forward_device(points, offsets, calib, maps, False), path.cml_conv1(w, b)))
grid, _ = path.forward_device(points, offsets, calib, maps, True)


def dense_conv1():
    outs = []
    for f in range(B):                                   # the reference runs CML per frame (batch 1)
        y = F.relu(F.conv3d(grid[f:f + 1], w, b, stride=(2, 1, 1), padding=(1, 1, 1)))
        outs.append(F.batch_norm(y, None, None, None, None, True, 0.0, 1e-6))
    return outs


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
t_dense = timed(dense_conv1, 3)
out = path.cml_conv1(w, b)
ref = torch.cat(dense_conv1())
err = ((out - ref).abs().max() / ref.abs().max()).item()
print(f'path with dense grid {t_path_grid:.3f} ms | path without grid {t_path_nogrid:.3f} ms | path + sparse conv1 {t_sparse:.3f} ms '
      f'(sparse conv1 alone {t_sparse - t_path_nogrid:.3f} ms) | dense route: grid fill {t_path_grid - t_path_nogrid:.3f} ms + torch Conv3d/BN3d fp32 '
      f'{t_dense:.3f} ms | sparse vs torch dense rel diff {err:.2e}  (batch {B})')
