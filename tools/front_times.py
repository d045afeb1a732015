"""Serialised CUDA-event times of the front stages (GPU box): operand pack, per-pixel GEMM, voxelisation, clears + row sort.
The map branch runs behind the point branch here (fusion mode 2), so every time is the stage's own. Usage: python tools/front_times.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvxnet_makise_b200 import synth, _lib
from mvxnet_makise_b200.pipeline import PointPath
from mvxnet_makise_b200.modules import pack_calib
B, P = 8, 120000
dev = torch.device('cuda')
frames = [synth.make_points(f, P) for f in range(B)]
offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in frames])]).tolist()
pts = torch.from_numpy(np.concatenate(frames, 0)).to(dev)
calib = torch.stack([pack_calib(synth.kitti_calib()) for _ in range(B)]).to(dev)
g = torch.Generator().manual_seed(1234)
maps = [torch.randn((B, 256, h, w), generator=g).to(dev) for (h, w) in synth.fpn_shapes()]
path = PointPath(synth.make_weights(0), synth.KITTI_GRID, device=dev)
_lib.set_fusion_mode(2)
for _ in range(3): path.forward_device(pts, offsets, calib, maps)
torch.cuda.synchronize()
steps = 10
_lib.check(_lib.lib.mvx_timing_enable(steps), 'te')
for _ in range(steps): path.forward_device(pts, offsets, calib, maps)
torch.cuda.synchronize()
seg = np.zeros(_lib.NUM_SEGMENTS); buf = (_lib.c_float * _lib.NUM_SEGMENTS)()
for c in range(steps):
    _lib.check(_lib.lib.mvx_timing_read(c, buf), 'tr'); seg += np.array(buf[:]) / steps
names = [(_lib.lib.mvx_timing_segment_name(i) or b'').decode() for i in range(_lib.NUM_SEGMENTS)]
d = {n: round(float(v), 4) for n, v in zip(names, seg) if n}
print('serialised front (ms per 8 frames): maps_nhwc', d['maps_nhwc'], 'pixel_gemm', d['pixel_gemm'], 'voxelize', d['voxelize'], 'rows_build', d['rows_build'], 'clear', d['clear'])
