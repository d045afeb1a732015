"""Print the CUDA-event stage times of the fused path for a given kernel configuration (GPU box).
Usage: python tools/stage_times.py [gemm_mode] ; env MVX_DBG for experiments"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth, _lib
from mvxnet_makise_b200.pipeline import PointPath
from mvxnet_makise_b200.modules import pack_calib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, P, steps = 8, 120000, 10
dev = torch.device('cuda')
frames = [synth.make_points(f, P) for f in range(B)]
offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in frames])]).tolist()
pts = torch.from_numpy(np.concatenate(frames, 0)).to(dev)
calib = torch.stack([pack_calib(synth.kitti_calib()) for _ in range(B)]).to(dev)
g = torch.Generator().manual_seed(1234)
maps = [torch.randn((B, 256, h, w), generator=g).to(dev) for (h, w) in synth.fpn_shapes()]
path = PointPath(synth.make_weights(0), synth.KITTI_GRID, device=dev)
_lib.set_gemm_mode(mode)
for _ in range(3): path.forward_device(pts, offsets, calib, maps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): path.forward_device(pts, offsets, calib, maps)
e1.record(); torch.cuda.synchronize()
print(f'mode={mode} step (overlapped schedule, 30 steps): {e0.elapsed_time(e1) / 30:.3f} ms = {B * 30 / e0.elapsed_time(e1) * 1e3:.0f} frames/s')
_lib.check(_lib.lib.mvx_timing_enable(steps), 'te')
for _ in range(steps): path.forward_device(pts, offsets, calib, maps)
torch.cuda.synchronize()
seg = np.zeros(_lib.NUM_SEGMENTS); buf = (_lib.c_float * _lib.NUM_SEGMENTS)()
for c in range(steps):
    _lib.check(_lib.lib.mvx_timing_read(c, buf), 'tr'); seg += np.array(buf[:]) / steps
names = [(_lib.lib.mvx_timing_segment_name(i) or b'').decode() for i in range(_lib.NUM_SEGMENTS)]
print(f'mode={mode} total={seg.sum():.3f} ms', {n: round(float(v), 3) for n, v in zip(names, seg) if n and v > 0.04})
