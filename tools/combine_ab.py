"""A/B of the two combine kernels on the bench inputs (GPU box): the run-structured kernel (default) against the first, row-by-row
version (mvx_set_fold_mode(2)). Y1 is bit-identical by construction; the BatchNorm sums group their fp32 partial sums differently,
so the voxel features agree to rounding. Repeated runs look for sporadic differences (a race in the cp.async ring would show here).
Usage: python tools/combine_ab.py [repeats] [gemm_mode]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth, _lib
from mvxnet_makise_b200.pipeline import PointPath
from mvxnet_makise_b200.modules import pack_calib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, P = 8, 120000
dev = torch.device('cuda')
frames = [synth.make_points(f, P) for f in range(B)]
offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in frames])]).tolist()
pts = torch.from_numpy(np.concatenate(frames, 0)).to(dev)
calib = torch.stack([pack_calib(synth.kitti_calib()) for _ in range(B)]).to(dev)
g = torch.Generator().manual_seed(1234)
maps = [torch.randn((B, 256, h, w), generator=g).to(dev) for (h, w) in synth.fpn_shapes()]
path = PointPath(synth.make_weights(0), synth.KITTI_GRID, device=dev)
_lib.set_gemm_mode(mode)


def run(fold_mode):
    _lib.set_fold_mode(fold_mode)
    path.forward_device(pts, offsets, calib, maps, want_grid=False)
    torch.cuda.synchronize()
    return [path.voxel_features(f)[0].clone() for f in range(B)]


ref = run(2)
worst = 0.0
for it in range(reps):
    got = run(0)
    for f in range(B):
        err = (got[f].double() - ref[f].double()).abs().max().item() / ref[f].abs().max().item()
        worst = max(worst, err)
        if err > 1e-5:
            print(f'iteration {it} frame {f}: rel err {err:.3e}')
_lib.set_fold_mode(0)
print(f'combine A/B: {reps} runs x {B} frames, worst rel err of the voxel features {worst:.3e}')
sys.exit(0 if worst <= 1e-5 or mode == 6 else 1)
