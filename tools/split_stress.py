"""Stress of concurrent sub-batch execution (GPU box): PointPath.forward_device_split against the batched call, many times,
with a description of WHAT differs when something does (frame, voxels, columns). A race anywhere in the path (asynchronous copy
rings, cross-proxy hand-backs, stream joins) shows up here first: small L2-resident frames on three concurrent streams.
Usage: python tools/split_stress.py [repeats] [n_split] [gemm_mode]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth, _lib
from mvxnet_makise_b200.pipeline import PointPath
from mvxnet_makise_b200.modules import pack_calib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n_split = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
_lib.set_gemm_mode(mode)
G = synth.KITTI_GRID
sd = synth.make_weights(9)
calib = synth.kitti_calib()
frames = [synth.make_points(140 + f, P) for f, P in enumerate((800, 1200, 500, 1500, 950))]
B = len(frames)
offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in frames])]).tolist()
pts = torch.from_numpy(np.concatenate(frames, 0)).cuda()
c32 = torch.stack([pack_calib(calib) for _ in range(B)]).cuda()
rng = np.random.default_rng(15)
maps = [torch.from_numpy(rng.standard_normal((B, 256, h, w), dtype=np.float32)).cuda() for h, w in [(13, 42), (7, 21), (4, 11)]]
ref = PointPath(sd, G)
ref.forward_device(pts, offsets, c32, maps, want_grid=False)
torch.cuda.synchronize()
feats = [ref.voxel_features(f)[0].clone() for f in range(B)]
# the batched call against itself first: is the single-stream path deterministic to rounding?
bad_single = 0
for it in range(reps):
    ref.forward_device(pts, offsets, c32, maps, want_grid=False)
    torch.cuda.synchronize()
    for f in range(B):
        d = (ref.voxel_features(f)[0] - feats[f]).abs().max().item() / feats[f].abs().max().item()
        if d > 1e-5:
            bad_single += 1
            print(f'single-stream run {it} frame {f}: rel err {d:.3e}')
path = PointPath(sd, G)
bad = 0
want_grid = len(sys.argv) > 4 and sys.argv[4] == 'grid'
if want_grid:
    g_ref, _ = ref.forward_device(pts, offsets, c32, maps)
    torch.cuda.synchronize()
    nz_ref = (g_ref != 0)
for it in range(reps):
    g, _ = path.forward_device_split(pts, offsets, c32, maps, want_grid, n_split)
    torch.cuda.synchronize()
    if want_grid:
        nz = (g != 0)
        if not torch.equal(nz, nz_ref):
            bad += 1
            for f in range(B):
                x = (nz[f] != nz_ref[f])
                if x.any():
                    where = x.nonzero()
                    print(f'split run {it} frame {f}: grid occupancy differs in {where.shape[0]} cells, first {where[:4].tolist()}; '
                          f'values there: split {g[f][x][:4].tolist()} ref {g_ref[f][x][:4].tolist()}')
    for f in range(B):
        vf = path.voxel_features(f)[0]
        diff = (vf - feats[f]).abs()
        d = diff.max().item() / feats[f].abs().max().item()
        if d > 1e-5:
            bad += 1
            rows = (diff.max(dim=1).values > 1e-5 * feats[f].abs().max()).nonzero().flatten()
            cols = (diff.max(dim=0).values > 1e-5 * feats[f].abs().max()).nonzero().flatten()
            print(f'split run {it} frame {f}: rel err {d:.3e}; {rows.numel()} of {vf.shape[0]} voxels (first {rows[:6].tolist()}), {cols.numel()} of 128 columns (first {cols[:8].tolist()})')
print(f'split stress: {reps} runs, n_split {n_split}, mode {mode}: {bad_single} single-stream and {bad} split frame results deviate')
sys.exit(1 if (bad or bad_single) else 0)
