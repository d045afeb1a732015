"""Diagnostic (GPU box): per-layer error of the fused path vs the oracle in fp32 and in fp64 ("truth").
Usage: python tools/diag_layers.py [path_a|path_b]"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth
from mvxnet_makise_b200.pipeline import PointPath
from oracle import pointpath_oracle as O

G = synth.KITTI_GRID
tag = sys.argv[1] if len(sys.argv) > 1 else 'path_a'
g = np.load(os.path.join(ROOT, 'tests', 'golden', tag + '.npz'))
rng = np.random.default_rng(int(g['map_seed']))
maps = [rng.standard_normal((1, 256, h, w), dtype=np.float32) for (h, w) in ((13, 42), (7, 21), (4, 11))]
sd_np = synth.make_weights(int(g['weight_seed']))
calib = synth.kitti_calib()


def chain(dtype):
    sd = {k: torch.from_numpy(v).to(dtype) for k, v in sd_np.items()}
    pcd6 = O.points_with_proj(g['pcd4'], calib)
    vox9, uidx = O.group(pcd6, G.velorange, G.voxelsize, G.T)
    voxels = torch.Tensor(vox9)
    im768 = O.feature_mapping(voxels, [torch.from_numpy(m) for m in maps], torch.Tensor(list(synth.KITTI_IMSIZE_HW)))
    outs = {}
    x = im768[None].to(dtype)
    names = ['head.fusion.fcn1.fc', 'head.fusion.conv1.conv', 'head.fusion.fcn2.fc', 'head.fusion.conv2.conv', 'head.fusion.fcn3.fc']
    for i, n in enumerate(names):
        x = O.crb(x, sd[n + '.weight'], sd[n + '.bias'])
        outs[f'L{i+1}'] = x[0]
    x23 = torch.concat([voxels[None][..., :7].to(dtype), x], dim=-1)
    x = O.vfe(x23, sd['backbone.svfe.vfe1.fcn.fc.weight'], sd['backbone.svfe.vfe1.fcn.fc.bias'])
    outs['V1'] = x[0]
    x = O.vfe(x, sd['backbone.svfe.vfe2.fcn.fc.weight'], sd['backbone.svfe.vfe2.fcn.fc.bias'])
    outs['V2'] = x[0]
    x = O.crb(x, sd['backbone.fcn.fc.weight'], sd['backbone.fcn.fc.bias'])
    outs['vfeat'] = torch.max(x, dim=2)[0][0]
    return voxels, outs


with torch.no_grad():
    voxels, r32 = chain(torch.float32)
    _, r64 = chain(torch.float64)
path = PointPath(sd_np, G)
path([g['pcd4']], [calib], [torch.from_numpy(m) for m in maps], want_grid=False)
torch.cuda.synchronize()
c = path.counts.cpu().numpy()[0]
N, K = int(c[0]), int(c[1])
cap = path.cap
cnt = path.region('vox_cnt', torch.int32, (1, cap))[0, :N].cpu().numpy()
dense_rows = np.concatenate([v * G.T + np.arange(k) for v, k in enumerate(cnt)])
stats = path.region('stats', torch.float64, (8, 1, 768, 2)).cpu()
R = N * G.T


def norm(y, layer, C):
    s = stats[layer, 0].reshape(-1)[:C * 2].reshape(C, 2)
    mean = s[:, 0] / R
    var = (s[:, 1] / R - mean * mean).clamp_min(0)
    return ((y.double() - mean) / torch.sqrt(var + 1e-6))


def err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


couts = [768, 128, 128, 16, 16]
for i, C in enumerate(couts):
    y = path.region(f'Y{i+1}', torch.float32, (1, cap + 128, C))[0, :K].cpu()
    gpu = norm(y, i, C)
    ref32 = r32[f'L{i+1}'].reshape(-1, C)[dense_rows]
    ref64 = r64[f'L{i+1}'].reshape(-1, C)[dense_rows]
    d = (gpu - ref64).abs().max(0)[0]
    worst = int(d.argmax())
    s = stats[i, 0].reshape(-1)[:C * 2].reshape(C, 2)
    mean = s[:, 0] / R; var = s[:, 1] / R - mean * mean
    print(f'L{i+1}: gpu-vs-ref32 {err(gpu, ref32):.3e}  gpu-vs-f64 {err(gpu, ref64):.3e}  ref32-vs-f64 {err(ref32, ref64):.3e}'
          f'  worst ch {worst}: var {var[worst]:.3e} mean {mean[worst]:.3e} absdiff {d[worst]:.3e} max|ref| {ref64.abs().max():.3f}')
vf, _ = path.voxel_features(0)
print(f'vfeat: gpu-vs-ref32 {err(vf.cpu(), r32["vfeat"]):.3e}  gpu-vs-f64 {err(vf.cpu(), r64["vfeat"]):.3e}  ref32-vs-f64 {err(r32["vfeat"], r64["vfeat"]):.3e}  max|ref| {r64["vfeat"].abs().max():.3f}')
print(f'golden(ref impl) vs f64: {err(torch.from_numpy(g["vfeat"]), r64["vfeat"]):.3e}')
