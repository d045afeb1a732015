"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + torch-CPU fp32, the reference's own arithmetic libraries) of the
reference's point-side hot path. Nothing under ``mvxnet_makise_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and only as the checker or as the timed CPU baseline.

Parity pinning: the reference ships no tests or golden vectors. Every function
here is pinned against the reference's own code executed in the build container
(``tests/golden/make_golden.py`` imports /root/reference through ``oracle/refshim.py`` and writes
``tests/golden/*.npz``; ``tests/test_oracle.py`` re-checks the restatement against those fixtures
on every run, and against the live reference where /root/reference exists).

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle() -> str:
    """gcc-compile oracle/voxel_oracle.c (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(['make', '-s', '-C', _HERE, 'all'], check=True)
    return os.path.join(_HERE, '_build', 'libvoxel_oracle.so')


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libvoxel_oracle.so')
        if not os.path.exists(path):
            build_c_oracle()
        lib = ctypes.CDLL(path)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
        lib.oracle_cell_index.argtypes = [vp, i64, i64, vp, vp, vp]
        lib.oracle_cell_index.restype = None
        lib.oracle_group_assign.argtypes = [vp, i64, i32, vp, vp, vp, vp]
        lib.oracle_group_assign.restype = i64
        lib.oracle_emit_group7.argtypes = [vp, i64, i64, i32, vp, vp, i64, vp]
        lib.oracle_emit_group7.restype = None
        lib.oracle_emit_group9.argtypes = [vp, i64, i64, i32, vp, vp, i64, vp, vp]
        lib.oracle_emit_group9.restype = None
        _LIB = lib
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- stage 1: voxelization
def cell_index(pcd: np.ndarray, velorange: Sequence[float], voxelsize: Sequence[float]) -> np.ndarray:
    """``((pts - low) / size).astype('int32')`` in fp64 — Preprocessing.py:67-69 / :86-90."""
    pcd = np.ascontiguousarray(pcd, dtype=np.float32)
    low = np.array(velorange[0:3], dtype=np.float64)
    size = np.array(voxelsize, dtype=np.float64)
    idx = np.empty((pcd.shape[0], 3), dtype=np.int32)
    _lib().oracle_cell_index(_p(pcd), pcd.shape[0], pcd.shape[1], _p(low), _p(size), _p(idx))
    return idx


def cell_index_numpy(pcd, velorange, voxelsize):
    """The literal numpy expression of Preprocessing.py:67-69 (cross-check of the C loop)."""
    pts = np.asarray(pcd, dtype=np.float32)[:, :3]
    low = np.array(velorange[0:3])
    return ((pts - low) / np.array(voxelsize)).astype('int32')


def group_assign(idx: np.ndarray, T: int):
    """Grouping pass of cpp/voxelutil.cpp:331-342 == Preprocessing.py:94-104.
    Returns (vid_of_point int64[P], slot_of_point int32[P], coords int64[V,3], cnt int64[V])."""
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    P = idx.shape[0]
    vid = np.empty(P, dtype=np.int64)
    slot = np.empty(P, dtype=np.int32)
    coords = np.empty((max(P, 1), 3), dtype=np.int64)
    cnt = np.empty(max(P, 1), dtype=np.int64)
    V = _lib().oracle_group_assign(_p(idx), P, T, _p(vid), _p(slot), _p(coords), _p(cnt))
    return vid, slot, coords[:V].copy(), cnt[:V].copy()


def cpp_group(pcd: np.ndarray, idx: np.ndarray, T: int):
    """`_group(pcd, idx, T)` — cpp/voxelutil.cpp:325-360. Returns (voxel (V,T,7) f32, (x,y,z) int64, cnt int64)."""
    pcd = np.ascontiguousarray(pcd, dtype=np.float32)      # pybind array_t<float> forcecast
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    vid, slot, coords, cnt = group_assign(idx, T)
    V = coords.shape[0]
    voxel = np.empty((V, T, 7), dtype=np.float32)
    _lib().oracle_emit_group7(_p(pcd), pcd.shape[0], pcd.shape[1], T, _p(vid), _p(slot), V, _p(voxel))
    return voxel, (coords[:, 0].copy(), coords[:, 1].copy(), coords[:, 2].copy()), cnt


def group_(pcd: np.ndarray, velorange, voxelsize, T: int):
    """`group_` without its in-place shuffle — Preprocessing.py:57-73 (fp32 centroid variant)."""
    idx = cell_index(pcd, velorange, voxelsize)
    voxel, uidx, vcnt = cpp_group(pcd, idx, T)
    center = voxel[..., :3].sum(axis=1) / vcnt[:, None]
    voxel[..., 3:6] = voxel[..., :3] - center[:, None, :]
    return voxel, np.array(uidx).T


def group(pcd6: np.ndarray, velorange, voxelsize, T: int):
    """numba `group` without its in-place shuffle — Preprocessing.py:75-116 (the one train.py:44 calls).
    pcd6 = (P,6) [x,y,z,r,row,col]. Returns (voxel (V,T,9) fp64, uidx (V,3) fp64)."""
    pcd6 = np.ascontiguousarray(pcd6, dtype=np.float32)
    idx = cell_index(pcd6, velorange, voxelsize)
    vid, slot, coords, cnt = group_assign(idx, T)
    V = coords.shape[0]
    voxel = np.empty((V, T, 9), dtype=np.float64)
    _lib().oracle_emit_group9(_p(pcd6), pcd6.shape[0], pcd6.shape[1], T, _p(vid), _p(slot), V, _p(cnt), _p(voxel))
    return voxel, coords.astype(np.float64)


# --------------------------------------------------------------------------- before the path: crop / cropToSight
def crop(pcd: np.ndarray, velorange: Sequence[float]) -> np.ndarray:
    """`crop(pcd, range)` — modules/data/Preprocessing.py:12-17 (fp32 coordinates compared against fp64 bounds)."""
    low = np.array(velorange[0:3])
    high = np.array(velorange[3:6])
    roi = pcd[:, :3]
    return pcd[np.all((low <= roi) & (roi < high), axis=1)]


def crop_to_sight(pcd: np.ndarray, calib: Dict[str, np.ndarray], imsize_wh: Sequence[int]) -> np.ndarray:
    """numpy branch of `cropToSight(pcd, calib, imsize)` — modules/data/Preprocessing.py:26-55 (imsize is (w, h))."""
    imsize = np.array(imsize_wh) - 1e-3
    points = np.empty((4, pcd.shape[0]), dtype='float32')
    points[:3] = pcd.T[:3]
    points[3] = 1
    points = calib['R0_rect'] @ calib['Tr_velo_to_cam'] @ points
    f = points[2] > 0
    pcd = pcd[f]
    points = points[:, f]
    points = calib['P2'] @ points
    points[:2] = points[:2] / points[2]
    points = points[:2].T
    f = np.all(points >= 0, axis=1) & np.all(points < imsize, axis=1)
    return pcd[f]


# --------------------------------------------------------------------------- stage 2a: projection
def lidar2img(pcd: torch.Tensor, calib: Dict[str, torch.Tensor]) -> torch.Tensor:
    """`lidar2Img(pcd, calib, uncheck=True)` — modules/utils/Calib.py:47-70. Returns (P,2) (u,v)."""
    pts = torch.empty((4, pcd.shape[0]))
    pts[:3] = pcd[:, :3].T
    pts[3] = 1
    pts = calib['R0_rect'] @ calib['Tr_velo_to_cam'] @ pts
    pts = calib['P2'] @ pts
    pts[:2] = pts[:2] / pts[2]
    return pts[:2].T


def points_with_proj(pcd4: np.ndarray, calib_np: Dict[str, np.ndarray]) -> np.ndarray:
    """train.py:31-35: pcd -> torch, proj = lidar2Img(...)[:, [1,0]], concat -> (P,6) numpy fp32."""
    pcd = torch.Tensor(np.asarray(pcd4, dtype=np.float32))
    calib = {k: torch.Tensor(np.asarray(v)) for k, v in calib_np.items()}
    proj = lidar2img(pcd, calib)[:, [1, 0]]
    return torch.concat([pcd, proj], dim=1).numpy()


def lidar2img_numpy(pcd: np.ndarray, calib: Dict[str, np.ndarray]) -> np.ndarray:
    """numpy branch of `lidar2Img(pcd, calib, uncheck=True)` - modules/utils/Calib.py:56-70, literally. The result dtype is
    numpy's promotion of the calibration dtype with the fp32 point buffer: float64 for the float64 dicts `readCalib` returns
    (Load.py:24-41), float32 for fp32 dicts."""
    points = np.empty((4, pcd.shape[0]), dtype='float32')
    points[:3] = pcd[:, :3].T
    points[3] = 1
    points = calib['R0_rect'] @ calib['Tr_velo_to_cam'] @ points
    points = calib['P2'] @ points
    points[:2] = points[:2] / points[2]
    return points[:2].T


def merged_points_with_proj(point_sets, calibs_np) -> np.ndarray:
    """train.py:29-42 (GT-paste): the scene through the TORCH branch of lidar2Img (`torch.Tensor` points and calibration, so
    fp32 whatever the dict held; columns swapped by `[:, [1, 0]]`), then every pasted object through the NUMPY branch
    (`proj[:, ::-1]`) with ITS OWN calibration dict as `readCalib` left it - float64 matrices give a float64 projection -
    concatenated in that order -> (P,6), float64 as soon as one pasted set is. `group` copies the projection into the
    voxel tensor and `torch.Tensor(voxel)` (train.py:125) rounds it to fp32 once."""
    out = [points_with_proj(point_sets[0], calibs_np[0])]
    for p, c in zip(point_sets[1:], calibs_np[1:]):
        p = np.asarray(p, dtype=np.float32)
        out.append(np.concatenate([p, lidar2img_numpy(p, c)[:, ::-1]], axis=1))                       # train.py:37-40
    return np.concatenate(out, axis=0)                                                               # train.py:42


# --------------------------------------------------------------------------- stage 2b: gather
def feature_mapping(voxels: torch.Tensor, features: List[torch.Tensor], imsize: torch.Tensor,
                    eps: float = 1e-6) -> torch.Tensor:
    """`featureMaping` for one frame — modules/imhead/Pipe.py:23-82.
    voxels (N,T,9) fp32 is MUTATED like the reference does (pad slots zeroed); features are the
    un-padded NCHW maps (this function pads a copy, it does not mutate the list). Returns (N,T,768)."""
    feats = [F.pad(f, (0, 1, 0, 1)) for f in features]                     # Pipe.py:47-48
    region = [imsize / torch.Tensor([*f.shape[-2:]]).to(imsize.device) for f in features]    # Pipe.py:41-45
    v = voxels
    origshape = v.shape[:-1]
    xyz = v[..., :3].reshape((-1, 3))
    zero = torch.all(xyz == 0, dim=1)                                      # Pipe.py:53-54
    proj = v[..., -2:].reshape((-1, 2))
    proj[zero] = 0
    zero = zero.reshape(origshape)
    v[zero] = 0                                                            # Pipe.py:58-59
    out = []
    for feature, rs in zip(feats, region):
        q = proj / rs - eps                                                # Pipe.py:62
        index = q.long()
        xi = (q[:, 0] - index[:, 0])[None, :]
        yi = (q[:, 1] - index[:, 1])[None, :]
        xi_, yi_ = 1 - xi, 1 - yi
        x, y = index[:, 0], index[:, 1]
        x1, y1 = x + 1, y + 1
        assert torch.max(x1) < feature.shape[-2] and torch.max(y1) < feature.shape[-1]
        g = feature[0, :, x, y] * xi * yi                                  # Pipe.py:72-75 (inverted weights)
        g = g + feature[0, :, x1, y] * xi_ * yi
        g = g + feature[0, :, x, y1] * xi * yi_
        g = g + feature[0, :, x1, y1] * xi_ * yi_
        out.append(g)
    out = torch.concat(out, dim=0).T
    out = out.reshape(origshape + (out.shape[-1],))
    out[zero] = 0                                                          # Pipe.py:80
    return out


# --------------------------------------------------------------------------- stage 3: layer stack
def crb(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """One FCN / CRB2d block on (1,N,T,Cin): relu(Linear) then BatchNorm2d(affine=False,
    track_running_stats=False) i.e. batch statistics, biased variance — modules/layers/Blocks.py:5-18, 31-40."""
    if w.dim() == 4:        # CRB2d: 1x1 Conv2d on the NCHW view (Blocks.py:31-40; Pipe.py:97-102 permutes)
        y = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w, b))
    else:                   # FCN: nn.Linear on the channels-last view (Blocks.py:13-15)
        y = F.relu(F.linear(x, w, b)).permute(0, 3, 1, 2)
    y = F.batch_norm(y, None, None, None, None, True, 0.0, eps)
    return y.permute(0, 2, 3, 1)


def fusion(x768: torch.Tensor, sd: Dict[str, torch.Tensor], eps: float = 1e-6) -> torch.Tensor:
    """`ImageFeatureFusion.forward` — modules/imhead/Pipe.py:94-104. (1,N,T,768) -> (1,N,T,16)."""
    x = x768
    for name in ('head.fusion.fcn1.fc', 'head.fusion.conv1.conv', 'head.fusion.fcn2.fc',
                 'head.fusion.conv2.conv', 'head.fusion.fcn3.fc'):
        x = crb(x, sd[name + '.weight'], sd[name + '.bias'], eps)
    return x


def vfe(x, w, b, eps=1e-6):
    """`VFE.forward` — modules/voxelnet/Pipe.py:12-18."""
    x = crb(x, w, b, eps)
    s = torch.max(x, dim=2, keepdim=True)[0].repeat(1, 1, x.shape[2], 1)
    return torch.concat([x, s], dim=-1)


def voxel_features(x23: torch.Tensor, sd, eps=1e-6) -> torch.Tensor:
    """SVFE + FCN + max over T — voxelnet/Pipe.py:24-29, VoxelNet.py:26-32. (1,N,T,23) -> (N,128)."""
    x = vfe(x23, sd['backbone.svfe.vfe1.fcn.fc.weight'], sd['backbone.svfe.vfe1.fcn.fc.bias'], eps)
    x = vfe(x, sd['backbone.svfe.vfe2.fcn.fc.weight'], sd['backbone.svfe.vfe2.fcn.fc.bias'], eps)
    x = crb(x, sd['backbone.fcn.fc.weight'], sd['backbone.fcn.fc.bias'], eps)
    x = torch.max(x, dim=2)[0]
    return torch.squeeze(x, dim=2).reshape((-1, x.shape[-1]))


# --------------------------------------------------------------------------- stage 4: scatter
def reindex(x: torch.Tensor, idx: torch.Tensor, voxelshape) -> torch.Tensor:
    """`VoxelNet.reindex` — modules/voxelnet/VoxelNet.py:16-22. idx (N,4) int64 [batch,ix,iy,iz]."""
    res = torch.zeros((1, x.shape[1], voxelshape[2], voxelshape[0], voxelshape[1]), dtype=x.dtype, device=x.device)   # VoxelNet.py:19 (cfg.device)
    res[idx[:, 0], :, idx[:, 3], idx[:, 1], idx[:, 2]] = x
    return res


# --------------------------------------------------------------------------- after the path: CML.conv1
def crb3d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride, padding, eps: float = 1e-6) -> torch.Tensor:
    """`CRB3d(cin, cout, k, s, p).forward` - modules/layers/Blocks.py:20-29: bn(relu(Conv3d(x))), BatchNorm3d(affine=False,
    track_running_stats=False) i.e. batch statistics, biased variance. Pinned on the reference's own CML by tests/golden/cml_a.npz."""
    y = F.relu(F.conv3d(x, w, b, stride=stride, padding=padding))
    return F.batch_norm(y, None, None, None, None, True, 0.0, eps)


def cml(grid: torch.Tensor, ws, bs, eps: float = 1e-6):
    """`CML.forward` - modules/voxelnet/Pipe.py:31-43: conv1 (2,1,1)/(1,1,1), conv2 1/(0,1,1), conv3 (2,1,1)/1. Returns all three outputs."""
    y1 = crb3d(grid, ws[0], bs[0], (2, 1, 1), (1, 1, 1), eps)
    y2 = crb3d(y1, ws[1], bs[1], 1, (0, 1, 1), eps)
    y3 = crb3d(y2, ws[2], bs[2], (2, 1, 1), 1, eps)
    return y1, y2, y3


def cml_conv1(grid: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """`CML.conv1` = CRB3d(128, 64, 3, (2,1,1), (1,1,1)) — modules/voxelnet/Pipe.py:31-43; CRB3d (modules/layers/Blocks.py):
    bn(relu(Conv3d(x))) with BatchNorm3d(affine=False, track_running_stats=False) i.e. batch statistics, biased variance.
    grid (1,128,nz,nx,ny) -> (1,64,(nz+1)//2,nx,ny)."""
    y = F.relu(F.conv3d(grid, w, b, stride=(2, 1, 1), padding=(1, 1, 1)))
    return F.batch_norm(y, None, None, None, None, True, 0.0, eps)


# --------------------------------------------------------------------------- whole path, one frame
def forward_frame(pcd4: np.ndarray, calib_np, fpn_maps: List[np.ndarray], sd_np, grid, imsize_hw,
                  eps: float = 1e-6, want_grid: bool = True, stages: dict | None = None,
                  dtype: torch.dtype = torch.float32, device: str = 'cpu'):
    """The reference's chain for ONE frame, shuffle disabled (trap 3):
    lidar2Img -> group -> [host glue train.py:118-128] -> featureMaping -> fusion -> concat
    (MVXNet.py:25-26) -> SVFE/FCN/max -> reindex. Returns dict of intermediate results.

    dtype=torch.float64 evaluates the SAME layer stack (stage 3) in double precision on the fp32 gather
    output: the rounding-free value of the reference algorithm, used by the tests to separate the fp32
    reference's own rounding noise (BatchNorm with 86 % identical pad rows normalises real rows to tens of
    sigma, so fp32 statistics noise is amplified) from errors of the implementation under test.

    device='cuda' evaluates the SAME torch expressions (stages 2b-4) with torch's CUDA kernels instead of its CPU
    kernels - still the checker, never the product: the full-size parity tests use it to get the fp64 value of all 8
    frames of the benchmarked batch in seconds, after pinning it on one frame against the CPU evaluation (fp64 GEMMs and
    reductions agree to ~1e-12). Projection and voxelization (stages 1-2a) always run on the CPU."""
    import time
    t = {}
    sd = {k: torch.from_numpy(np.asarray(v)).to(device=device, dtype=dtype) for k, v in sd_np.items()}
    t0 = time.perf_counter()
    # a list of calibrations: pcd4 is the matching list of point sets (scene + pasted objects, train.py:29-42)
    pcd6 = merged_points_with_proj(pcd4, calib_np) if isinstance(calib_np, (list, tuple)) else points_with_proj(pcd4, calib_np)
    t['project'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    voxel9, uidx = group(pcd6, grid.velorange, grid.voxelsize, grid.T)
    t['voxelize'] = time.perf_counter() - t0
    voxels = torch.Tensor(voxel9).to(device)                              # fp64 -> fp32 (train.py:125)
    idx = torch.LongTensor(np.concatenate([np.zeros((uidx.shape[0], 1)), uidx], axis=1)).to(device)   # train.py:119,126
    imsize = torch.Tensor(list(imsize_hw)).to(device)
    feats = [torch.from_numpy(np.asarray(m)).to(device) for m in fpn_maps]
    t0 = time.perf_counter()
    im768 = feature_mapping(voxels, feats, imsize, eps)
    t['gather'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    im16 = fusion(im768[None].to(dtype), sd, eps)
    t['fusion'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    x23 = torch.concat([voxels[None][..., :7].to(dtype), im16], dim=-1)  # MVXNet.py:26
    vfeat = voxel_features(x23, sd, eps)
    t['vfe'] = time.perf_counter() - t0
    out = {'voxels9': voxels, 'idx': idx, 'im768': im768, 'im16': im16[0], 'vfeat': vfeat, 'times': t}
    if want_grid:
        t0 = time.perf_counter()
        out['grid'] = reindex(vfeat, idx, grid.voxelshape)
        t['scatter'] = time.perf_counter() - t0
    if stages is not None:
        stages.update(t)
    return out


# --------------------------------------------------------------------------- training mode: gradients of the hot-path layers
def backward_frame(pcd4: np.ndarray, calib_np, fpn_maps: List[np.ndarray], sd_np, grid, imsize_hw, d_vfeat: np.ndarray,
                   eps: float = 1e-6, dtype: torch.dtype = torch.float64):
    """What `loss.backward()` (train.py:140-166) leaves in `.grad` of the 8 hot-path layers for ONE frame, when the
    gradient arriving at the voxel features (the input of `reindex`/CML, VoxelNet.py:33-36) is `d_vfeat` (N,128):
    autograd through the dense reference chain featureMaping -> ImageFeatureFusion -> concat -> SVFE/FCN/max
    (Pipe.py:84-104; MVXNet.py:25-26; voxelnet/Pipe.py:5-29; VoxelNet.py:27-32). The FPN maps are inputs (no gradient).
    Returns (vfeat (N,128), {state-dict name: gradient})."""
    params = {k: torch.from_numpy(np.asarray(v)).to(dtype).requires_grad_(True) for k, v in sd_np.items()}
    pcd6 = points_with_proj(pcd4, calib_np)
    voxel9, _ = group(pcd6, grid.velorange, grid.voxelsize, grid.T)
    voxels = torch.Tensor(voxel9)
    feats = [torch.from_numpy(m) for m in fpn_maps]
    im768 = feature_mapping(voxels, feats, torch.Tensor(list(imsize_hw)), eps)
    im16 = fusion(im768[None].to(dtype), params, eps)
    x23 = torch.concat([voxels[None][..., :7].to(dtype), im16], dim=-1)
    vfeat = voxel_features(x23, params, eps)
    (vfeat * torch.from_numpy(np.asarray(d_vfeat)).to(dtype)).sum().backward()
    return vfeat.detach(), {k: p.grad for k, p in params.items()}
