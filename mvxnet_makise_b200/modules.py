"""Host-side mirror of the reference's nn.Module surface for the hot path (SURVEY.md §8b seam 3).

Same class names, constructor arguments, parameter names/shapes (checkpoint compatible) and forward
signatures as the reference; the arithmetic runs in the sm_100a kernels behind the C ABI.
Forward only in this round: the kernels do not record autograd graphs (module outputs are detached).

  lidar2Img            modules/utils/Calib.py:47-70
  featureMaping        modules/imhead/Pipe.py:23-82
  FCN, CRB2d           modules/layers/Blocks.py:5-18, 31-40
  ImageFeatureFusion   modules/imhead/Pipe.py:84-104
  VFE, SVFE            modules/voxelnet/Pipe.py:5-29
  VoxelNetHead         modules/voxelnet/VoxelNet.py:16-34 (svfe, fcn, max over T, reindex)
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

EPS = 1e-6          # cfg.eps with half: False (modules/config/Config.py:11-13)
SAMPLENUM = 35      # cfg.samplenum (config.yml:21)
VOXELSHAPE = (352, 400, 10)


def pack_calib(calib: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[R0_rect @ Tr_velo_to_cam | P2] as 32 fp32 values. The 4x4 product is formed on the host in fp32 exactly as
    the reference forms it (Calib.py:65: `calib['R0_rect'] @ calib['Tr_velo_to_cam'] @ points` associates left)."""
    r0 = torch.as_tensor(calib['R0_rect'], dtype=torch.float32).cpu()
    tr = torch.as_tensor(calib['Tr_velo_to_cam'], dtype=torch.float32).cpu()
    p2 = torch.as_tensor(calib['P2'], dtype=torch.float32).cpu()
    return torch.cat([(r0 @ tr).reshape(-1), p2.reshape(-1)]).contiguous()


def calib_is_f64(calib: Dict) -> bool:
    """True when the calibration holds float64 matrices - what `readCalib` returns (Load.py:24-41: `np.zeros((4, 4))` and
    `np.concatenate([float32 block, [[0, 0, 0, 1]]])` are float64). The reference's numpy branches then evaluate the whole
    projection in fp64 (cropToSight at Load.py:73; lidar2Img of the pasted ground-truth sets, train.py:36-39); its torch
    branches see `torch.Tensor(calib[k])`, i.e. fp32 (Load.py:75-76, train.py:113-115)."""
    return any((isinstance(v, np.ndarray) and v.dtype == np.float64) or (isinstance(v, torch.Tensor) and v.dtype == torch.float64)
               for v in (calib['R0_rect'], calib['Tr_velo_to_cam'], calib['P2']))


def pack_calib64(calib: Dict) -> torch.Tensor:
    """[R0_rect @ Tr_velo_to_cam | P2] as 32 fp64 values, the 4x4 product formed by numpy in fp64 like Calib.py:65 does."""
    r0, tr, p2 = (np.asarray(calib[k].cpu() if isinstance(calib[k], torch.Tensor) else calib[k], dtype=np.float64)
                  for k in ('R0_rect', 'Tr_velo_to_cam', 'P2'))
    return torch.from_numpy(np.concatenate([(r0 @ tr).reshape(-1), p2.reshape(-1)])).contiguous()


def lidar2Img(pcd, calib: dict, uncheck: bool = False):
    """Project points to the image; returns (N,2) in (width, height) order like the reference.
    Accepts numpy or torch input and returns the same kind (torch results stay on the GPU). numpy points with a float64
    calibration (readCalib's dicts) follow the reference's numpy branch: fp64 arithmetic, float64 result."""
    _lib.require_cuda()
    assert pcd.ndim == 2, 'Point cloud should be in (N, 3 + C)'
    as_numpy = isinstance(pcd, np.ndarray)
    pts = torch.from_numpy(np.ascontiguousarray(pcd, dtype=np.float32)).cuda() if as_numpy \
        else pcd.to(device='cuda', dtype=torch.float32).contiguous()
    if as_numpy and calib_is_f64(calib):
        c64 = pack_calib64(calib).cuda()
        out = torch.empty((pts.shape[0], 2), dtype=torch.float64, device=pts.device)
        check(lib.mvx_lidar2img_f64(ptr(pts), pts.shape[1], pts.shape[0], ptr(c64), ptr(out), stream_ptr()), 'lidar2img_f64')
        if not uncheck:
            m = c64[:16].reshape(4, 4)
            depth = pts[:, :3].double() @ m[2, :3] + m[2, 3]
            out = out[depth > 0]
        return out.cpu().numpy()
    c32 = pack_calib(calib).cuda()
    out = torch.empty((pts.shape[0], 2), dtype=torch.float32, device=pts.device)
    check(lib.mvx_lidar2img(ptr(pts), pts.shape[1], pts.shape[0], ptr(c32), ptr(out), stream_ptr()), 'lidar2img')
    if not uncheck:   # Calib.py:66-67: drop points behind the camera (depth = row 2 of (R0@Tr)@p)
        m = c32[:16].reshape(4, 4)
        depth = pts[:, :3] @ m[2, :3] + m[2, 3]
        out = out[depth > 0]
    return out.cpu().numpy() if as_numpy else out


def featureMaping(voxels, features: List[torch.Tensor], calibs, imsize: torch.Tensor, eps: float = EPS):
    """Drop-in for ``featureMaping`` (Pipe.py:23-82). voxels: batch * (N,T,C>=9 is NOT required: exactly the
    reference layout (N,T,9)); features: 3 maps (batch,256,Hf,Wf); returns batch * (N,T,768).
    Side effects kept: pad slots (x==y==z==0) of ``voxels`` are zeroed in place (Pipe.py:53-59).
    Unlike the reference the ``features`` list is not replaced by padded copies (the pad is a bounds predicate)."""
    _lib.require_cuda()
    res = []
    hw = [(int(f.shape[-2]), int(f.shape[-1])) for f in features]
    C = int(features[0].shape[1])
    mh = (ctypes.c_int32 * 3)(*[h for h, _ in hw])
    mw = (ctypes.c_int32 * 3)(*[w for _, w in hw])
    nbytes = ctypes.c_size_t()
    check(lib.mvx_maps_nhwc_bytes(mh, mw, C, ctypes.byref(nbytes)), 'maps_nhwc_bytes')
    imh, imw = float(imsize[0]), float(imsize[1])
    for i, v in enumerate(voxels):
        assert v.is_cuda and v.dtype == torch.float32 and v.shape[-1] == 9
        vc = v if v.is_contiguous() else v.contiguous()
        R = vc.numel() // 9
        maps = [f[i].contiguous() for f in features]
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=v.device)
        out = torch.empty(tuple(v.shape[:-1]) + (3 * C,), dtype=torch.float32, device=v.device)
        mp = (ctypes.c_void_p * 3)(*[m.data_ptr() for m in maps])
        check(lib.mvx_feature_mapping(ptr(vc), R, mp, mh, mw, C, imh, imw, float(eps), ptr(out), ptr(ws), nbytes.value,
                                      stream_ptr()), 'feature_mapping')
        if vc is not v:
            v.copy_(vc)
        res.append(out)
    return res


def _pad_cols(x2d: torch.Tensor, mult: int = 16) -> torch.Tensor:
    cin = x2d.shape[1]
    pad = (-cin) % mult
    if pad == 0:
        return x2d.contiguous()
    return torch.cat([x2d, x2d.new_zeros((x2d.shape[0], pad))], dim=1).contiguous()


def _wt(weight: torch.Tensor, mult: int = 16) -> torch.Tensor:
    """(Cout, Cin[,1,1]) -> W^T (Cin_pad, Cout) fp32 contiguous, zero rows for the padded inputs."""
    w = weight.detach().reshape(weight.shape[0], -1).to(torch.float32)
    wt = w.t().contiguous()
    pad = (-wt.shape[0]) % mult
    if pad:
        wt = torch.cat([wt, wt.new_zeros((pad, wt.shape[1]))], dim=0).contiguous()
    return wt


def _layer_forward(kind: str, x: torch.Tensor, weight, bias, eps: float, T: int = 1):
    """x (..., Cin) CUDA fp32 -> BN(relu(x W^T + b)) with batch statistics over all leading dims."""
    _lib.require_cuda()
    lead = x.shape[:-1]
    cout = weight.shape[0]
    x2 = _pad_cols(x.detach().reshape(-1, x.shape[-1]).to(torch.float32))
    wt = _wt(weight)
    b = bias.detach().to(torch.float32).contiguous()
    R, cin = x2.shape
    nbytes = ctypes.c_size_t()
    check(lib.mvx_layer_workspace_bytes(cin, cout, ctypes.byref(nbytes)), 'layer_workspace_bytes')
    stats = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
    if kind == 'fcn':
        y = torch.empty((R, cout), dtype=torch.float32, device=x.device)
        check(lib.mvx_fcn_forward(ptr(x2), R, cin, ptr(wt), ptr(b), cout, float(eps), ptr(y), ptr(stats), stream_ptr()),
              'fcn_forward')
        return y.reshape(lead + (cout,))
    vmax = torch.empty((R // T, cout), dtype=torch.int32, device=x.device)
    if kind == 'vfe':
        y = torch.empty((R, 2 * cout), dtype=torch.float32, device=x.device)
        check(lib.mvx_vfe_forward(ptr(x2), R, T, cin, ptr(wt), ptr(b), cout, float(eps), ptr(y), ptr(stats), ptr(vmax),
                                  stream_ptr()), 'vfe_forward')
        return y.reshape(lead + (2 * cout,))
    y = torch.empty((R // T, cout), dtype=torch.float32, device=x.device)
    check(lib.mvx_fcn_max_forward(ptr(x2), R, T, cin, ptr(wt), ptr(b), cout, float(eps), ptr(y), ptr(stats), ptr(vmax),
                                  stream_ptr()), 'fcn_max_forward')
    return y


class FCN(nn.Module):
    """Blocks.py:5-18: relu(Linear) -> BatchNorm2d(affine=False, track_running_stats=False). Input (batch,h,w,c)."""

    def __init__(self, cin, cout, eps: float = EPS):
        super().__init__()
        self.fc = nn.Linear(cin, cout)
        self.eps = eps

    def forward(self, x):
        return _layer_forward('fcn', x, self.fc.weight, self.fc.bias, self.eps)


class CRB2d(nn.Module):
    """Blocks.py:31-40 for the 1x1/stride-1/no-pad case the path uses (Pipe.py:89,91). Input (batch,c,h,w)."""

    def __init__(self, cin, cout, k, s, p, eps: float = EPS):
        super().__init__()
        if (k, s, p) != (1, 1, 0):
            raise NotImplementedError('only the 1x1 CRB2d of the fusion stack is on the hot path')
        self.conv = nn.Conv2d(cin, cout, k, s, p)
        self.eps = eps

    def forward(self, x):
        y = _layer_forward('fcn', x.permute(0, 2, 3, 1), self.conv.weight, self.conv.bias, self.eps)
        return y.permute(0, 3, 1, 2)


class ImageFeatureFusion(nn.Module):
    """Pipe.py:84-104: 768 -> 768 -> 128 -> 128 -> 16 -> 16."""

    def __init__(self):
        super().__init__()
        self.fcn1 = FCN(768, 768)
        self.conv1 = CRB2d(768, 128, 1, 1, 0)
        self.fcn2 = FCN(128, 128)
        self.conv2 = CRB2d(128, 16, 1, 1, 0)
        self.fcn3 = FCN(16, 16)

    def forward(self, x):
        x = self.fcn1(x)
        x = self.conv1(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        x = self.fcn2(x)
        x = self.conv2(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        return self.fcn3(x)


class VFE(nn.Module):
    """voxelnet/Pipe.py:5-18: FCN -> max over T -> tile -> concat. Input (batch,N,T,cin) -> (batch,N,T,2*cout)."""

    def __init__(self, cin, cout, sampleNum):
        super().__init__()
        self.fcn = FCN(cin, cout)
        self.sampleNum = sampleNum

    def forward(self, x):
        assert x.shape[2] == self.sampleNum
        return _layer_forward('vfe', x, self.fcn.fc.weight, self.fcn.fc.bias, self.fcn.eps, T=self.sampleNum)


class SVFE(nn.Module):
    """voxelnet/Pipe.py:20-29."""

    def __init__(self, sampleNum=SAMPLENUM):
        super().__init__()
        self.vfe1 = VFE(7 + 16, 16, sampleNum)
        self.vfe2 = VFE(32, 64, sampleNum)

    def forward(self, x):
        return self.vfe2(self.vfe1(x))


def reindex(x: torch.Tensor, idx: torch.Tensor, voxelshape: Sequence[int] = VOXELSHAPE) -> torch.Tensor:
    """Drop-in for ``VoxelNet.reindex`` (VoxelNet.py:16-22): (N,C) + idx (N,4) int64 -> (1,C,nz,nx,ny)."""
    _lib.require_cuda()
    nx, ny, nz = (int(s) for s in voxelshape)
    x = x.detach().to(torch.float32).contiguous()
    idx = idx.to(torch.int64).contiguous()
    N, C = x.shape
    out = torch.empty((1, C, nz, nx, ny), dtype=torch.float32, device=x.device)
    G = nx * ny * nz
    mapws = torch.empty(G + G // 32 + 1, dtype=torch.int32, device=x.device)
    check(lib.mvx_scatter_dense(ptr(x), ptr(idx), N, C, nx, ny, nz, ptr(out), ptr(mapws), stream_ptr()), 'scatter_dense')
    return out


class VoxelNetHead(nn.Module):
    """The hot-path part of ``VoxelNet`` (VoxelNet.py:9-34): svfe, fcn, max over T, reindex. ``cml``/``rpn`` stay the
    reference's stock modules and consume the returned grid. Parameter names match ``backbone.svfe.*``/``backbone.fcn.*``."""

    def __init__(self, sampleNum=SAMPLENUM, voxelshape=VOXELSHAPE):
        super().__init__()
        self.svfe = SVFE(sampleNum)
        self.fcn = FCN(128, 128)
        self.sampleNum = sampleNum
        self.voxelshape = tuple(voxelshape)

    reindex = staticmethod(reindex)

    def voxel_features(self, x):
        x = self.svfe(x)
        return _layer_forward('fcn_max', x, self.fcn.fc.weight, self.fcn.fc.bias, self.fcn.eps, T=self.sampleNum)

    def forward(self, x, idx):
        return reindex(self.voxel_features(x), idx, self.voxelshape)
