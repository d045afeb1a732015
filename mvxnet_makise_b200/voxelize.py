"""Host side of stage 1: the reference's voxelization entry points on the B200 kernels.

Mirrors (same names, argument meaning and return types):
  * ``cpp._group``                     — cpp/voxelutil.cpp:325-360 via modules/Extension.py:1-3
  * ``group`` (numba) / ``group_``     — modules/data/Preprocessing.py:75-116 / :57-73
  * ``crop`` / ``cropTensor`` / ``cropToSight`` — modules/data/Preprocessing.py:12-55 (the step before the path)
The in-function shuffle of ``group``/``group_`` is kept (``shuffle=True`` default); parity is defined on the
post-shuffle order, so tests call with ``shuffle=False`` (SURVEY.md trap 3).
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib
from . import anchors as _anchors
from ._lib import lib, check, ptr, stream_ptr


def _cap_for(max_points: int) -> int:
    return max(128, (int(max_points) + 127) // 128 * 128)


class VoxelBatch:
    """Device-side result of one batched voxelization (compact form)."""

    def __init__(self, B, cap, device):
        i32 = dict(dtype=torch.int32, device=device)
        self.B, self.cap = B, cap
        self.counts = torch.empty((B, 4), **i32)
        self.vox_coord = torch.empty((B, cap, 4), **i32)
        self.vox_cnt = torch.empty((B, cap), **i32)
        self.vox_row0 = torch.empty((B, cap + 1), **i32)
        self.row_point = torch.empty((B, cap), **i32)
        self.row_vox = torch.empty((B, cap), **i32)
        self.cell2vid = None

    def as_struct(self) -> _lib.VoxelOut:
        return _lib.VoxelOut(ptr(self.counts), ptr(self.vox_coord), ptr(self.vox_cnt), ptr(self.vox_row0),
                             ptr(self.row_point), ptr(self.row_vox), ptr(self.cell2vid))


def voxelize(points: torch.Tensor, offsets: Sequence[int], T: int, grid: _lib.Grid | None = None,
             cell_idx: torch.Tensor | None = None, want_cell2vid: bool = False) -> VoxelBatch:
    """Batched deterministic voxelization of concatenated CUDA points (sum P, stride) fp32.

    offsets: host list [B+1] of point offsets. Either ``grid`` (cell index computed in fp64 on the GPU) or
    ``cell_idx`` ((sum P, 3) int32 CUDA, the caller-computed idx of ``_group``) must be given."""
    _lib.require_cuda()
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2
    B = len(offsets) - 1
    maxp = max([offsets[i + 1] - offsets[i] for i in range(B)] + [0])
    cap = _cap_for(maxp)
    out = VoxelBatch(B, cap, points.device)
    if want_cell2vid:
        G = grid.shape[0] * grid.shape[1] * grid.shape[2]
        out.cell2vid = torch.empty((B, G), dtype=torch.int32, device=points.device)
    nbytes = ctypes.c_size_t()
    check(lib.mvx_voxelize_workspace_bytes(B, cap, ctypes.byref(nbytes)), 'voxelize_workspace_bytes')
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=points.device)
    off = (ctypes.c_int32 * (B + 1))(*[int(o) for o in offsets])
    st = out.as_struct()
    check(lib.mvx_voxelize(ctypes.byref(grid) if grid is not None else None, B, cap, ptr(points), points.shape[1], off,
                           ptr(cell_idx), int(T), ctypes.byref(st), ptr(ws), nbytes.value, stream_ptr()), 'voxelize')
    out._ws = ws   # keep alive until the stream has consumed it
    return out


def _to_cuda_f32(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device='cuda', dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def cpp_group(pcd, idx, samplesPerVoxel: int):
    """Drop-in for ``cpp._group(pcd, idx, samplesPerVoxel)`` (voxelutil.cpp:325-360).

    pcd (P, C>=4) float (force-cast to fp32 like pybind's array_t<float>), idx (P,3) int (force-cast to int32).
    Returns (voxel (V,T,7) float32, (x,y,z) int64[V] each, cnt int64[V]) as numpy arrays."""
    T = int(samplesPerVoxel)
    if np.ndim(pcd) != 2 or np.ndim(idx) != 2:
        raise ValueError('array has incorrect number of dimensions: expected 2')   # pybind unchecked<2>()
    pts = _to_cuda_f32(pcd)
    if isinstance(idx, torch.Tensor):
        ci = idx.to(device='cuda', dtype=torch.int32).contiguous()
    else:
        ci = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).cuda()
    P = ci.shape[0]
    if P == 0:
        z = np.zeros(0, dtype=np.int64)
        return np.zeros((0, T, 7), dtype=np.float32), (z, z.copy(), z.copy()), z.copy()
    vb = voxelize(pts, [0, P], T, grid=None, cell_idx=ci)
    V = int(vb.counts[0, 0].item())
    dev = pts.device
    voxel = torch.empty((V, T, 7), dtype=torch.float32, device=dev)
    xyzc = torch.empty((4, V), dtype=torch.int64, device=dev)
    check(lib.mvx_group_emit7(ptr(pts), pts.shape[1], V, T, ptr(vb.vox_coord), ptr(vb.vox_cnt), ptr(vb.vox_row0),
                              ptr(vb.row_point), ptr(voxel), ptr(xyzc[0]), ptr(xyzc[1]), ptr(xyzc[2]), ptr(xyzc[3]),
                              stream_ptr()), 'group_emit7')
    h = xyzc.cpu().numpy()
    return voxel.cpu().numpy(), (h[0].copy(), h[1].copy(), h[2].copy()), h[3].copy()


def group_(pcd: np.ndarray, range: Sequence[float], size: Sequence[float], samplesPerVoxel: int, shuffle: bool = True):
    """Drop-in for ``group_`` (Preprocessing.py:57-73): the reference's own numpy glue around ``cpp._group``."""
    if shuffle:
        np.random.shuffle(pcd)
    pts = pcd[:, :3]
    low = np.array(range[0:3])
    idx = ((pts - low) / size).astype('int32')
    voxel, uidx, vcnt = cpp_group(pcd, idx, samplesPerVoxel)
    center = voxel[..., :3].sum(axis=1) / vcnt[:, None]
    voxel[..., 3:6] = voxel[..., :3] - center[:, None, :]
    return voxel, np.array(uidx).T


def group(pcd: np.ndarray, range: Sequence[float], size: Sequence[float], samplesPerVoxel: int, shuffle: bool = True,
          voxelshape: Sequence[int] | None = None, device_out: bool = False):
    """Drop-in for the numba ``group`` (Preprocessing.py:75-116), the variant train.py:44 calls.

    pcd (P,6) fp32 [x,y,z,r,row,col]; returns (voxel (V,T,9) float64, uidx (V,3) float64) numpy arrays.
    The cell index is computed on the GPU in fp64 exactly like ``((pts - low) / size).astype(int32)``.
    ``voxelshape`` defaults to round((hi-lo)/size). With ``device_out`` the fp32 voxel tensor (what
    train.py:125 would upload) and an int64 uidx stay on the GPU."""
    T = int(samplesPerVoxel)
    if shuffle:
        np.random.shuffle(pcd)
    if voxelshape is None:
        voxelshape = [int(round((range[i + 3] - range[i]) / size[i])) for i in (0, 1, 2)]
    g = _lib.make_grid(range, size, voxelshape, T)
    pts = _to_cuda_f32(pcd)
    P = pts.shape[0]
    if P == 0:
        return np.empty((0, T, 9)), np.empty((0, 3))
    vb = voxelize(pts, [0, P], T, grid=g)
    c = vb.counts[0].cpu()
    if int(c[2]) != 0:
        raise IndexError(f'{int(c[2])} points fall outside the voxel grid (the reference expects cropped input)')
    V = int(c[0])
    dev = pts.device
    if device_out:
        v32 = torch.empty((V, T, 9), dtype=torch.float32, device=dev)
        check(lib.mvx_group_emit9(ptr(pts), pts.shape[1], V, T, ptr(vb.vox_coord), ptr(vb.vox_cnt), ptr(vb.vox_row0),
                                  ptr(vb.row_point), None, ptr(v32), None, stream_ptr()), 'group_emit9')
        return v32, vb.vox_coord[0, :V, :3].to(torch.int64)
    v64 = torch.empty((V, T, 9), dtype=torch.float64, device=dev)
    u64 = torch.empty((V, 3), dtype=torch.float64, device=dev)
    check(lib.mvx_group_emit9(ptr(pts), pts.shape[1], V, T, ptr(vb.vox_coord), ptr(vb.vox_cnt), ptr(vb.vox_row0),
                              ptr(vb.row_point), ptr(v64), None, ptr(u64), stream_ptr()), 'group_emit9')
    return v64.cpu().numpy(), u64.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------------
# Before the path: range crop and field-of-view crop with order-preserving compaction (Preprocessing.py:12-55)
def crop_frames(points: torch.Tensor, offsets: Sequence[int], range6: Sequence[float] | None = None,
                calib32: torch.Tensor | None = None, imsize_wh: Sequence[float] | None = None,
                calib64: torch.Tensor | None = None):
    """Batched device entry. points (sum P, C>=3) fp32 CUDA, offsets host [B+1]; range6 = velorange or None;
    calib32 (B,32) CUDA ([R0@Tr | P2], `modules.pack_calib`) + imsize (w, h) or None. Returns (out, counts): `out` has
    the layout of `points` with the kept points of frame f compacted, in input order, at rows offsets[f] ..
    offsets[f] + counts[f]; counts (B,) int32 CUDA. calib64 (B,32) float64 CUDA (`modules.pack_calib64`) instead of calib32
    runs the sight test in fp64 - the reference's numpy branch with readCalib's float64 matrices (Load.py:73)."""
    _lib.require_cuda()
    assert points.is_cuda and points.dtype == torch.float32 and points.is_contiguous() and points.dim() == 2
    B = len(offsets) - 1
    off = (ctypes.c_int32 * (B + 1))(*[int(o) for o in offsets])
    maxp = max([offsets[i + 1] - offsets[i] for i in range(B)] + [0])
    nbytes = ctypes.c_size_t()
    check(lib.mvx_crop_workspace_bytes(B, int(offsets[-1]), int(maxp), ctypes.byref(nbytes)), 'crop_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=points.device)
    out = torch.empty_like(points)
    counts = torch.empty(B, dtype=torch.int32, device=points.device)
    rng = (ctypes.c_double * 6)(*[float(v) for v in range6]) if range6 is not None else None
    w, h = (float(imsize_wh[0]), float(imsize_wh[1])) if imsize_wh is not None else (0.0, 0.0)
    if calib32 is not None:
        assert calib32.is_cuda and calib32.dtype == torch.float32 and calib32.is_contiguous() and tuple(calib32.shape) == (B, 32)
        assert imsize_wh is not None
    if calib64 is not None:
        assert calib32 is None and imsize_wh is not None
        assert calib64.is_cuda and calib64.dtype == torch.float64 and calib64.is_contiguous() and tuple(calib64.shape) == (B, 32)
        check(lib.mvx_crop_points_f64(ptr(points), points.shape[1], B, off, rng, ptr(calib64), w, h, ptr(out), ptr(counts), ptr(ws),
                                      ws.numel(), stream_ptr()), 'crop_points_f64')
        return out, counts
    check(lib.mvx_crop_points(ptr(points), points.shape[1], B, off, rng, ptr(calib32), w, h, ptr(out), ptr(counts), ptr(ws),
                              ws.numel(), stream_ptr()), 'crop_points')
    return out, counts


def _crop_one(pcd, range6, calib, imsize):
    from .modules import pack_calib, pack_calib64, calib_is_f64
    as_numpy = isinstance(pcd, np.ndarray)
    assert pcd.ndim == 2
    x = _to_cuda_f32(np.ascontiguousarray(pcd, dtype=np.float32) if as_numpy else pcd)
    # numpy points + float64 calibration (readCalib's dict, Load.py:73): the reference's numpy branch runs in fp64
    f64 = calib is not None and as_numpy and calib_is_f64(calib)
    c32 = pack_calib(calib)[None].to(x.device) if calib is not None and not f64 else None
    c64 = pack_calib64(calib)[None].to(x.device) if f64 else None
    out, counts = crop_frames(x, [0, x.shape[0]], range6, c32, imsize, calib64=c64)
    kept = out[:int(counts[0].item())]
    if as_numpy:
        return kept.cpu().numpy().astype(pcd.dtype, copy=False)
    return kept.to(pcd.dtype)


def crop(pcd, range: Sequence[float]):
    """`crop(pcd, range)` — Preprocessing.py:12-17: points with low <= xyz < high, input order kept. numpy in -> numpy out,
    torch in -> torch (CUDA) out (`cropTensor`, Preprocessing.py:19-24, is the same filter)."""
    return _crop_one(pcd, list(range), None, None)


cropTensor = crop


def cropToSight(pcd, calib: dict, imsize: Sequence[int]):
    """`cropToSight(pcd, calib, imsize)` — Preprocessing.py:26-55: points in front of the camera whose projection falls
    inside the image (imsize is (w, h), minus the reference's 1e-3 fudge), input order kept."""
    return _crop_one(pcd, None, calib, imsize)


def cropFrame(pcd, range: Sequence[float], calib: dict, imsize: Sequence[int]):
    """crop + cropToSight in one pass (what Load.py:59,73 / cropdata.py:32-65 do back to back)."""
    return _crop_one(pcd, list(range), calib, imsize)


class VoxelUtil:
    """The object ``modules/Extension.py`` exposes as ``cpp`` (voxelutil.cpp:362-368): all four bound functions, each
    on the CUDA kernels (``_group``: csrc/voxelize.cu; the anchor / rotated-IoU functions: csrc/anchors.cu)."""

    _group = staticmethod(cpp_group)
    _classifyAnchors = staticmethod(_anchors._classifyAnchors)
    bboxOverlap = staticmethod(_anchors.bboxOverlap)
    bboxIntersection = staticmethod(_anchors.bboxIntersection)


cpp = VoxelUtil()
