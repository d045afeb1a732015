"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by mvxnet_makise_b200/).

ctypes front of oracle/iou_oracle.c (rotated IoU + anchor classification, cpp/voxelutil.cpp:15-316) and torch-CPU
restatements of the python that surrounds it in the reference:
  bbox3d2bev       modules/Calc.py:9-36   (+ getRotationMatrices :9-13)
  createAnchors    modules/data/Preprocessing.py:118-142
  start cells      modules/Calc.py:91-94  (nls / nws of classifyAnchors)
Pinned against the live reference (oracle/_ref voxelutil, modules.Calc is not importable: shapely) by
tests/test_oracle_iou.py and against tests/golden/anchors_a.npz.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libiou_oracle.so')
        src = os.path.join(_HERE, 'iou_oracle.c')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(['make', '-s', '-C', _HERE, '_build/libiou_oracle.so'], check=True)
        lib = ctypes.CDLL(path)
        vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
        lib.iou_oracle_pairwise.argtypes = [vp, i64, vp, i64, ctypes.c_int, vp]
        lib.iou_oracle_pairwise.restype = None
        lib.iou_oracle_classify.argtypes = [vp, i64, vp, i64, i64, i64, vp, vp, f32, f32, vp, vp, vp, i64, vp]
        lib.iou_oracle_classify.restype = None
        _LIB = lib
    return _LIB


def _f32(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int64)


def pairwise(b1, b2, mode: str = 'iou') -> np.ndarray:
    """bboxOverlap ('iou') / bboxIntersection ('inter') of (N,4,2) against (M,4,2) corner quads -> (N,M) fp32."""
    b1, b2 = _f32(b1), _f32(b2)
    out = np.empty((b1.shape[0], b2.shape[0]), np.float32)
    _lib().iou_oracle_pairwise(b1.ctypes.data, b1.shape[0], b2.ctypes.data, b2.shape[0], 0 if mode == 'iou' else 1, out.ctypes.data)
    return out


def classify(gts, anchors, nls, nws, neg_thr: float, pos_thr: float):
    """`_classifyAnchors` (voxelutil.cpp:141-316): ((px,py,pz), (nx,ny,nz), gi) int64 arrays + number of ground truths whose
    start cell lies outside the anchor grid."""
    gts, anchors, nls, nws = _f32(gts), _f32(anchors), _i64(nls), _i64(nws)
    L, W, A = anchors.shape[:3]
    cap = 4096
    while True:
        pos, neg, gi = np.empty((cap, 3), np.int64), np.empty((cap, 3), np.int64), np.empty(cap, np.int64)
        counts = np.zeros(3, np.int64)
        _lib().iou_oracle_classify(gts.ctypes.data, gts.shape[0], anchors.ctypes.data, L, W, A, nls.ctypes.data, nws.ctypes.data,
                                   neg_thr, pos_thr, pos.ctypes.data, neg.ctypes.data, gi.ctypes.data, cap, counts.ctypes.data)
        if max(counts[0], counts[1]) <= cap:
            break
        cap = int(max(counts[0], counts[1]))
    npos, nneg = int(counts[0]), int(counts[1])
    pos, neg = pos[:npos], neg[:nneg]
    return ((pos[:, 0].copy(), pos[:, 1].copy(), pos[:, 2].copy()), (neg[:, 0].copy(), neg[:, 1].copy(), neg[:, 2].copy()),
            gi[:npos].copy(), int(counts[2]))


# ---- the python around the native calls (torch CPU, the reference's own arithmetic library) ---------------------------------
def rotation_matrices(r: torch.Tensor) -> torch.Tensor:   # Calc.py:9-13
    rcos = torch.cos(r).reshape((-1, 1))
    rsin = torch.sin(r).reshape((-1, 1))
    return torch.concat([rcos, -rsin, rsin, rcos], dim=1).reshape((-1, 2, 2))


def bbox3d2bev(bbox3ds: torch.Tensor) -> torch.Tensor:   # Calc.py:15-36: (..., 7) xyzlwhr -> (..., 4, 2) corners
    origshape = bbox3ds.shape[:-1]
    b = bbox3ds.reshape((-1, bbox3ds.shape[-1]))
    res = torch.tensor([[0.5, 0.5], [-0.5, 0.5], [-0.5, -0.5], [0.5, -0.5]], dtype=torch.float32)
    res = torch.tile(res, (b.shape[0], 1, 1)) * b[:, None, [3, 4]]
    res = res @ rotation_matrices(b[:, 6])
    res = res + b[:, None, [0, 1]]
    return res.reshape(origshape + (4, 2)) if len(origshape) > 0 else res[0]


def create_anchors(l: int, w: int, rng, size) -> torch.Tensor:   # Preprocessing.py:118-142 -> (l, w, 14)
    ls = (rng[3] - rng[0]) / l
    ws = (rng[4] - rng[1]) / w
    x = torch.linspace(rng[0] + ls / 2, rng[3] - ls / 2, l)
    y = torch.linspace(rng[1] + ws / 2, rng[4] - ws / 2, w)
    x, y = torch.meshgrid(x, y, indexing='ij')
    g = torch.concat([x[..., None], y[..., None]], dim=2)
    size = torch.tile(torch.tensor(size, dtype=torch.float32), (l, w, 1))
    z = torch.full((l, w, 1), -1.0)
    t = torch.zeros((l, w, 1))
    t2 = torch.full((l, w, 1), torch.pi / 2)
    return torch.concat([torch.concat([g, z, size, t], dim=2), torch.concat([g, z, size, t2], dim=2)], dim=2)


def anchor_bevs(anchors14: torch.Tensor) -> torch.Tensor:   # train.py: anchors.reshape(l, w, 2, 7) -> bev corners (l, w, 2, 4, 2)
    l, w = anchors14.shape[:2]
    return bbox3d2bev(anchors14.reshape(l, w, 2, 7))


def start_cells(gt_centers: torch.Tensor, n_l: int, n_w: int, velorange):   # Calc.py:91-94
    l = (velorange[3] - velorange[0]) / n_l
    w = (velorange[4] - velorange[1]) / n_w
    nls = ((gt_centers[:, 0] - velorange[0] - l / 2) / l + 0.5).long()
    nws = ((gt_centers[:, 1] - velorange[1] - w / 2) / w + 0.5).long()
    return nls, nws
