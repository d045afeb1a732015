"""Host side of the label-side native functions on the B200 kernels of ``csrc/anchors.cu``.

Mirrors (same names, argument meaning, return types and append order):
  * ``cpp.bboxOverlap`` / ``cpp.bboxIntersection``  — cpp/voxelutil.cpp:96-139 (caller modules/augment/Augment.py:54)
  * ``cpp._classifyAnchors``                         — cpp/voxelutil.cpp:141-316
  * ``classifyAnchors``                              — modules/Calc.py:88-96 (caller train.py:46)
numpy / CPU-tensor arguments give numpy results like the pybind module; CUDA tensors stay on the device.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr


def _dev_f32(a, ndim: int, what: str) -> torch.Tensor:
    if np.ndim(a) != ndim:
        raise ValueError(f'array has incorrect number of dimensions: {np.ndim(a)}; expected {ndim}')   # pybind unchecked<n>()
    if isinstance(a, torch.Tensor):
        return a.detach().to(device='cuda', dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _dev_i64(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.detach().to(device='cuda', dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()


def _on_device(*args) -> bool:
    return any(isinstance(a, torch.Tensor) and a.is_cuda for a in args)


def _pairwise(bboxes1, bboxes2, mode: int):
    _lib.require_cuda()
    b1, b2 = _dev_f32(bboxes1, 3, 'bboxes1'), _dev_f32(bboxes2, 3, 'bboxes2')
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=b1.device)
    check(lib.mvx_bbox_pairwise(ptr(b1), n, ptr(b2), m, mode, ptr(out), stream_ptr()), 'bbox_pairwise')
    return out if _on_device(bboxes1, bboxes2) else out.cpu().numpy()


def bboxOverlap(bboxes1, bboxes2):
    """``cpp.bboxOverlap(bboxes1 (N,4,2), bboxes2 (M,4,2)) -> (N,M)`` rotated IoU of BEV corner quads (voxelutil.cpp:96-116)."""
    return _pairwise(bboxes1, bboxes2, 0)


def bboxIntersection(bboxes1, bboxes2):
    """``cpp.bboxIntersection`` — intersection areas (voxelutil.cpp:118-139)."""
    return _pairwise(bboxes1, bboxes2, 1)


class AnchorClassifier:
    """Keeps the anchor BEV quads (L,W,A,4,2) on the device (2.25 MB for the 176x200x2 KITTI grid; train.py:59-61 builds them
    once) and classifies the ground truths of one or several frames per call."""

    def __init__(self, anchor_bevs):
        _lib.require_cuda()
        self.anchors = _dev_f32(anchor_bevs, 5, 'anchors')
        if self.anchors.shape[3:] != (4, 2):
            raise ValueError('anchors must be (L, W, A, 4, 2) corner quads')
        self.L, self.W, self.A = (int(s) for s in self.anchors.shape[:3])

    def classify_device(self, gts, nls, nws, negThr: float, posThr: float):
        """Returns device tensors (pos (npos,3), neg (nneg,3), gi (npos), n_outside). One host read of the counts."""
        g = _dev_f32(gts, 3, 'gts')
        G = g.shape[0]
        nl, nw = _dev_i64(nls), _dev_i64(nws)
        dev = self.anchors.device
        nbytes = ctypes.c_size_t()
        check(lib.mvx_classify_anchors_workspace_bytes(G, self.L, self.W, self.A, ctypes.byref(nbytes)), 'classify_anchors_workspace_bytes')
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        counts = torch.empty(4, dtype=torch.int64, device=dev)
        cap = max(1024, 128 * G * self.A)
        while True:
            pos = torch.empty((cap, 3), dtype=torch.int64, device=dev)
            neg = torch.empty((cap, 3), dtype=torch.int64, device=dev)
            gi = torch.empty(cap, dtype=torch.int64, device=dev)
            check(lib.mvx_classify_anchors(ptr(g), G, ptr(self.anchors), self.L, self.W, self.A, ptr(nl), ptr(nw), float(negThr),
                                           float(posThr), ptr(pos), ptr(neg), ptr(gi), cap, ptr(counts), ptr(ws), nbytes.value,
                                           stream_ptr()), 'classify_anchors')
            c = counts.cpu()
            npos, nneg = int(c[0]), int(c[1])
            if max(npos, nneg) <= cap:
                return pos[:npos], neg[:nneg], gi[:npos], int(c[2])
            cap = max(npos, nneg)

    def __call__(self, gts, nls, nws, negThr: float, posThr: float):
        pos, neg, gi, outside = self.classify_device(gts, nls, nws, negThr, posThr)
        if outside:
            raise IndexError(f'{outside} ground truths start outside the anchor grid (an unchecked read in the reference)')
        p, n = pos.t().contiguous(), neg.t().contiguous()
        if _on_device(gts):
            return (p[0], p[1], p[2]), (n[0], n[1], n[2]), gi
        p, n = p.cpu().numpy(), n.cpu().numpy()
        return (p[0].copy(), p[1].copy(), p[2].copy()), (n[0].copy(), n[1].copy(), n[2].copy()), gi.cpu().numpy()


def _classifyAnchors(gts, anchors, nls, nws, negThr: float, posThr: float):
    """``cpp._classifyAnchors(gts (G,4,2), anchors (L,W,A,4,2), nls int64[G], nws int64[G], negThr, posThr)
    -> ((px,py,pz), (nx,ny,nz), gi)`` int64 arrays in the reference's append order (voxelutil.cpp:141-316)."""
    a = anchors if isinstance(anchors, AnchorClassifier) else AnchorClassifier(anchors)
    return a(gts, nls, nws, negThr, posThr)


def classifyAnchors(gts, gtCenters: torch.Tensor, anchors, velorange: Sequence[float], negThr: float, posThr: float):
    """``classifyAnchors`` of modules/Calc.py:88-96: start cell of every ground truth from its centre (torch fp32 on the host,
    the reference's own arithmetic; a handful of values), then the native classification. ``anchors`` may be the (L,W,A,4,2)
    tensor or an ``AnchorClassifier`` holding it on the device."""
    n_l, n_w = (anchors.L, anchors.W) if isinstance(anchors, AnchorClassifier) else anchors.shape[:2]
    cell_l = (velorange[3] - velorange[0]) / n_l
    cell_w = (velorange[4] - velorange[1]) / n_w
    c = gtCenters.detach().cpu() if isinstance(gtCenters, torch.Tensor) else torch.as_tensor(gtCenters)
    nls = ((c[:, 0] - velorange[0] - cell_l / 2) / cell_l + 0.5).long()
    nws = ((c[:, 1] - velorange[1] - cell_w / 2) / cell_w + 0.5).long()
    return _classifyAnchors(gts, anchors, nls, nws, negThr, posThr)
