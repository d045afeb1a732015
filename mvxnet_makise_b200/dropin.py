"""Drop-in for the reference's own objects: the unmodified `MVXNet` model and `train.py` loop keep working, the hot path
underneath them runs on the sm_100a kernels.

    from MVXNet import MVXNet                       # the reference's class, untouched
    from mvxnet_makise_b200.dropin import accelerate, install_extension
    install_extension()                             # modules/Extension.py seam: cpp._group / _classifyAnchors / bboxOverlap / bboxIntersection
    model = accelerate(MVXNet().to(device))         # same parameters, same state_dict, same forward signature
    score, reg = model(voxel, img, idx, [calib], imsize)     # train.py:131, unchanged
    loss.backward(); opt.step()                     # train.py:161-162, unchanged: gradients reach head.fusion.* / backbone.svfe.* / backbone.fcn.*

What `accelerate` replaces is the body of `MVXNet.forward` (MVXNet.py:21-27) between the image backbone and the CML:
`featureMaping` -> `ImageFeatureFusion` -> concat -> `SVFE` -> `FCN` -> max over T -> `reindex` become one call of the fused
path on the dense-voxel entry (`PointPath.forward_voxels`), wrapped in a `torch.autograd.Function` whose backward is the CUDA
backward of the eight layers (`mvx_pointpath_backward`). `head.extractor` (frozen torchvision backbone), `backbone.cml` and
`backbone.rpn` stay the reference's own modules. The model's `nn.Parameter`s remain the single source of truth: they are read
on every forward (an optimiser step is picked up through the tensors' version counters) and receive `.grad` like any other
parameter, so `model.parameters()`, `AdamW`, `state_dict()` / `load_state_dict()` and checkpoints are unaffected.
"""
from __future__ import annotations

import sys
import types
from typing import List, Sequence

import numpy as np
import torch

from . import synth
from .pipeline import PointPath

_HOT = [name for name, *_ in synth.HOT_LAYERS]


def _get(module, dotted: str):
    for part in dotted.split('.'):
        module = getattr(module, part)
    return module


def hot_parameters(model) -> List[torch.nn.Parameter]:
    """The 16 tensors of the eight hot-path layers, checkpoint order: weight, bias per layer."""
    out = []
    for name in _HOT:
        out += [_get(model, name + '.weight'), _get(model, name + '.bias')]
    return out


def _push_parameters(path: PointPath, params: Sequence[torch.Tensor]):
    """parameters -> the W^T (Cin_pad, Cout) / bias tensors the kernels read; skipped while the version counters stand still"""
    versions = tuple((p.data_ptr(), p._version) for p in params)
    if getattr(path, '_param_versions', None) == versions:
        return
    with torch.no_grad():
        for l, (_, cin, cout, _) in enumerate(synth.HOT_LAYERS):
            w, b = params[2 * l], params[2 * l + 1]
            path.wt[l][:cin].copy_(w.reshape(cout, cin).t())
            path.bias[l].copy_(b)
    path._param_versions = versions


class HotPathFunction(torch.autograd.Function):
    """grid = hot_path(voxels, idx, fpn maps; 16 parameters). Differentiable with respect to the parameters only: the image
    backbone is frozen (Head.py:9-11) and the voxel tensor is data."""

    @staticmethod
    def forward(ctx, path: PointPath, voxels, idx, maps, *params):
        _push_parameters(path, params)
        nz, nx, ny = path.grid.shape[2], path.grid.shape[0], path.grid.shape[1]
        B = len(voxels) if isinstance(voxels, (list, tuple)) else 1
        grid = torch.empty((B, 128, nz, nx, ny), dtype=torch.float32, device=path.device)   # a fresh tensor: the CML saves it for ITS backward
        path.forward_voxels(voxels, idx, maps, want_grid=True, train=True, grid_out=grid)
        ctx.path = path
        ctx.call_id = path._train_call = getattr(path, '_train_call', 0) + 1
        return grid

    @staticmethod
    def backward(ctx, d_grid):
        path = ctx.path
        if path._train_call != ctx.call_id:
            raise RuntimeError('the hot path ran another training forward before this backward: its saved activations are gone '
                               '(one forward/backward at a time per accelerated model)')
        flat = torch.empty(sum(int(np.prod(s)) for _, _, s in PointPath.grad_layout()), dtype=torch.float32, device=path.device)
        path.backward(d_grid=d_grid.contiguous(), grad_flat=flat, accumulate=False)
        grads = [flat[o:o + int(np.prod(shape))].view(*shape) for _, o, shape in PointPath.grad_layout()]
        return (None, None, None, None, *grads)


def hot_path(path: PointPath, voxels, idx, maps, params: Sequence[torch.Tensor]) -> torch.Tensor:
    """The dense grid (B,128,nz,nx,ny) for the reference's (voxels, idx) arguments. Records an autograd node when gradients are
    enabled and a parameter wants one; otherwise takes the inference route (pixel-first fcn1, no saved activations)."""
    maps = [m.detach().to(torch.float32).contiguous() for m in maps]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return HotPathFunction.apply(path, voxels, idx, maps, *params)
    _push_parameters(path, params)
    grid, _ = path.forward_voxels(voxels, idx, maps, want_grid=True, train=False)
    return grid


def _forward(self, voxels, imgs, idx, calibs, imsize):
    """`MVXNet.forward(voxels, imgs, idx, calibs, imsize)` (MVXNet.py:21-27): same arguments, same (score, reg) result.
    `calibs` is accepted and unused, as in the reference's featureMaping (the projection already sits in voxels[..., 7:9])."""
    path = self._mvx_path
    hw = self._mvx_imsize.get(id(imsize))
    if hw is None:                                   # cfg.imsize travels as a tensor (train.py:69): read it once, not per step
        hw = (float(imsize[0]), float(imsize[1]))
        self._mvx_imsize[id(imsize)] = hw
    path.imsize_hw = hw
    feats = self.head.extractor(imgs)                # FPN levels '0','1','2' (Pipe.py:17-21), frozen
    grid = hot_path(path, voxels, idx.to(torch.int64), feats, hot_parameters(self))
    nx, ny = path.grid.shape[0], path.grid.shape[1]
    outs = []
    for b in range(grid.shape[0]):                   # CML / RPN are batch-1 in the reference (VoxelNet.py:35-37)
        x = self.backbone.cml(grid[b:b + 1])
        outs.append(self.backbone.rpn(x.reshape((1, -1, nx, ny))))
    return outs[0] if len(outs) == 1 else outs


def accelerate(model, grid: synth.GridSpec = synth.KITTI_GRID, imsize_hw: Sequence[int] = synth.KITTI_IMSIZE_HW, eps: float = 1e-6):
    """Swap the hot path of a reference `MVXNet` instance (already on its CUDA device) for the fused sm_100a path, in place.
    `grid` / `eps` mirror config.yml (velorange, voxelshape, samplenum; eps 1e-6 with half: False). Returns the model."""
    params = hot_parameters(model)                   # raises AttributeError if this is not the reference's module tree
    dev = params[0].device
    if dev.type != 'cuda':
        raise RuntimeError('accelerate() needs the model on a CUDA device; mvxnet_makise_b200 has no CPU fallback')
    sd = {name + suffix: _get(model, name + suffix).detach() for name in _HOT for suffix in ('.weight', '.bias')}
    model._mvx_path = PointPath(sd, grid, imsize_hw, eps, device=dev)
    model._mvx_imsize = {}
    model.forward = types.MethodType(_forward, model)
    return model


def install_extension():
    """The `modules/Extension.py` seam (Extension.py:1-3): point every already-imported user of the reference's `cpp` object
    (modules/data/Preprocessing.py:3, modules/Calc.py:4, modules/augment/Augment.py:6) at the CUDA implementation. Returns
    the names of the modules that were patched. A maintainer editing the reference writes instead, in modules/Extension.py:
    `from mvxnet_makise_b200.voxelize import cpp`."""
    from .voxelize import cpp
    patched = []
    for name in ('modules.Extension', 'modules.data.Preprocessing', 'modules.Calc', 'modules.augment.Augment'):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, 'cpp'):
            mod.cpp = cpp
            patched.append(name)
    return patched
