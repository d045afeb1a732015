"""The fused batched point path (stages 1-4) behind one call — what ``MVXNet.forward`` does between the image
backbone and ``VoxelNet.cml`` (MVXNet.py:21-27; Head.py:14-22; VoxelNet.py:24-34), with the CPU half of
train.py:26-49 (projection + voxelization) moved onto the GPU.

    path = PointPath(state_dict)                      # reference parameter names
    grid, counts = path(points_list, calibs, fpn_maps)   # (B,128,nz,nx,ny) fp32 on the GPU, counts (B,4) int32

Frames are independent (per-frame BatchNorm statistics, as in the batch-1 reference), so multi-GPU use is pure
frame sharding: see ``mvxnet_makise_b200.dist``.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib, synth
from ._lib import lib, check, ptr, stream_ptr
from .modules import _wt, pack_calib, pack_calib64, calib_is_f64

LAYER_NAMES = [name for name, *_ in synth.HOT_LAYERS]


class _HostSlot:
    """One complete set of host-entry buffers (sub-batch input buffers + workspaces, device outputs, pinned results)."""

    def __init__(self):
        self.subs, self.used = [], False
        self.grid_out = self.counts = self.out_counts = self.out_head = self.done = None


class HostStep:
    """Handle of one asynchronous `PointPath.forward_host(..., sync=False)` call."""

    def __init__(self, slot: _HostSlot):
        self._slot = slot
        self.event = slot.done

    def wait(self):
        """Blocks until this call's results are in host memory; returns (grid on device, counts on host, head block on host)."""
        self.event.synchronize()
        return self._slot.grid_out, self._slot.out_counts, self._slot.out_head


class PointPath:
    def __init__(self, state_dict: Dict[str, torch.Tensor], grid: synth.GridSpec = synth.KITTI_GRID,
                 imsize_hw: Sequence[int] = synth.KITTI_IMSIZE_HW, eps: float = 1e-6, device='cuda'):
        _lib.require_cuda()
        self.device = torch.device(device)
        self.grid_spec = grid
        self.grid = _lib.make_grid(grid.velorange, grid.voxelsize, grid.voxelshape, grid.T)
        self.imsize_hw = (float(imsize_hw[0]), float(imsize_hw[1]))
        self.eps = float(eps)
        self.load_state_dict(state_dict)
        self._ws = None
        self._ws_key = None
        self._args = None
        self._subs = None          # sub-batch contexts of the pipelined host entry (forward_host)
        self._sub_chunk = 0
        self.host_chunk = 1        # frames per sub-batch of forward_host: H2D of chunk j+1 overlaps compute of chunk j
        self.host_streams = 2      # sub-batches alternate between this many compute streams (their kernels may overlap)
        self.host_taper = True     # split the last sub-batch in two: less compute left after the last copy has landed
        self.host_slots = 2        # complete buffer sets of forward_host, used alternately: call s+1's copies overlap call s's kernels
        # The reference shuffles the points INSIDE group / group_ (Preprocessing.py:66,86), so "the first T points of a voxel"
        # is a random T-subset, redrawn every epoch. None: shuffle in training mode (forward_train), keep the caller's order in
        # inference; True / False force it. `shuffle_generator` seeds it (torch.Generator on the device); the permutation
        # applied by the last call is kept in `last_perm` (global row indices into the call's point array).
        self.shuffle = None
        self.shuffle_generator = None
        self.last_perm = None

    def load_state_dict(self, sd):
        self.wt, self.bias = [], []
        for name in LAYER_NAMES:
            w = torch.as_tensor(np.asarray(sd[name + '.weight']) if not isinstance(sd[name + '.weight'], torch.Tensor)
                                else sd[name + '.weight'])
            b = torch.as_tensor(np.asarray(sd[name + '.bias']) if not isinstance(sd[name + '.bias'], torch.Tensor)
                                else sd[name + '.bias'])
            self.wt.append(_wt(w.to(self.device)))
            self.bias.append(b.detach().to(self.device, torch.float32).contiguous())

    # ------------------------------------------------------------------------------------------------
    def _prepare(self, B: int, cap: int, map_hw, own_outputs: bool = True, train: bool = False):
        key = (B, cap, tuple(map_hw), own_outputs, train)
        if self._ws_key == key:
            return
        a = _lib.PointPathArgs()
        a.grid = self.grid
        a.B, a.cap = B, cap
        for l, (h, w) in enumerate(map_hw):
            a.map_h[l], a.map_w[l] = h, w
        a.map_c = 256
        nbytes = ctypes.c_size_t()
        check(lib.mvx_pointpath_workspace_bytes(ctypes.byref(a), ctypes.byref(nbytes)), 'pointpath_workspace_bytes')
        offs = (ctypes.c_int64 * _lib.WS_REGIONS)()
        check(lib.mvx_pointpath_layout(ctypes.byref(a), offs), 'pointpath_layout')
        self.layout = {}
        for r in range(_lib.WS_REGIONS):
            n = lib.mvx_pointpath_layout_name(r)
            if n:
                self.layout[n.decode()] = int(offs[r])
        self._bws = None
        if train:     # training workspace: + the raw output of the last FCN; backward scratch (gradient buffers)
            fwd_b, bwd_b = ctypes.c_size_t(), ctypes.c_size_t()
            check(lib.mvx_pointpath_train_workspace_bytes(ctypes.byref(a), ctypes.byref(fwd_b), ctypes.byref(bwd_b)),
                  'pointpath_train_workspace_bytes')
            nbytes = fwd_b
            self._bws = torch.empty(bwd_b.value, dtype=torch.uint8, device=self.device)
        self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        self._ws_key = key
        self.cap, self.B = cap, B
        if own_outputs:
            self._alloc_outputs(B)

    def _alloc_outputs(self, B: int):
        nz, nx, ny = self.grid.shape[2], self.grid.shape[0], self.grid.shape[1]
        if getattr(self, 'grid_out', None) is None or self.grid_out.shape[0] != B:
            self.grid_out = torch.empty((B, 128, nz, nx, ny), dtype=torch.float32, device=self.device)
            self.counts = torch.empty((B, 4), dtype=torch.int32, device=self.device)

    def region(self, name: str, dtype, shape):
        """Typed view into a named workspace region (tests / diagnostics)."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        off = self.layout[name]
        return self._ws[off:off + n].view(dtype).view(*shape)

    # ------------------------------------------------------------------------------------------------
    def forward_device(self, points: torch.Tensor, offsets: Sequence[int], calib32: torch.Tensor,
                       maps: List[torch.Tensor], want_grid: bool = True, cap: int | None = None,
                       grid_out: torch.Tensor | None = None, counts: torch.Tensor | None = None, train: bool = False,
                       point_calib: torch.Tensor | None = None, shuffle: bool | None = None,
                       calib64: torch.Tensor | None = None, calib_f64: torch.Tensor | None = None):
        """points (sum P, stride>=4) fp32 CUDA, offsets host [B+1], calib32 (B,32) CUDA, maps 3 x (B,256,Hf,Wf) CUDA.
        grid_out / counts: optional caller-owned outputs ((B,128,nz,nx,ny) fp32, (B,4) int32, contiguous).
        train=True keeps every activation for `backward` (row-first fcn1; mvx_pointpath_forward_train).
        point_calib: optional (sum P) int32 CUDA — merged point sets (GT-paste, train.py:29-42): calibration set of every
        point; calib32 is then the (n_sets, 32) table it indexes.
        calib64 (n_sets, 32) float64 CUDA + calib_f64 (n_sets) int32 CUDA: calibration sets flagged 1 are projected in fp64
        (the reference's numpy lidar2Img with readCalib's float64 matrices, train.py:36-39) and rounded to fp32 once.
        shuffle: permute every frame's points on the device first, like the in-function shuffle of the reference's `group`
        (None: `self.shuffle`, which defaults to training mode only)."""
        B = len(offsets) - 1
        if shuffle is None:
            shuffle = getattr(self, 'shuffle', None)
        if shuffle is None:
            shuffle = train
        self.last_perm = None
        if shuffle and points.shape[0] > 0:
            gen = getattr(self, 'shuffle_generator', None)
            perm = torch.cat([torch.randperm(int(offsets[f + 1]) - int(offsets[f]), device=points.device, generator=gen) + int(offsets[f])
                              for f in range(B)])
            points = points.index_select(0, perm)
            if point_calib is not None:
                point_calib = point_calib.index_select(0, perm).contiguous()
            self.last_perm = perm
        maxp = max(offsets[i + 1] - offsets[i] for i in range(B))
        cap = cap or max(128, (maxp + 127) // 128 * 128)
        map_hw = [(int(m.shape[-2]), int(m.shape[-1])) for m in maps]
        self._prepare(B, cap, map_hw, own_outputs=counts is None, train=train)
        if counts is not None:
            assert counts.is_contiguous() and counts.shape == (B, 4) and (grid_out is None or grid_out.is_contiguous())
            self.grid_out, self.counts = grid_out, counts
        self._subs_active = False
        a = _lib.PointPathArgs()
        a.grid = self.grid
        a.B, a.cap = B, cap
        if points.shape[0] == 0:     # all-empty batch: the kernels read no point, but the ABI wants a valid pointer
            points = torch.zeros((1, max(4, points.shape[1])), dtype=torch.float32, device=self.device)
        a.points, a.point_stride = points.data_ptr(), points.shape[1]
        off = (ctypes.c_int32 * (B + 1))(*[int(o) for o in offsets])
        a.pt_off_host = off
        a.calib32 = calib32.data_ptr()
        if point_calib is not None:
            assert point_calib.is_cuda and point_calib.dtype == torch.int32 and point_calib.is_contiguous()
            assert point_calib.numel() == offsets[-1] and calib32.dim() == 2 and calib32.shape[1] == 32
            a.point_calib = point_calib.data_ptr()
        if calib64 is not None:
            assert calib64.is_cuda and calib64.dtype == torch.float64 and calib64.is_contiguous() and calib64.shape == calib32.shape
            assert calib_f64 is not None and calib_f64.is_cuda and calib_f64.dtype == torch.int32 and calib_f64.numel() == calib32.shape[0]
            a.calib64, a.calib_f64 = calib64.data_ptr(), calib_f64.data_ptr()
        for l in range(3):
            assert maps[l].is_contiguous() and maps[l].shape[0] == B and maps[l].shape[1] == 256
            a.maps[l] = maps[l].data_ptr()
            a.map_h[l], a.map_w[l] = map_hw[l]
        a.map_c = 256
        a.imsize_h, a.imsize_w = self.imsize_hw
        a.gather_eps, a.bn_eps = self.eps, self.eps
        for l in range(8):
            a.wt[l] = self.wt[l].data_ptr()
            a.bias[l] = self.bias[l].data_ptr()
        a.grid_out = self.grid_out.data_ptr() if want_grid else None
        a.counts = self.counts.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        a.stream = torch.cuda.current_stream().cuda_stream
        if train:
            check(lib.mvx_pointpath_forward_train(ctypes.byref(a)), 'pointpath_forward_train')
            self._train_args = (a, off, points, calib32, maps, point_calib, calib64, calib_f64)      # kept alive for backward()
        else:
            check(lib.mvx_pointpath_forward(ctypes.byref(a)), 'pointpath_forward')
            self._train_args = None
        self._last_args = (a, off, points, calib32, maps, point_calib, calib64, calib_f64)           # kept alive for cml_conv1()
        return (self.grid_out if want_grid else None), self.counts

    # ---- the reference's own arguments: MVXNet.forward(voxels, imgs, idx, calibs, imsize) (MVXNet.py:21-27) ----------------
    def forward_voxels(self, voxels, idx, maps: List[torch.Tensor], want_grid: bool = True, train: bool = False,
                       grid_out: torch.Tensor | None = None):
        """Dense-voxel entry: `voxels` = the (1, N, T, 9) / (N, T, 9) fp32 CUDA tensor [x,y,z,dx,dy,dz,r,row,col] that
        `pre.group` + train.py:118-128 produce, `idx` = (N, 4) int64 [batch, ix, iy, iz]; or lists of B such pairs (one per
        frame, per-frame BatchNorm statistics like the batch-1 reference). maps: 3 x (B,256,Hf,Wf) CUDA.
        Stage 1 and the projection are the caller's (the reference runs them on the CPU); stages 2b-4 run here. Pad slots
        (x == y == z == 0, Pipe.py:53-54) are zeroed IN PLACE in `voxels` like featureMaping does (Pipe.py:58-59).
        Costs one device->host read of the per-frame real-slot counts (the reference's featureMaping synchronises three
        times, Pipe.py:71). Returns (grid (B,128,nz,nx,ny), counts (B,4))."""
        orig = list(voxels) if isinstance(voxels, (list, tuple)) else [voxels]
        il = list(idx) if isinstance(idx, (list, tuple)) else [idx]
        assert len(orig) == len(il)
        T = int(self.grid.T)
        vl = [v.reshape(-1, T, 9) for v in orig]              # a view when the caller's tensor is contiguous
        for v, i in zip(vl, il):
            assert v.is_cuda and v.dtype == torch.float32 and i.is_cuda and i.dtype == torch.int64 and i.shape == (v.shape[0], 4)
        B = len(vl)
        single_inplace = B == 1 and vl[0].is_contiguous() and vl[0].data_ptr() == orig[0].data_ptr()
        vcat = vl[0] if single_inplace else torch.cat(vl, dim=0).contiguous()
        icat = il[0].contiguous() if B == 1 else torch.cat(il, dim=0).contiguous()
        voff = np.concatenate([[0], np.cumsum([v.shape[0] for v in vl])]).astype(np.int64).tolist()
        voff_c = (ctypes.c_int32 * (B + 1))(*[int(o) for o in voff])
        cnt = torch.empty((B, 4), dtype=torch.int32, device=self.device)
        check(lib.mvx_dense_voxel_counts(ptr(vcat), voff_c, B, T, ptr(cnt), stream_ptr()), 'dense_voxel_counts')
        c = cnt.cpu()
        need = max(int(c[:, 1].max()), int(c[:, 0].max()), 1)
        cap = (need + 4095) // 4096 * 4096                     # capacity bucket: no re-allocation for every new frame size
        map_hw = [(int(m.shape[-2]), int(m.shape[-1])) for m in maps]
        self._prepare(B, cap, map_hw, own_outputs=grid_out is None, train=train)
        if grid_out is not None:
            assert grid_out.is_contiguous() and grid_out.dtype == torch.float32
            self.grid_out = grid_out
            if getattr(self, 'counts', None) is None or self.counts.shape[0] != B:
                self.counts = torch.empty((B, 4), dtype=torch.int32, device=self.device)
        self._subs_active = False
        a = _lib.PointPathArgs()
        a.grid = self.grid
        a.B, a.cap = B, cap
        a.voxels_dense, a.voxel_idx, a.vox_off_host = vcat.data_ptr(), icat.data_ptr(), voff_c
        for l in range(3):
            assert maps[l].is_cuda and maps[l].is_contiguous() and maps[l].shape[0] == B and maps[l].shape[1] == 256
            a.maps[l] = maps[l].data_ptr()
            a.map_h[l], a.map_w[l] = map_hw[l]
        a.map_c = 256
        a.imsize_h, a.imsize_w = self.imsize_hw
        a.gather_eps, a.bn_eps = self.eps, self.eps
        for l in range(8):
            a.wt[l] = self.wt[l].data_ptr()
            a.bias[l] = self.bias[l].data_ptr()
        a.grid_out = self.grid_out.data_ptr() if want_grid else None
        a.counts = self.counts.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        a.stream = torch.cuda.current_stream().cuda_stream
        keep = (a, voff_c, vcat, icat, maps)
        if train:
            check(lib.mvx_pointpath_forward_train(ctypes.byref(a)), 'pointpath_forward_train')
            self._train_args = keep
        else:
            check(lib.mvx_pointpath_forward(ctypes.byref(a)), 'pointpath_forward')
            self._train_args = None
        self._last_args = keep
        if not single_inplace:                                 # the in-place side effect, for callers that handed in views / lists
            for f, v in enumerate(orig):
                v.copy_(vcat[voff[f]:voff[f + 1]].reshape(v.shape))
        return (self.grid_out if want_grid else None), self.counts

    # ---- after the path: sparse hand-off to CML.conv1 -------------------------------------
    def cml_conv1(self, weight: torch.Tensor, bias: torch.Tensor, eps: float | None = None) -> torch.Tensor:
        """`backbone.cml.conv1` = CRB3d(128, 64, 3, (2,1,1), (1,1,1)) (voxelnet/Pipe.py:31-43) of the dense grid of the LAST
        forward, computed sparsely from the voxel features (the forward may have run with want_grid=False).
        weight (64,128,3,3,3), bias (64) as in the reference checkpoint (`backbone.cml.conv1.conv.*`).
        Returns (B, 64, (nz+1)//2, nx, ny) fp32 on the GPU, per-frame BatchNorm statistics like the batch-1 reference."""
        if getattr(self, '_last_args', None) is None or getattr(self, '_subs_active', False):
            raise RuntimeError('cml_conv1() needs a preceding forward (device entry) on this PointPath')
        a = self._last_args[0]
        w = weight.detach().to(self.device, torch.float32).contiguous()
        b = bias.detach().to(self.device, torch.float32).contiguous()
        assert tuple(w.shape) == (64, 128, 3, 3, 3) and tuple(b.shape) == (64,)
        nz, nx, ny = self.grid.shape[2], self.grid.shape[0], self.grid.shape[1]
        nbytes = ctypes.c_size_t()
        check(lib.mvx_cml_conv1_workspace_bytes(ctypes.byref(a), ctypes.byref(nbytes)), 'cml_conv1_workspace_bytes')
        if getattr(self, '_cml_ws', None) is None or self._cml_ws.numel() < nbytes.value:
            self._cml_ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        out = torch.empty((self.B, 64, (nz + 2 - 3) // 2 + 1, nx, ny), dtype=torch.float32, device=self.device)
        a.stream = torch.cuda.current_stream().cuda_stream
        check(lib.mvx_cml_conv1_sparse(ctypes.byref(a), ptr(w), ptr(b), float(self.eps if eps is None else eps), ptr(out),
                                       ptr(self._cml_ws), self._cml_ws.numel()), 'cml_conv1_sparse')
        return out

    # ---- training mode -------------------------------------------------------------------------------------
    def forward_train(self, points, offsets, calib32, maps, want_grid: bool = True, cap: int | None = None,
                      shuffle: bool | None = None):
        """Forward that keeps the activations the backward needs (BASELINE.json configs[3]). Shuffles every frame's points
        first unless `shuffle=False` / `self.shuffle = False` (the reference's `group` does, Preprocessing.py:86)."""
        return self.forward_device(points, offsets, calib32, maps, want_grid, cap, train=True, shuffle=shuffle)

    @staticmethod
    def grad_layout():
        """(name, offset, shape) of every parameter gradient inside the flat bucket: checkpoint names/order."""
        out, o = [], 0
        for name, cin, cout, is_conv in synth.HOT_LAYERS:
            out.append((name + '.weight', o, (cout, cin, 1, 1) if is_conv else (cout, cin)))
            o += cout * cin
            out.append((name + '.bias', o, (cout,)))
            o += cout
        return out

    def backward(self, d_vfeat: torch.Tensor | None = None, d_grid: torch.Tensor | None = None,
                 grad_flat: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
        """Gradients of the 8 hot-path layers for the last forward_train call, summed over its frames, as ONE flat fp32
        bucket (the message of the NCCL all-reduce; `grad_layout()` names its slices). Give dLoss/d voxel features
        (B, cap, 128) or dLoss/d dense grid (B,128,nz,nx,ny)."""
        if getattr(self, '_train_args', None) is None:
            raise RuntimeError('backward() needs a preceding forward_train() on this PointPath')
        a = self._train_args[0]
        n = int(lib.mvx_grad_floats())
        if grad_flat is None:
            grad_flat = torch.empty(n, dtype=torch.float32, device=self.device)
            accumulate = False
        assert grad_flat.is_contiguous() and grad_flat.numel() == n and grad_flat.dtype == torch.float32
        for t, shape in ((d_vfeat, (self.B, self.cap, 128)), (d_grid, tuple(self.grid_out.shape) if self.grid_out is not None else None)):
            if t is not None:
                assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and tuple(t.shape) == shape, (tuple(t.shape), shape)
        a.stream = torch.cuda.current_stream().cuda_stream
        check(lib.mvx_pointpath_backward(ctypes.byref(a), ptr(d_vfeat), ptr(d_grid), grad_flat.data_ptr(), int(bool(accumulate)),
                                         self._bws.data_ptr(), self._bws.numel()), 'pointpath_backward')
        return grad_flat

    def grads(self, grad_flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Named views into a flat gradient bucket (reference state-dict names)."""
        return {name: grad_flat[o:o + int(np.prod(shape))].view(*shape) for name, o, shape in self.grad_layout()}

    def forward_device_split(self, points: torch.Tensor, offsets: Sequence[int], calib32: torch.Tensor,
                             maps: List[torch.Tensor], want_grid: bool = True, n_split: int = 2):
        """Same inputs/outputs as forward_device, but the batch is cut into `n_split` sub-batches that run CONCURRENTLY on
        their own streams and workspaces (frames are independent): the memory-bound kernels of one sub-batch (grid fill,
        combine) overlap the tensor-core kernels of the other. Joins back into the current stream."""
        B = len(offsets) - 1
        n_split = max(1, min(int(n_split), B))
        if n_split == 1:
            return self.forward_device(points, offsets, calib32, maps, want_grid)
        key = (B, n_split)
        if getattr(self, '_split_key', None) != key:
            bounds = [round(i * B / n_split) for i in range(n_split + 1)]
            self._split = []
            for i in range(n_split):
                c = self._child()
                c.f0, c.f1 = bounds[i], bounds[i + 1]
                c.stream = torch.cuda.Stream(device=self.device)
                self._split.append(c)
            self._split_key = key
        self._alloc_outputs(B)
        self.B = B
        cur = torch.cuda.current_stream()
        for c in self._split:
            p0, p1 = int(offsets[c.f0]), int(offsets[c.f1])
            c.stream.wait_stream(cur)
            with torch.cuda.stream(c.stream):
                c.forward_device(points[p0:p1], [int(o) - p0 for o in offsets[c.f0:c.f1 + 1]], calib32[c.f0:c.f1],
                                 [m[c.f0:c.f1] for m in maps], want_grid,
                                 grid_out=self.grid_out[c.f0:c.f1] if want_grid else None, counts=self.counts[c.f0:c.f1])
        for c in self._split:
            cur.wait_stream(c.stream)
        self._subs, self._sub_chunk, self._subs_active = self._split, None, True
        return (self.grid_out if want_grid else None), self.counts

    def __call__(self, points_list: List, calibs: List[dict], fpn_maps: List, want_grid: bool = True,
                 shuffle: bool | None = None):
        """points_list: B arrays/tensors (P_f, >=4) [x,y,z,r]; calibs: B dicts of 4x4 matrices (Load.py:24-41);
        fpn_maps: 3 tensors (B,256,Hf,Wf) (FPN levels '0','1','2').
        A frame may also be a LIST of point sets with a LIST of calibration dicts, one per set — the scene followed by the
        pasted ground-truth objects of the GT-paste augmentation, each projected through its own calibration and then
        merged in that order (train.py:29-42)."""
        def as_t(p):
            return torch.as_tensor(np.asarray(p, dtype=np.float32)) if not isinstance(p, torch.Tensor) else p
        merged = any(isinstance(c, (list, tuple)) for c in calibs)
        pts, table, pc, table64, is64 = [], [], [], [], []
        for p, c in zip(points_list, calibs):
            sets, cals = (list(p), list(c)) if isinstance(c, (list, tuple)) else ([p], [c])
            assert len(sets) == len(cals), 'one calibration per point set'
            sets = [as_t(q) for q in sets]
            pts.append(torch.cat([q[:, :4].to(torch.float32) for q in sets], dim=0))
            if merged:
                for k, (q, cal) in enumerate(zip(sets, cals)):
                    pc.append(np.full(q.shape[0], len(table), dtype=np.int32))
                    table.append(pack_calib(cal))
                    table64.append(pack_calib64(cal))
                    # train.py:31-33 projects the scene (set 0) with torch in fp32 whatever the dict holds; the pasted sets go
                    # through the numpy branch (train.py:36-39), which is fp64 when their dicts are readCalib's float64 ones
                    is64.append(1 if (k > 0 and calib_is_f64(cal)) else 0)
            else:
                table.append(pack_calib(cals[0]))
        offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in pts])]).tolist()
        points = torch.cat(pts, dim=0).contiguous().to(self.device, non_blocking=True)
        calib32 = torch.stack(table).to(self.device, non_blocking=True)
        point_calib = torch.from_numpy(np.concatenate(pc)).to(self.device, non_blocking=True) if merged else None
        c64 = f64 = None
        if merged and any(is64):
            c64 = torch.stack(table64).to(self.device)
            f64 = torch.tensor(is64, dtype=torch.int32, device=self.device)
        maps = [torch.as_tensor(m).to(self.device, torch.float32).contiguous() for m in fpn_maps]
        return self.forward_device(points, offsets, calib32, maps, want_grid, point_calib=point_calib, shuffle=shuffle,
                                   calib64=c64, calib_f64=f64)

    # ---- host-buffer entry: H2D of this batch's inputs, the fused path, D2H of the counts ---------------
    def _child(self):
        c = object.__new__(PointPath)
        c.device, c.grid_spec, c.grid, c.imsize_hw, c.eps = self.device, self.grid_spec, self.grid, self.imsize_hw, self.eps
        c.wt, c.bias = self.wt, self.bias          # shared weights
        c._ws = c._ws_key = c._args = c._subs = None
        c._sub_chunk, c.host_chunk, c.host_taper = 0, 0, False
        c.shuffle, c.shuffle_generator, c.last_perm = self.shuffle, self.shuffle_generator, None
        return c

    def forward_host(self, points_host: torch.Tensor, offsets: Sequence[int], calib32_host: torch.Tensor,
                     maps_host: List[torch.Tensor], want_grid: bool = True, head_rows: int = 1024, sync: bool = True):
        """Inputs in (ideally pinned) HOST memory: points (sum P, 4) fp32, calib32 (B,32), maps 3 x (B,256,Hf,Wf). The maps may
        also be CUDA tensors (the reference produces them on the GPU, Head.py:14-22): then only points + calib cross PCIe.
        Frames are independent, so the batch is cut into sub-batches of `self.host_chunk` frames: the H2D copy of
        sub-batch j+1 (copy stream) overlaps the kernels of sub-batch j (compute streams); each sub-batch has its own
        input buffers and workspace and writes its slice of the batch outputs. Reads back the per-frame counts and
        the first `head_rows` voxel feature rows of frame 0 (the host-visible result).

        Steps are pipelined ACROSS calls as well: there are `self.host_slots` (2) complete sets of input buffers,
        workspaces and outputs, used alternately, and the copies of call s+1 wait only for the kernels that last read
        the same slot (call s-1), not for call s - so while call s computes, the copy engine already moves call s+1's
        inputs. sync=True (default) waits for this call's results and returns (grid on device, counts on host, head block
        on host); sync=False returns a `HostStep` handle at once: `.wait()` blocks until the results of THAT call are in
        its pinned buffers and returns the same triple (valid until the slot is reused two calls later)."""
        dev = self.device
        B = len(offsets) - 1
        # sub-batch schedule: `host_chunk` frames per sub-batch, or an explicit list of sizes; the default tapers the tail
        # (..., 1, 1): the step ends one sub-batch of compute after the LAST copy lands, so the last sub-batch should be small
        if isinstance(self.host_chunk, (list, tuple)):
            sizes = [int(c) for c in self.host_chunk]
        else:
            chunk = max(1, min(int(self.host_chunk) or B, B))
            sizes = [chunk] * (B // chunk) + ([B % chunk] if B % chunk else [])
            if self.host_taper and chunk > 1 and len(sizes) >= 2 and sizes[-1] == chunk:
                sizes = sizes[:-1] + [chunk - chunk // 2, chunk // 2] if chunk > 2 else sizes[:-1] + [1, 1]
        assert sum(sizes) == B and all(c > 0 for c in sizes), sizes
        bounds = [0]
        for c in sizes:
            bounds.append(bounds[-1] + c)
        # contexts are keyed on capacity BUCKETS, not on the exact per-frame point counts: real sweeps differ in size from
        # call to call, and rebuilding the sub-batch contexts would re-allocate multi-GB workspaces and pinned buffers
        maxp = max([int(offsets[i + 1]) - int(offsets[i]) for i in range(B)] + [1])
        cap = (maxp + 4095) // 4096 * 4096
        stride = int(points_host.shape[1])
        n_slots = max(1, int(getattr(self, 'host_slots', 2)))
        key = (B, cap, stride, tuple(tuple(m.shape[1:]) + (m.is_cuda,) for m in maps_host), tuple(sizes), int(head_rows), n_slots)
        if getattr(self, '_in_key', None) != key:
            self._slots = []
            for _ in range(n_slots):
                slot = _HostSlot()
                for f0, f1 in zip(bounds[:-1], bounds[1:]):
                    c = self._child()
                    c.f0, c.f1 = f0, f1
                    c.in_points = torch.empty(((f1 - f0) * cap, stride), dtype=torch.float32, device=dev)
                    c.in_calib = torch.empty((f1 - f0, 32), dtype=torch.float32, device=dev)
                    c.in_maps = [None if m.is_cuda else torch.empty((f1 - f0,) + tuple(m.shape[1:]), dtype=torch.float32, device=dev) for m in maps_host]
                    c.ev = torch.cuda.Event()
                    slot.subs.append(c)
                nz, nx, ny = self.grid.shape[2], self.grid.shape[0], self.grid.shape[1]
                slot.grid_out = torch.empty((B, 128, nz, nx, ny), dtype=torch.float32, device=dev) if want_grid else None
                slot.counts = torch.empty((B, 4), dtype=torch.int32, device=dev)
                slot.out_counts = torch.empty((B, 4), dtype=torch.int32).pin_memory()
                slot.out_head = torch.empty((min(int(head_rows), cap), 128), dtype=torch.float32).pin_memory()
                slot.done = torch.cuda.Event()
                self._slots.append(slot)
            self._copy_stream = torch.cuda.Stream(device=dev)
            ns = max(1, min(int(self.host_streams), len(sizes)))
            self._compute_streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]
            self._host_calls = 0
            self._in_key = key
        slot = self._slots[self._host_calls % len(self._slots)]
        self._host_calls += 1
        if want_grid and slot.grid_out is None:
            nz, nx, ny = self.grid.shape[2], self.grid.shape[0], self.grid.shape[1]
            slot.grid_out = torch.empty((B, 128, nz, nx, ny), dtype=torch.float32, device=dev)
        self._h2d_bytes = (points_host.numel() + calib32_host.numel() + sum(m.numel() for m in maps_host if not m.is_cuda)) * 4
        self.B, self.grid_out, self.counts = B, slot.grid_out, slot.counts
        cur = torch.cuda.current_stream()
        cs = self._copy_stream
        ns = len(self._compute_streams)
        if slot.used:
            cs.wait_event(slot.done)             # the kernels of the call that last used this slot have read its inputs
        with torch.cuda.stream(cs):              # every copy is queued up front: the copy engine never waits for the CPU
            for c in slot.subs:
                p0, p1 = int(offsets[c.f0]), int(offsets[c.f1])
                c.offsets = [int(o) - p0 for o in offsets[c.f0:c.f1 + 1]]
                c.n_points = p1 - p0
                if p1 > p0:
                    c.in_points[:p1 - p0].copy_(points_host[p0:p1], non_blocking=True)
                c.in_calib.copy_(calib32_host[c.f0:c.f1], non_blocking=True)
                for d, h in zip(c.in_maps, maps_host):
                    if d is not None:
                        d.copy_(h[c.f0:c.f1], non_blocking=True)
                c.ev.record(cs)
        for j, c in enumerate(slot.subs):
            ks = self._compute_streams[j % ns]
            if j < ns:
                ks.wait_stream(cur)              # work the caller queued before this call (weight updates ...) is ordered first
            with torch.cuda.stream(ks):
                ks.wait_event(c.ev)
                sub_maps = [d if d is not None else h[c.f0:c.f1] for d, h in zip(c.in_maps, maps_host)]   # device-resident maps: a view
                c.forward_device(c.in_points[:c.n_points], c.offsets, c.in_calib, sub_maps, want_grid, cap=cap,
                                 grid_out=slot.grid_out[c.f0:c.f1] if want_grid else None, counts=slot.counts[c.f0:c.f1])
        k0 = self._compute_streams[0]
        for s_ in self._compute_streams[1:]:
            k0.wait_stream(s_)
        with torch.cuda.stream(k0):              # device -> host read of the step's result
            slot.out_counts.copy_(slot.counts, non_blocking=True)
            c0 = slot.subs[0]
            n_head = slot.out_head.shape[0]
            slot.out_head.copy_(c0.region('vfeat', torch.float32, (c0.B, c0.cap, 128))[0, :n_head], non_blocking=True)
            slot.done.record(k0)
        slot.used = True
        cur.wait_event(slot.done)                # later work on the caller's stream sees this call's device outputs
        self._subs, self._sub_chunk, self._subs_active = slot.subs, None, True
        self._d2h_bytes = slot.out_counts.numel() * 4 + slot.out_head.numel() * 4
        step = HostStep(slot)
        return step.wait() if sync else step

    @property
    def h2d_bytes(self):
        return self._h2d_bytes

    @property
    def d2h_bytes(self):
        return self._d2h_bytes

    # ---- compact outputs (reference voxel order), for callers that do not want the dense grid -----------
    def voxel_features(self, f: int):
        """(N_f,128) fp32 features and (N_f,4) int64 idx [batch, ix, iy, iz] of frame f (after a forward)."""
        n = int(self.counts[f, 0].item())
        src, fl = self, f
        if getattr(self, '_subs_active', False):   # the last forward ran in sub-batches: frame f lives in one of them
            src = next(c for c in self._subs if c.f0 <= f < c.f1)
            fl = f - src.f0
        vfeat = src.region('vfeat', torch.float32, (src.B, src.cap, 128))[fl, :n]
        coord = src.region('vox_coord', torch.int32, (src.B, src.cap, 4))[fl, :n, :3].to(torch.int64)
        idx = torch.cat([torch.full((n, 1), f, dtype=torch.int64, device=coord.device), coord], dim=1)
        return vfeat, idx
