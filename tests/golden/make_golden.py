"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Build-container only (needs /root/reference; see oracle/refshim.py for the three import shims).
Run from the repo root:   python tests/golden/make_golden.py
The fixtures pin oracle/pointpath_oracle.py (tests/test_oracle.py) and, through it, the CUDA path.
Seeded inputs that are regenerated at test time (weights, FPN maps) carry a sha256 in the fixture.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SMALL_FPN = [(13, 42), (7, 21), (4, 11)]      # spatially reduced FPN levels (featureMaping is shape-agnostic)


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def small_maps(seed: int, shapes=SMALL_FPN):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((1, 256, h, w), dtype=np.float32) for (h, w) in shapes]


def voxel_case_points(seed: int, P: int, dense_frac: float = 0.3) -> np.ndarray:
    """Synthetic frame + a dense cluster (forces > T points per voxel) + exact cell-boundary floats
    and their fp32 neighbours (SURVEY.md trap 1) + duplicated points."""
    rng = np.random.default_rng(seed)
    base = synth.make_points(seed, P)
    n_dense = int(P * dense_frac)
    centers = base[rng.integers(0, P, 12)]
    dense = centers[rng.integers(0, 12, n_dense)].copy()
    dense[:, :3] += rng.normal(0, 0.12, (n_dense, 3)).astype(np.float32)
    # boundary candidates: k*size + low and +-1 ulp
    r, s = synth.KITTI_VELORANGE, synth.KITTI_GRID.voxelsize
    bx = (np.arange(1, 352, 7) * s[0] + r[0]).astype(np.float32)
    by = (np.arange(1, 400, 9) * s[1] + r[1]).astype(np.float32)
    bz = (np.arange(1, 10) * s[2] + r[2]).astype(np.float32)
    bnd = []
    for d, vals in enumerate((bx, by, bz)):
        for v in vals:
            for w in (np.nextafter(v, np.float32(-1e9)), v, np.nextafter(v, np.float32(1e9))):
                p = base[rng.integers(0, P)].copy()
                p[d] = w
                bnd.append(p)
    bnd = np.array(bnd, dtype=np.float32)
    dup = base[:50].copy()
    pts = np.concatenate([base, dense, bnd, dup], axis=0)
    pts = synth._crop(pts, r)                                   # keep the cropdata.py invariants
    pts = synth._crop_to_sight(pts, synth.kitti_calib(), (synth.KITTI_IMSIZE_HW[1], synth.KITTI_IMSIZE_HW[0]))
    pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts, dtype=np.float32)


def run_ref_group9(m, pcd6):
    """numba `group` with its in-function shuffle disabled (SURVEY.md trap 3)."""
    import numpy
    orig = numpy.random.shuffle
    numpy.random.shuffle = lambda a: None
    try:
        v, u = m.pre.group.py_func(pcd6.copy(), list(m.cfg.velorange), list(m.cfg.voxelsize), m.cfg.samplenum)
    finally:
        numpy.random.shuffle = orig
    return v, u


def main():
    m = refshim.load()
    torch.manual_seed(0)
    calib_np = synth.kitti_calib()
    calib_t = {k: torch.Tensor(v) for k, v in calib_np.items()}

    # ---------------------------------------------------------------- voxelization + projection
    for tag, seed, P in (('vox_a', 3, 2500), ('vox_b', 4, 6000)):
        pcd4 = voxel_case_points(seed, P)
        proj = m.calib.lidar2Img(torch.Tensor(pcd4), calib_t, True)          # Calib.py:47-70
        pcd6 = torch.concat([torch.Tensor(pcd4), proj[:, [1, 0]]], dim=1).numpy()   # train.py:32-35
        low = np.array(m.cfg.velorange[0:3])
        idx = ((pcd4[:, :3] - low) / m.cfg.voxelsize).astype('int32')         # Preprocessing.py:67-69
        vox7, uidx7, cnt7 = m.cpp._group(pcd4, idx, m.cfg.samplenum)          # voxelutil.cpp:325-360
        vox9, uidx9 = run_ref_group9(m, pcd6)                                 # Preprocessing.py:75-116
        assert np.array_equal(np.array(uidx7).T, uidx9.astype(np.int64))
        np.savez_compressed(os.path.join(OUT, f'{tag}.npz'), pcd4=pcd4, proj_uv=proj.numpy(), idx=idx,
                            vox7=vox7, uidx7=np.array(uidx7).T, cnt7=cnt7, vox9=vox9, uidx9=uidx9)
        print(tag, 'P', pcd4.shape[0], 'V', vox7.shape[0], 'capped', int((cnt7 == m.cfg.samplenum).sum()),
              'maxcnt', int(cnt7.max()))

    # ---------------------------------------------------------------- gather + layer stack + scatter
    for tag, seed, P in (('path_a', 5, 1500), ('path_b', 6, 4000)):
        pcd4 = voxel_case_points(seed, P, dense_frac=0.5)
        proj = m.calib.lidar2Img(torch.Tensor(pcd4), calib_t, True)
        pcd6 = torch.concat([torch.Tensor(pcd4), proj[:, [1, 0]]], dim=1).numpy()
        vox9, uidx9 = run_ref_group9(m, pcd6)
        voxels = torch.Tensor(vox9)[None]                                     # train.py:118,125
        idx = torch.LongTensor(np.concatenate([np.zeros((uidx9.shape[0], 1)), uidx9], axis=1))
        maps = small_maps(seed + 100)
        sd_np = synth.make_weights(seed)
        imsize = torch.Tensor(m.cfg.imsize)

        fusion = m.imhead_pipe.ImageFeatureFusion()
        svfe = m.vpipe.SVFE(m.cfg.samplenum)
        fcn = m.layers.FCN(128, 128)
        sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
        fusion.load_state_dict({k[len('head.fusion.'):]: v for k, v in sd.items() if k.startswith('head.fusion.')})
        svfe.load_state_dict({k[len('backbone.svfe.'):]: v for k, v in sd.items() if k.startswith('backbone.svfe.')})
        fcn.load_state_dict({k[len('backbone.fcn.'):]: v for k, v in sd.items() if k.startswith('backbone.fcn.')})
        with torch.no_grad():
            feats = [torch.from_numpy(x.copy()) for x in maps]
            im768 = m.imhead_pipe.featureMaping(voxels, feats, [calib_t], imsize)[0]   # mutates voxels
            im16 = fusion(im768[None])
            x23 = torch.concat([voxels[..., :7], im16], dim=-1)               # MVXNet.py:26
            x = svfe(x23)
            x = fcn(x)
            x = torch.max(x, dim=2)[0]
            vfeat = torch.squeeze(x, dim=2).reshape((-1, 128))
            grid = m.VoxelNet.reindex(vfeat, idx)
        real = np.flatnonzero((voxels[0, ..., :3] != 0).any(-1).reshape(-1).numpy())
        rows = np.sort(np.random.default_rng(seed).choice(real, 96, replace=False))
        np.savez_compressed(
            os.path.join(OUT, f'{tag}.npz'), pcd4=pcd4, map_seed=seed + 100, weight_seed=seed,
            maps_sha=sha(*maps), weights_sha=sha(*[sd_np[k] for k in sorted(sd_np)]),
            voxels9_after=voxels[0].numpy(), idx=idx.numpy(),
            im768_rows=rows, im768_sample=im768.reshape(-1, 768)[rows].numpy(),
            im768_colsum=im768.reshape(-1, 768).double().sum(0).numpy(),
            im16=im16[0].numpy(), vfeat=vfeat.numpy(),
            grid_nonzero=int((grid != 0).sum()), grid_sha=sha(grid.numpy()), grid_shape=np.array(grid.shape))
        print(tag, 'P', pcd4.shape[0], 'N', vox9.shape[0], 'grid nz', int((grid != 0).sum()))


if __name__ == '__main__':
    main()
