"""Generates tests/golden/cml_a.npz by running the UNMODIFIED reference's `CML` (modules/voxelnet/Pipe.py:31-43, built from
`CRB3d`, modules/layers/Blocks.py:20-29) on a small sparse grid. Pins oracle.cml_conv1 / oracle.crb3d, the checker of the sparse
hand-off kernel (csrc/sparse_conv.cu). Build-container only (needs /root/reference).
Run from the repo root:   python tests/golden/make_golden_cml.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    m = refshim.load()
    torch.manual_seed(11)
    cml = m.vpipe.CML()                                    # conv1 CRB3d(128,64,3,(2,1,1),(1,1,1)), conv2 CRB3d(64,64,3,1,(0,1,1)), conv3 CRB3d(64,64,3,(2,1,1),1)
    rng = np.random.default_rng(4)
    nz, nx, ny = 10, 24, 32
    G = nz * nx * ny
    cells = np.sort(rng.choice(G, 300, replace=False))     # ~4 % occupied, like a KITTI frame's 1.5 %
    feats = rng.standard_normal((300, 128)).astype(np.float32)
    grid = torch.zeros((1, 128, nz, nx, ny))
    iz, rem = np.divmod(cells, nx * ny)
    ix, iy = np.divmod(rem, ny)
    grid[0, :, iz, ix, iy] = torch.from_numpy(feats).T
    with torch.no_grad():
        y1 = cml.conv1(grid)
        y2 = cml.conv2(y1)
        y3 = cml.conv3(y2)
    out = dict(cells=cells.astype(np.int64), feats=feats, shape=np.array([nz, nx, ny]),
               y1=y1.numpy(), y2=y2.numpy(), y3=y3.numpy())
    for i, c in enumerate((cml.conv1, cml.conv2, cml.conv3), 1):
        out[f'w{i}'] = c.conv.weight.detach().numpy()
        out[f'b{i}'] = c.conv.bias.detach().numpy()
    np.savez_compressed(os.path.join(OUT, 'cml_a.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
