"""Dev tool (CPU): prints the agreement of the COMPACT backward formulas (oracle/compact_backward.py: weighted pad rows,
argmax routing, BatchNorm batch-stat backward with multiplicities) with torch autograd through the dense oracle chain.
The same check runs as a test (tests/test_oracle.py::test_compact_backward_equals_dense_autograd)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth
from oracle import pointpath_oracle as O
from oracle import compact_backward as CB

G = synth.KITTI_GRID


def main(P=1200, seed=11):
    pts = synth.make_points(seed, P)
    maps = [np.random.default_rng(1).standard_normal((1, 256, h, w), dtype=np.float32) for (h, w) in ((13, 42), (7, 21), (4, 11))]
    sd_np = synth.make_weights(4)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    idx = O.cell_index(pts, G.velorange, G.voxelsize)
    N = O.group_assign(idx, G.T)[2].shape[0]
    Gw = np.random.default_rng(2).standard_normal((N, 128))
    vf, gref = O.backward_frame(pts, synth.kitti_calib(), maps, sd_np, G, synth.KITTI_IMSIZE_HW, Gw)
    pcd6 = O.points_with_proj(pts, synth.kitti_calib())
    voxels = torch.Tensor(O.group(pcd6, G.velorange, G.voxelsize, G.T)[0])
    im768 = O.feature_mapping(voxels, [torch.from_numpy(m) for m in maps], torch.Tensor(list(synth.KITTI_IMSIZE_HW))).double()
    cnt, rows, row_v = CB.compact_rows(voxels, G.T)
    A1 = torch.cat([im768.reshape(-1, 768)[rows], torch.zeros(1, 768, dtype=torch.float64)], 0)
    vox7c = torch.cat([voxels[..., :7].double().reshape(-1, 7)[rows], torch.zeros(1, 7, dtype=torch.float64)], 0)
    out, st, aux = CB.compact_forward(A1, vox7c, sd, cnt, row_v, G.T)
    print('forward max diff', (out - vf).abs().max().item())
    grads = CB.compact_backward(st, aux, Gw, row_v)
    for k in sorted(gref):
        g = grads[k].reshape(gref[k].shape)
        print(f'{k:34s} rel err {((g - gref[k]).abs().max() / gref[k].abs().max()).item():.3e}')


if __name__ == '__main__':
    main()
