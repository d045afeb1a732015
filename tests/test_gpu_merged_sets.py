"""GPU parity test of merged point sets (the data format GT-paste hands to the path, train.py:29-42): the
scene and every pasted ground-truth object carry their own calibration; each set is projected through its own matrices, then all
sets are concatenated in order and voxelized together. Checker: the oracle port fed the same list of (point set, calibration)."""
import numpy as np
import pytest
import torch

from _util import same_occupancy

from mvxnet_makise_b200 import synth
from oracle import pointpath_oracle as O

pytestmark = pytest.mark.gpu
G = synth.KITTI_GRID


def other_calib(seed):
    """a different camera: focal length, principal point and the lidar->camera translation moved"""
    rng = np.random.default_rng(seed)
    c = {k: np.array(v, dtype=np.float32, copy=True) for k, v in synth.kitti_calib().items()}
    c['P2'][0, 0] += np.float32(rng.uniform(-15, 15)); c['P2'][1, 1] = c['P2'][0, 0]
    c['P2'][0, 2] += np.float32(rng.uniform(-8, 8)); c['P2'][1, 2] += np.float32(rng.uniform(-4, 4))
    c['Tr_velo_to_cam'][:3, 3] += rng.uniform(-0.05, 0.05, 3).astype(np.float32)
    return c


def pasted_object(seed, n, calib):
    """a car-sized cluster of returns, inside the field of view of ITS OWN camera (the ground-truth database is cut with
    cropToSight against each object's calibration, create_gtdatabase.py; featureMaping asserts in-image projections)"""
    rng = np.random.default_rng(seed)
    cloud = synth.make_points(500 + seed, 40_000)
    centre = cloud[rng.integers(0, cloud.shape[0]), :3]
    near = cloud[(np.abs(cloud[:, :3] - centre) < np.array([2.0, 0.9, 0.75], np.float32)).all(1)]
    near = O.crop_to_sight(near, calib, (synth.KITTI_IMSIZE_HW[1], synth.KITTI_IMSIZE_HW[0]))
    # densify: jittered copies keep the cluster car-sized but give it n distinct points
    reps = int(np.ceil(n / max(near.shape[0], 1)))
    p = np.tile(near, (reps, 1))[:n].copy()
    p[:, :3] += rng.normal(0, 0.02, (p.shape[0], 3)).astype(np.float32)
    return O.crop_to_sight(p, calib, (synth.KITTI_IMSIZE_HW[1], synth.KITTI_IMSIZE_HW[0]))


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def test_merged_point_sets_with_their_own_calibrations():
    from mvxnet_makise_b200.pipeline import PointPath
    shapes = [(13, 42), (7, 21), (4, 11)]
    rng = np.random.default_rng(17)
    maps = [rng.standard_normal((3, 256, h, w), dtype=np.float32) for h, w in shapes]
    sd = synth.make_weights(4)
    base = synth.kitti_calib()
    calibs = [base, [base, other_calib(1), other_calib(2), other_calib(3)], [base, other_calib(4)]]
    frames = [
        synth.make_points(60, 2500),                                                                   # plain frame, one calibration
        [synth.make_points(61, 2000)] + [pasted_object(k, n, calibs[1][k]) for k, n in ((1, 400), (2, 250), (3, 600))],
        [synth.make_points(62, 1500), pasted_object(4, 300, calibs[2][1])],
    ]
    assert all(q.shape[0] > 50 for fr in frames[1:] for q in fr[1:])
    path = PointPath(sd, G)
    grids, counts = path(frames, calibs, [torch.from_numpy(m) for m in maps])
    torch.cuda.synchronize()
    counts = counts.cpu().numpy()
    assert np.all(counts[:, 2] == 0)
    cap = path.cap
    for f in range(3):
        sets, cals = (frames[f], calibs[f]) if isinstance(calibs[f], list) else ([frames[f]], [calibs[f]])
        fm = [m[f:f + 1] for m in maps]
        with torch.no_grad():
            ref = O.forward_frame(sets, cals, fm, sd, G, synth.KITTI_IMSIZE_HW)
            ref64 = O.forward_frame(sets, cals, fm, sd, G, synth.KITTI_IMSIZE_HW, dtype=torch.float64)
        N = ref['idx'].shape[0]
        K = int((ref['voxels9'][..., :3] != 0).any(-1).sum())
        assert counts[f, 0] == N and counts[f, 1] == K
        vfeat, idx = path.voxel_features(f)
        assert np.array_equal(idx.cpu().numpy()[:, 1:], ref['idx'].numpy()[:, 1:]), 'voxel coordinates differ'
        # per-row projection (row, col): bit-exact against the reference's per-set lidar2Img
        cnt = (ref['voxels9'][..., :3] != 0).any(-1).sum(1).numpy()
        dense_rows = np.concatenate([v * G.T + np.arange(c) for v, c in enumerate(cnt)])
        proj = path.region('proj', torch.float32, (3, cap + 128, 2))[f, :K].cpu()
        assert torch.equal(proj, ref['voxels9'].reshape(-1, 9)[dense_rows][:, 7:9]), f'frame {f}: projections differ'
        assert rel(vfeat.cpu(), ref64['vfeat']) < 1e-4
        assert rel(vfeat.cpu(), ref['vfeat']) <= rel(ref['vfeat'], ref64['vfeat']) + 1e-4
        g = grids[f].cpu()
        assert same_occupancy(g, ref64['grid'][0]) and rel(g, ref64['grid'][0]) < 1e-4
    # the calibration matters: the same merged frame pushed through ONE calibration gives different projections
    path2 = PointPath(sd, G)
    merged = np.concatenate(frames[1], axis=0)
    path2([merged], [base], [torch.from_numpy(m[1:2]) for m in maps])
    p_one = path2.region('proj', torch.float32, (1, path2.cap + 128, 2))[0, :counts[1, 1]].cpu()
    p_own = path.region('proj', torch.float32, (3, cap + 128, 2))[1, :counts[1, 1]].cpu()
    assert not torch.equal(p_one, p_own)


def test_merged_sets_with_float64_calibrations(golden_dir):
    """GT-paste with the calibration dicts as `readCalib` leaves them (float64 matrices, Load.py:24-41; LoadGT.py:31): the
    reference projects the scene with torch in fp32 (train.py:31-33) and every pasted set with numpy in fp64 (train.py:36-39),
    rounds once to fp32 (train.py:125). Golden: the UNMODIFIED lidar2Img (tests/golden/make_golden_crop.py)."""
    import os
    from mvxnet_makise_b200.pipeline import PointPath
    from mvxnet_makise_b200 import modules as M
    g = np.load(os.path.join(golden_dir, 'crop_a.npz'))
    c64 = synth.kitti_calib_f64()
    other = {'P2': g['m_other_P2'], 'R0_rect': g['m_other_R0'], 'Tr_velo_to_cam': g['m_other_Tr']}
    sets, cals = [g['m_scene'], g['m_p1'], g['m_p2']], [c64, c64, other]
    # numpy lidar2Img with a float64 dict: float64 result, equal to the reference's
    uv = M.lidar2Img(g['m_p2'], other, True)
    assert uv.dtype == np.float64 and np.array_equal(uv, g['m_uv64_p2'])
    rng = np.random.default_rng(3)
    maps = [rng.standard_normal((1, 256, h, w), dtype=np.float32) for h, w in [(13, 42), (7, 21), (4, 11)]]
    sd = synth.make_weights(6)
    path = PointPath(sd, G)
    _, counts = path([sets], [cals], [torch.from_numpy(m) for m in maps], want_grid=False)
    torch.cuda.synchronize()
    K = int(counts[0, 1])
    merged = g['m_merged32']                                               # (P,6) fp32 [x y z r row col] as the model sees it
    rp = path.region('row_point', torch.int32, (1, path.cap))[0, :K].cpu().numpy()
    proj = path.region('proj', torch.float32, (1, path.cap + 128, 2))[0, :K].cpu().numpy()
    assert np.array_equal(proj.view(np.uint32), np.ascontiguousarray(merged[rp][:, 4:6]).view(np.uint32)), 'projections differ'
    with torch.no_grad():
        ref64 = O.forward_frame(sets, cals, maps, sd, G, synth.KITTI_IMSIZE_HW, dtype=torch.float64, want_grid=False)
    vfeat, idx = path.voxel_features(0)
    assert np.array_equal(idx.cpu().numpy()[:, 1:], ref64['idx'].numpy()[:, 1:])
    assert rel(vfeat.cpu(), ref64['vfeat']) < 1e-4
