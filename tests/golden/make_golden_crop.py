"""Generates tests/golden/crop_a.npz by running the UNMODIFIED reference's `crop` / `cropToSight`
(modules/data/Preprocessing.py:12-55) on raw synthetic sweeps. Build-container only (needs /root/reference).
Run from the repo root:   python tests/golden/make_golden_crop.py"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def boundary_points(rng, n=300):
    """points sitting exactly on the range bounds and their fp32 neighbours"""
    r = synth.KITTI_VELORANGE
    pts = []
    for d in range(3):
        for b in (r[d], r[3 + d]):
            v = np.float32(b)
            for w in (np.nextafter(v, np.float32(-1e9)), v, np.nextafter(v, np.float32(1e9))):
                for _ in range(n // 18):
                    p = np.array([rng.uniform(1, 60), rng.uniform(-30, 30), rng.uniform(-2.5, 0.5), rng.uniform()], dtype=np.float32)
                    p[d] = w
                    pts.append(p)
    return np.array(pts, dtype=np.float32)


def border_flip_points(calib32, calib64, imsize_wh, n_want=40, seed=3):
    """points whose fp32 and fp64 sight decisions DIFFER (projection within an fp32 rounding error of the right image border):
    without them the float64-calibration goldens could not tell the two arithmetic paths apart. Found by bisection on the
    fp32 y coordinate (which moves u) down to two ADJACENT floats that straddle the border in fp64; kept where the fp32
    evaluation decides differently for one of them."""
    from oracle import pointpath_oracle as O
    rng = np.random.default_rng(seed)
    n = 4000
    lim_w = imsize_wh[0] - 1e-3

    def u64(pts):
        return O.lidar2img_numpy(pts, calib64)[:, 0]

    x = rng.uniform(6, 60, n).astype(np.float32)
    z = rng.uniform(-1.5, 0.2, n).astype(np.float32)
    lo = np.zeros(n, np.float32)                  # y = 0: near the image centre (inside)
    hi = (-x * np.float32(1.2)).astype(np.float32)  # far to the right: u > w (outside)
    mk = lambda y: np.stack([x, y, z, np.zeros(n, np.float32)], axis=1).astype(np.float32)
    ok = (u64(mk(lo)) < lim_w) & (u64(mk(hi)) >= lim_w)
    for _ in range(40):
        mid = ((lo.astype(np.float64) + hi.astype(np.float64)) / 2).astype(np.float32)
        inside = u64(mk(mid)) < lim_w
        lo = np.where(inside, mid, lo)
        hi = np.where(inside, hi, mid)
    cand = np.concatenate([mk(lo)[ok], mk(hi)[ok], mk(np.nextafter(lo, np.float32(10)))[ok], mk(np.nextafter(hi, np.float32(-1e9)))[ok]])
    cand[:, 3] = np.arange(cand.shape[0])
    k32 = set(O.crop_to_sight(cand, calib32, imsize_wh)[:, 3].astype(int))
    k64 = set(O.crop_to_sight(cand, calib64, imsize_wh)[:, 3].astype(int))
    flips = cand[sorted(k32 ^ k64)][:n_want]
    near = cand[rng.permutation(cand.shape[0])[:n_want]]      # and border points on which both agree
    return np.concatenate([flips, near], axis=0).astype(np.float32), flips.shape[0]


def main():
    m = refshim.load()
    calib = synth.kitti_calib()
    calib64 = synth.kitti_calib_f64()      # the dict exactly as readCalib builds it (float64 matrices): Load.py:73's case
    imsize_wh = (synth.KITTI_IMSIZE_HW[1], synth.KITTI_IMSIZE_HW[0])
    out = {}
    for tag, seed, P in (('a', 21, 6000), ('b', 22, 20000)):
        rng = np.random.default_rng(seed)
        flips, nflip = border_flip_points(calib, calib64, imsize_wh, seed=seed)
        print(tag, 'border points whose fp32 / fp64 decisions differ:', nflip, 'of', flips.shape[0])
        raw = np.concatenate([synth.make_raw_sweep(seed, P), boundary_points(rng), flips], axis=0)
        raw = np.ascontiguousarray(raw[rng.permutation(raw.shape[0])])
        # tag every point with its index in column 3 so the fixture can store indices instead of points
        tagged = raw.copy()
        tagged[:, 3] = np.arange(raw.shape[0], dtype=np.float32)
        c = m.pre.crop(tagged.copy(), m.cfg.velorange)                           # Preprocessing.py:12-17
        s = m.pre.cropToSight(tagged.copy(), calib, imsize_wh)                   # Preprocessing.py:26-55 (numpy branch)
        cs = m.pre.cropToSight(c.copy(), calib, imsize_wh)                       # Load.py:59,73 order
        s64 = m.pre.cropToSight(tagged.copy(), calib64, imsize_wh)               # numpy branch in fp64
        cs64 = m.pre.cropToSight(c.copy(), calib64, imsize_wh)
        out.update({f'raw_{tag}': raw, f'crop_{tag}': c[:, 3].astype(np.int32), f'sight_{tag}': s[:, 3].astype(np.int32),
                    f'both_{tag}': cs[:, 3].astype(np.int32), f'sight64_{tag}': s64[:, 3].astype(np.int32),
                    f'both64_{tag}': cs64[:, 3].astype(np.int32)})
        print(tag, 'fp64 calib: sight', s64.shape[0], 'both', cs64.shape[0], 'differs from fp32 in',
              len(set(s[:, 3].astype(int)) ^ set(s64[:, 3].astype(int))), 'points')
        print(tag, 'raw', raw.shape[0], 'crop', c.shape[0], 'sight', s.shape[0], 'both', cs.shape[0])
    # GT-paste data format with readCalib-style float64 dicts (train.py:29-42): scene through the torch branch, pasted sets
    # through the numpy branch of the UNMODIFIED lidar2Img, merged like train.py:42; stored as the fp32 the model sees
    rng = np.random.default_rng(5)
    scene, pasted1, pasted2 = synth.make_points(33, 3000), synth.make_points(34, 500), synth.make_points(35, 300)
    other = {k: v.copy() for k, v in calib64.items()}
    other['P2'][0, 0] += np.float32(9.25); other['P2'][1, 2] -= np.float32(1.5); other['Tr_velo_to_cam'][1, 3] += np.float32(0.0127)
    import torch
    pcd = torch.Tensor(scene)
    ct = {k: torch.Tensor(v) for k, v in calib64.items()}                        # Load.py:75-76
    merged = [torch.concat([pcd, m.calib.lidar2Img(pcd, ct, True)[:, [1, 0]]], dim=1).numpy()]     # train.py:31-35
    for ap, ac in ((pasted1, calib64), (pasted2, other)):
        proj = m.calib.lidar2Img(ap, ac, True)[:, ::-1]                          # train.py:37-38 (numpy branch, float64)
        assert proj.dtype == np.float64
        merged.append(np.concatenate([ap, proj], axis=1))
    merged = np.concatenate(merged, axis=0)
    assert merged.dtype == np.float64
    uv64 = m.calib.lidar2Img(pasted2, other, True)
    out.update(dict(m_scene=scene, m_p1=pasted1, m_p2=pasted2, m_other_P2=other['P2'], m_other_Tr=other['Tr_velo_to_cam'],
                    m_other_R0=other['R0_rect'], m_merged32=merged.astype(np.float32), m_uv64_p2=uv64))
    np.savez_compressed(os.path.join(OUT, 'crop_a.npz'), **out)


if __name__ == '__main__':
    main()
