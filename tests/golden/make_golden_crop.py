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


def main():
    m = refshim.load()
    calib = synth.kitti_calib()
    imsize_wh = (synth.KITTI_IMSIZE_HW[1], synth.KITTI_IMSIZE_HW[0])
    out = {}
    for tag, seed, P in (('a', 21, 6000), ('b', 22, 20000)):
        rng = np.random.default_rng(seed)
        raw = np.concatenate([synth.make_raw_sweep(seed, P), boundary_points(rng)], axis=0)
        raw = np.ascontiguousarray(raw[rng.permutation(raw.shape[0])])
        # tag every point with its index in column 3 so the fixture can store indices instead of points
        tagged = raw.copy()
        tagged[:, 3] = np.arange(raw.shape[0], dtype=np.float32)
        c = m.pre.crop(tagged.copy(), m.cfg.velorange)                           # Preprocessing.py:12-17
        s = m.pre.cropToSight(tagged.copy(), calib, imsize_wh)                   # Preprocessing.py:26-55 (numpy branch)
        cs = m.pre.cropToSight(c.copy(), calib, imsize_wh)                       # Load.py:59,73 order
        out.update({f'raw_{tag}': raw, f'crop_{tag}': c[:, 3].astype(np.int32), f'sight_{tag}': s[:, 3].astype(np.int32),
                    f'both_{tag}': cs[:, 3].astype(np.int32)})
        print(tag, 'raw', raw.shape[0], 'crop', c.shape[0], 'sight', s.shape[0], 'both', cs.shape[0])
    np.savez_compressed(os.path.join(OUT, 'crop_a.npz'), **out)


if __name__ == '__main__':
    main()
