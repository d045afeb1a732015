"""Generates tests/golden/anchors_a.npz by running the UNMODIFIED reference: `createAnchors` (Preprocessing.py:118-142),
`bbox3d2bev` + `classifyAnchors` (modules/Calc.py:15-36, 88-96) and the compiled cpp/voxelutil.cpp (`_classifyAnchors`,
`bboxOverlap`, `bboxIntersection`). Build-container only (needs /root/reference).
Run from the repo root:   python tests/golden/make_golden_anchors.py

The pairwise functions of the reference fill their second quad through the BOX index (voxelutil.cpp:108,129); they are
driven here in the one configuration where that is harmless: a single box in bboxes2 whose last corner equals the first
corner of the quad Q that a preceding `_classifyAnchors` call left in the global `r2` — the call then evaluates every quad
of bboxes1 against Q with the reference's own clipper (Q counter-clockwise, so `r2` is not re-oriented in place)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
KITTI_VELORANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
CARSIZE = [3.9, 1.6, 1.56]


def random_boxes(rng, G, velorange, axis_aligned=4):
    b = np.zeros((G, 7), np.float32)
    b[:, 0] = rng.uniform(velorange[0] + 1, velorange[3] - 1, G)
    b[:, 1] = rng.uniform(velorange[1] + 1, velorange[4] - 1, G)
    b[:, 2] = -1
    b[:, 3] = rng.uniform(3.2, 4.6, G)
    b[:, 4] = rng.uniform(1.4, 1.9, G)
    b[:, 5] = 1.5
    b[:, 6] = rng.uniform(-np.pi, np.pi, G)
    b[:axis_aligned, 6] = np.array([0, np.pi / 2, 0.02, np.pi / 2 - 0.03, -np.pi / 2, np.pi], np.float32)[:axis_aligned]
    return torch.from_numpy(b)


def star_quad(rng, centre, flip=False):
    while True:
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        rad = rng.uniform(0.5, 3, 4)
        q = np.stack([centre[0] + rad * np.cos(ang), centre[1] + rad * np.sin(ang)], 1).astype(np.float32)
        x, y = q[:, 0].astype(np.float64), q[:, 1].astype(np.float64)
        if 0.5 * np.sum(x * np.roll(y, -1) - y * np.roll(x, -1)) > 0.05:
            return q[::-1].copy() if flip else q


def ref_vs_quad(vu, b1, q, fn):
    q = np.ascontiguousarray(q, np.float32)
    z = np.zeros(1, np.int64)
    vu._classifyAnchors(q[None], q[None, None, None], z, z, 0.45, 0.6)   # leaves q in the global r2
    b2 = np.zeros((1, 4, 2), np.float32)
    b2[0, 3] = q[0]
    return fn(np.ascontiguousarray(b1, np.float32), b2)[:, 0].copy()


def main():
    m = refshim.load()
    calc = refshim.load_calc()
    vu = m.cpp
    out = {}
    # ---- KITTI anchor grid (train.py:59-61) with car-sized ground truths, thresholds of train.py:46 ------------------
    anchors = m.pre.createAnchors(176, 200, KITTI_VELORANGE, CARSIZE)
    abev = calc.bbox3d2bev(anchors.reshape(anchors.shape[:2] + (-1, 7)))
    out['kitti_anchor_probe'] = abev[::25, ::25].numpy().copy()   # the test rebuilds the anchors and checks these
    for tag, seed, G in (('k1', 11, 14), ('k2', 12, 40)):
        rng = np.random.default_rng(seed)
        b3 = random_boxes(rng, G, KITTI_VELORANGE)
        bev = calc.bbox3d2bev(b3)
        pi, ni, gi = calc.classifyAnchors(bev, b3[:, [0, 1]], abev, KITTI_VELORANGE, 0.45, 0.6)
        out.update({f'{tag}_boxes': b3.numpy(), f'{tag}_bev': bev.numpy(), f'{tag}_pi': np.stack(pi), f'{tag}_ni': np.stack(ni),
                    f'{tag}_gi': np.asarray(gi)})
        print(tag, 'G', G, 'pos', len(gi), 'neg', len(ni[0]))
    # ---- small dense grid, wide boxes, low thresholds: long walks in every direction, border hits ----------------------
    rng = np.random.default_rng(13)
    vr = [0.0, -8.0, -3.0, 16.0, 8.0, 1.0]
    anchors = m.pre.createAnchors(40, 50, vr, [2.5, 1.2, 1.5])
    abev = calc.bbox3d2bev(anchors.reshape(anchors.shape[:2] + (-1, 7)))
    b3 = random_boxes(rng, 30, vr, axis_aligned=6)
    b3[:, 3] = torch.from_numpy(rng.uniform(1.5, 6.0, 30).astype(np.float32))
    b3[:, 4] = torch.from_numpy(rng.uniform(0.8, 3.0, 30).astype(np.float32))
    b3[6, :2] = torch.tensor([0.3, -7.8])     # corner of the grid
    b3[7, :2] = torch.tensor([15.8, 7.7])
    bev = calc.bbox3d2bev(b3)
    pi, ni, gi = calc.classifyAnchors(bev, b3[:, [0, 1]], abev, vr, 0.2, 0.35)
    out.update({'s_range': np.array(vr), 's_size': np.array([2.5, 1.2, 1.5]), 's_anchor_bev': abev.numpy(), 's_boxes': b3.numpy(),
                's_bev': bev.numpy(), 's_pi': np.stack(pi), 's_ni': np.stack(ni), 's_gi': np.asarray(gi)})
    print('small', 'pos', len(gi), 'neg', len(ni[0]))
    # ---- pairwise clipper pin: general quads (clockwise ones too) against counter-clockwise quads Q ------------------
    rng = np.random.default_rng(14)
    qs, b1s, inters, ious = [], [], [], []
    for t in range(24):
        c = rng.uniform(-5, 5, 2)
        q = star_quad(rng, c)
        b1 = np.stack([star_quad(rng, c + rng.uniform(-3, 3, 2), flip=(i % 3 == 0)) for i in range(40)])
        if t < 4:   # rectangles sharing edges / corners with Q's bounding rectangle: degenerate contacts
            q = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], np.float32) * np.float32(1 + t) + c.astype(np.float32)
            b1[:8] = q[None] + np.array([[2 + 2 * t, 0], [0, 2 + 2 * t], [1, 1], [0, 0], [0.5, 0], [-2 - 2 * t, -2 - 2 * t], [1e-7, 0],
                                          [0, 3 + 2 * t]], np.float32)[:, None, :]
        qs.append(q)
        b1s.append(b1)
        inters.append(ref_vs_quad(vu, b1, q, vu.bboxIntersection))
        ious.append(ref_vs_quad(vu, b1, q, vu.bboxOverlap))
    out.update({'pw_q': np.stack(qs), 'pw_b1': np.stack(b1s), 'pw_inter': np.stack(inters), 'pw_iou': np.stack(ious)})
    np.savez_compressed(os.path.join(OUT, 'anchors_a.npz'), **out)
    print('wrote anchors_a.npz', os.path.getsize(os.path.join(OUT, 'anchors_a.npz')), 'bytes')


if __name__ == '__main__':
    main()
