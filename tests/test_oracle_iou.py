"""CPU tests: the rotated-IoU / anchor-classification oracle (oracle/iou_oracle.c, .py) against the golden vectors generated from
the UNMODIFIED reference (tests/golden/make_golden_anchors.py) and, where /root/reference exists, against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import iou_oracle as IO
from oracle import refshim

KITTI_VELORANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
CARSIZE = [3.9, 1.6, 1.56]


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'anchors_a.npz'))


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def check_lists(res, g, tag):
    pi, ni, gi = res[0], res[1], res[2]
    assert np.array_equal(np.stack(pi), g[f'{tag}_pi']), f'{tag}: positive index list differs'
    assert np.array_equal(np.stack(ni), g[f'{tag}_ni']), f'{tag}: not-negative index list differs'
    assert np.array_equal(gi, g[f'{tag}_gi']), f'{tag}: ground-truth index list differs'


def test_python_restatements_match_reference(gold):
    abev = IO.anchor_bevs(IO.create_anchors(176, 200, KITTI_VELORANGE, CARSIZE))
    assert abev.shape == (176, 200, 2, 4, 2)
    assert same_bits(abev[::25, ::25].numpy(), gold['kitti_anchor_probe'])
    for tag in ('k1', 'k2', 's'):
        assert same_bits(IO.bbox3d2bev(torch.from_numpy(gold[f'{tag}_boxes'])).numpy(), gold[f'{tag}_bev'])
    sb = IO.anchor_bevs(IO.create_anchors(40, 50, list(gold['s_range']), list(gold['s_size'])))
    assert same_bits(sb.numpy(), gold['s_anchor_bev'])


@pytest.mark.parametrize('tag', ['k1', 'k2'])
def test_classify_oracle_vs_golden_kitti(gold, tag):
    abev = IO.anchor_bevs(IO.create_anchors(176, 200, KITTI_VELORANGE, CARSIZE))
    b3 = torch.from_numpy(gold[f'{tag}_boxes'])
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    res = IO.classify(gold[f'{tag}_bev'], abev, nls, nws, 0.45, 0.6)
    assert res[3] == 0
    check_lists(res, gold, tag)


def test_classify_oracle_vs_golden_small(gold):
    b3 = torch.from_numpy(gold['s_boxes'])
    nls, nws = IO.start_cells(b3[:, [0, 1]], 40, 50, list(gold['s_range']))
    res = IO.classify(gold['s_bev'], gold['s_anchor_bev'], nls, nws, 0.2, 0.35)
    check_lists(res, gold, 's')
    assert len(res[2]) > 200   # the case is a dense one


def test_pairwise_oracle_vs_golden(gold):
    for q, b1, inter, iou in zip(gold['pw_q'], gold['pw_b1'], gold['pw_inter'], gold['pw_iou']):
        assert same_bits(IO.pairwise(b1, q[None], 'inter')[:, 0], inter)
        assert same_bits(IO.pairwise(b1, q[None], 'iou')[:, 0], iou)


def test_pairwise_properties():
    rng = np.random.default_rng(5)
    sq = np.array([[1, 1], [-1, 1], [-1, -1], [1, -1]], np.float32)
    assert IO.pairwise(sq[None], sq[None], 'inter')[0, 0] == 4.0 and IO.pairwise(sq[None], sq[None], 'iou')[0, 0] == 1.0
    assert IO.pairwise(sq[None], (sq + np.float32(5))[None], 'inter')[0, 0] == 0.0
    half = sq + np.array([1, 0], np.float32)
    assert abs(IO.pairwise(sq[None], half[None], 'inter')[0, 0] - 2.0) < 1e-6
    assert abs(IO.pairwise(sq[None], half[None], 'iou')[0, 0] - 1 / 3) < 1e-6
    # clockwise input is re-oriented: same intersection
    assert IO.pairwise(sq[::-1].copy()[None], half[None], 'inter')[0, 0] == IO.pairwise(sq[None], half[None], 'inter')[0, 0]
    # rotated rectangles: intersection is symmetric up to rounding and bounded by both areas
    b = np.zeros((64, 7), np.float32)
    b[:, :2] = rng.uniform(-2, 2, (64, 2)); b[:, 3] = rng.uniform(1, 4, 64); b[:, 4] = rng.uniform(1, 3, 64); b[:, 6] = rng.uniform(-3, 3, 64)
    bev = IO.bbox3d2bev(torch.from_numpy(b)).numpy()
    m = IO.pairwise(bev, bev, 'inter')
    area = b[:, 3] * b[:, 4]
    assert np.allclose(m, m.T, atol=1e-4) and np.allclose(np.diag(m), area, rtol=1e-5)
    assert np.all(m <= np.minimum(area[:, None], area[None, :]) + 1e-4) and np.all(m >= -1e-4)


def test_classify_outside_start_cell_is_counted():
    abev = IO.anchor_bevs(IO.create_anchors(8, 8, [0, -4, -3, 8, 4, 1], [2.0, 1.0, 1.5]))
    sq = np.array([[[1, 1], [-1, 1], [-1, -1], [1, -1]]], np.float32)
    res = IO.classify(np.concatenate([sq, sq + np.float32(3)]), abev, np.array([-1, 3]), np.array([2, 3]), 0.1, 0.3)
    assert res[3] == 1 and np.all(res[2] == 1)


@pytest.mark.skipif(not refshim.available(), reason='needs /root/reference (build container)')
def test_oracle_vs_live_reference():
    vu = refshim.load_voxelutil()
    rng = np.random.default_rng(77)
    abev = IO.anchor_bevs(IO.create_anchors(88, 100, KITTI_VELORANGE, CARSIZE)).numpy()
    for trial in range(6):
        G = 25
        b = np.zeros((G, 7), np.float32)
        b[:, 0] = rng.uniform(1, 69, G); b[:, 1] = rng.uniform(-39, 39, G)
        b[:, 3] = rng.uniform(2.5, 6, G); b[:, 4] = rng.uniform(1.2, 3, G); b[:, 6] = rng.uniform(-3.2, 3.2, G)
        b3 = torch.from_numpy(b)
        bev = IO.bbox3d2bev(b3).numpy()
        nls, nws = IO.start_cells(b3[:, [0, 1]], 88, 100, KITTI_VELORANGE)
        neg_thr, pos_thr = (0.45, 0.6) if trial % 2 else (0.15, 0.3)
        ref = vu._classifyAnchors(bev, abev, nls.numpy(), nws.numpy(), neg_thr, pos_thr)
        ours = IO.classify(bev, abev, nls, nws, neg_thr, pos_thr)
        for k in range(3):
            assert np.array_equal(ref[0][k], ours[0][k]) and np.array_equal(ref[1][k], ours[1][k])
        assert np.array_equal(ref[2], ours[2])
