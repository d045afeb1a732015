"""GPU parity tests of the label-side kernels (csrc/anchors.cu) through the C ABI: bit-exact against the golden vectors of the
unmodified reference, against the C oracle on seeded inputs, and (where the prebuilt oracle/_ref module loads) against the
reference's own compiled `_classifyAnchors` at the full KITTI anchor grid."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KITTI_VELORANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
CARSIZE = [3.9, 1.6, 1.56]


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'anchors_a.npz'))


@pytest.fixture(scope='module')
def kitti_anchor_bevs():
    from oracle import iou_oracle as IO
    return IO.anchor_bevs(IO.create_anchors(176, 200, KITTI_VELORANGE, CARSIZE))


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float32).view(np.uint32), np.asarray(b, np.float32).view(np.uint32))


def assert_lists(res, pi, ni, gi, what):
    assert np.array_equal(np.stack(res[0]), np.stack(pi)), f'{what}: positive list differs'
    assert np.array_equal(np.stack(res[1]), np.stack(ni)), f'{what}: not-negative list differs'
    assert np.array_equal(res[2], gi), f'{what}: ground-truth index list differs'


def random_boxes(rng, G, vr, lw=((2.5, 6.0), (1.2, 3.0))):
    b = np.zeros((G, 7), np.float32)
    b[:, 0] = rng.uniform(vr[0] + 0.5, vr[3] - 0.5, G)
    b[:, 1] = rng.uniform(vr[1] + 0.5, vr[4] - 0.5, G)
    b[:, 3] = rng.uniform(*lw[0], G)
    b[:, 4] = rng.uniform(*lw[1], G)
    b[:, 6] = rng.uniform(-3.2, 3.2, G)
    return torch.from_numpy(b)


@pytest.mark.parametrize('tag', ['k1', 'k2'])
def test_classify_vs_reference_golden_kitti(gold, kitti_anchor_bevs, tag):
    from mvxnet_makise_b200.voxelize import cpp
    from mvxnet_makise_b200.anchors import classifyAnchors
    b3 = torch.from_numpy(gold[f'{tag}_boxes'])
    bev = torch.from_numpy(gold[f'{tag}_bev'])
    res = classifyAnchors(bev, b3[:, [0, 1]], kitti_anchor_bevs, KITTI_VELORANGE, 0.45, 0.6)   # Calc.py:88-96 surface
    assert_lists(res, gold[f'{tag}_pi'], gold[f'{tag}_ni'], gold[f'{tag}_gi'], tag)
    assert all(isinstance(a, np.ndarray) and a.dtype == np.int64 for a in (*res[0], *res[1], res[2]))   # pybind return types
    from oracle import iou_oracle as IO
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    res2 = cpp._classifyAnchors(bev.numpy(), kitti_anchor_bevs.numpy(), nls.numpy(), nws.numpy(), 0.45, 0.6)   # voxelutil surface
    assert_lists(res2, gold[f'{tag}_pi'], gold[f'{tag}_ni'], gold[f'{tag}_gi'], tag)


def test_classify_vs_reference_golden_small_dense(gold):
    from mvxnet_makise_b200.anchors import classifyAnchors
    b3 = torch.from_numpy(gold['s_boxes'])
    res = classifyAnchors(torch.from_numpy(gold['s_bev']), b3[:, [0, 1]], torch.from_numpy(gold['s_anchor_bev']),
                          list(gold['s_range']), 0.2, 0.35)
    assert_lists(res, gold['s_pi'], gold['s_ni'], gold['s_gi'], 'small')


def test_pairwise_vs_reference_golden(gold):
    from mvxnet_makise_b200.voxelize import cpp
    for q, b1, inter, iou in zip(gold['pw_q'], gold['pw_b1'], gold['pw_inter'], gold['pw_iou']):
        assert same_bits(cpp.bboxIntersection(b1, q[None])[:, 0], inter)
        assert same_bits(cpp.bboxOverlap(b1, q[None])[:, 0], iou)


def test_pairwise_vs_oracle_large():
    """400 x 600 rotated rectangles and general quads (some clockwise, some degenerate): bit-exact against the C oracle."""
    from mvxnet_makise_b200.voxelize import cpp
    from oracle import iou_oracle as IO
    rng = np.random.default_rng(31)
    b1 = IO.bbox3d2bev(random_boxes(rng, 400, [0, -10, 0, 20, 10, 0])).numpy()
    b2 = IO.bbox3d2bev(random_boxes(rng, 600, [0, -10, 0, 20, 10, 0])).numpy()
    b2[::7] = b2[::7, ::-1]                       # clockwise
    b2[5] = b2[5, 0]                              # a point
    b2[6, 2:] = b2[6, 1]                          # a segment
    b1[3] = b2[9]                                 # identical quads
    b1[4] = b2[10] + np.float32(1e-6)
    b2[11:40] += rng.uniform(-3, 3, (29, 4, 2)).astype(np.float32)   # general (possibly self-intersecting) quads
    for mode, fn in (('iou', cpp.bboxOverlap), ('inter', cpp.bboxIntersection)):
        ours, ref = fn(b1, b2), IO.pairwise(b1, b2, mode)
        assert ours.shape == (400, 600) and ours.dtype == np.float32
        assert np.array_equal(ours.view(np.uint32), ref.view(np.uint32)), f'{mode}: {np.sum(ours.view(np.uint32) != ref.view(np.uint32))} pairs differ'
    # CUDA tensors stay on the device; Augment.py:54's call shape (1, M)
    d = cpp.bboxOverlap(torch.from_numpy(b1[:1]).cuda(), torch.from_numpy(b2).cuda())
    assert d.is_cuda and d.shape == (1, 600) and same_bits(d.cpu().numpy(), IO.pairwise(b1[:1], b2, 'iou'))


def test_pairwise_empty_and_errors():
    from mvxnet_makise_b200.voxelize import cpp
    e = np.zeros((0, 4, 2), np.float32)
    sq = np.array([[[1, 1], [-1, 1], [-1, -1], [1, -1]]], np.float32)
    assert cpp.bboxOverlap(e, sq).shape == (0, 1) and cpp.bboxIntersection(sq, e).shape == (1, 0)
    assert cpp.bboxOverlap(sq, sq)[0, 0] == 1.0 and cpp.bboxIntersection(sq, sq)[0, 0] == 4.0
    with pytest.raises(ValueError):
        cpp.bboxOverlap(np.zeros((4, 2), np.float32), sq)   # pybind unchecked<3>() error type


@pytest.mark.parametrize('thr', [(0.45, 0.6), (0.12, 0.25), (0.6, 0.45)])
def test_classify_vs_oracle_many_ground_truths(kitti_anchor_bevs, thr):
    """300 ground truths (several frames' worth in one call) on the full 176x200x2 grid, wide boxes so that the walks are long
    and hit the grid border; thresholds incl. an inverted pair."""
    from mvxnet_makise_b200.anchors import AnchorClassifier
    from oracle import iou_oracle as IO
    rng = np.random.default_rng(41)
    b3 = random_boxes(rng, 300, KITTI_VELORANGE, lw=((2.0, 9.0), (1.0, 5.0)))
    b3[:8, 6] = torch.tensor([0, np.pi / 2, -np.pi / 2, np.pi, 1e-3, np.pi / 2 + 1e-3, 0.7853982, -0.7853982])
    b3[8, :2] = torch.tensor([0.1, -39.9])
    b3[9, :2] = torch.tensor([70.3, 39.9])
    bev = IO.bbox3d2bev(b3)
    bev[10] = bev[10].flip(0)        # clockwise ground truth: signed (negative) area enters the IoU like in the reference
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    clf = AnchorClassifier(kitti_anchor_bevs)
    ours = clf(bev, nls, nws, *thr)
    ref = IO.classify(bev, kitti_anchor_bevs, nls, nws, *thr)
    assert ref[3] == 0 and len(ref[1][0]) > 300
    assert_lists(ours, *ref[:3], f'thr={thr}')


def test_classify_edge_cases(kitti_anchor_bevs):
    from mvxnet_makise_b200 import _lib
    from mvxnet_makise_b200.anchors import AnchorClassifier
    from oracle import iou_oracle as IO
    clf = AnchorClassifier(kitti_anchor_bevs)
    # no ground truth at all
    res = clf(np.zeros((0, 4, 2), np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64), 0.45, 0.6)
    assert all(len(a) == 0 for a in (*res[0], *res[1], res[2]))
    # start cell outside the grid: counted, no entries (the reference reads out of bounds there)
    b3 = torch.tensor([[10.0, 0.0, -1, 3.9, 1.6, 1.5, 0.0], [30.0, 5.0, -1, 3.9, 1.6, 1.5, 0.3]])
    bev = IO.bbox3d2bev(b3)
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    nls[0] = 176
    pos, neg, gi, outside = clf.classify_device(bev, nls, nws, 0.45, 0.6)
    ref = IO.classify(bev, kitti_anchor_bevs, nls, nws, 0.45, 0.6)
    assert outside == 1 == ref[3] and np.array_equal(gi.cpu().numpy(), ref[2]) and torch.all(gi == 1)
    with pytest.raises(IndexError):
        clf(bev, nls, nws, 0.45, 0.6)
    # capacity smaller than the result: everything is counted, the stored prefix is correct, nothing is written past cap
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    ref = IO.classify(bev, kitti_anchor_bevs, nls, nws, 0.45, 0.6)
    g, nl, nw = bev.cuda().contiguous(), nls.cuda(), nws.cuda()
    n = ctypes.c_size_t()
    _lib.check(_lib.lib.mvx_classify_anchors_workspace_bytes(2, 176, 200, 2, ctypes.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device='cuda')
    cap = 3
    posb = torch.full((cap + 4, 3), -7, dtype=torch.int64, device='cuda')
    negb, gib = posb.clone(), torch.full((cap + 4,), -7, dtype=torch.int64, device='cuda')
    counts = torch.empty(4, dtype=torch.int64, device='cuda')
    p = _lib.ptr
    _lib.check(_lib.lib.mvx_classify_anchors(p(g), 2, p(clf.anchors), 176, 200, 2, p(nl), p(nw), 0.45, 0.6, p(posb), p(negb), p(gib), cap,
                                             p(counts), p(ws), n.value, _lib.stream_ptr()))
    c = counts.cpu().numpy()
    assert c[0] == len(ref[2]) > cap and c[1] == len(ref[1][0]) and c[2] == 0
    assert np.array_equal(posb[:cap].cpu().numpy(), np.stack(ref[0]).T[:cap]) and torch.all(posb[cap:] == -7)
    assert np.array_equal(negb[:cap].cpu().numpy(), np.stack(ref[1]).T[:cap]) and torch.all(negb[cap:] == -7)
    assert _lib.lib.mvx_classify_anchors(p(g), 2, p(clf.anchors), 176, 200, 2, p(nl), p(nw), 0.45, 0.6, p(posb), p(negb), p(gib), cap,
                                         p(counts), p(ws), 8, _lib.stream_ptr()) == -3   # MVX_ESPACE


def test_classify_vs_live_reference_module(kitti_anchor_bevs):
    """The reference's own compiled extension (oracle/_ref, prebuilt in the build container and shipped with the snapshot)."""
    import glob
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not glob.glob(os.path.join(here, 'oracle', '_ref', 'voxelutil*.so')):
        pytest.skip('oracle/_ref not built')
    from oracle import refshim, iou_oracle as IO
    from mvxnet_makise_b200.voxelize import cpp
    vu = refshim.load_voxelutil()
    rng = np.random.default_rng(51)
    b3 = random_boxes(rng, 60, KITTI_VELORANGE, lw=((3.0, 5.0), (1.4, 2.2)))
    bev = IO.bbox3d2bev(b3).numpy()
    nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, KITTI_VELORANGE)
    ab = kitti_anchor_bevs.numpy()
    ref = vu._classifyAnchors(bev, ab, nls.numpy(), nws.numpy(), 0.45, 0.6)
    ours = cpp._classifyAnchors(bev, ab, nls.numpy(), nws.numpy(), 0.45, 0.6)
    assert len(ref[2]) > 50
    assert_lists(ours, *ref, 'live reference')
