"""Dev tool (GPU): cost of the label-side kernels next to the reference's compiled extension on the host.
Workload: the ground truths of one training batch per GPU (BASELINE configs[3]: 16 frames x 12 cars, train.py:28 pastes up to 12)
on the 176x200x2 KITTI anchor grid, and rotated-IoU matrices."""
import glob, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200.anchors import AnchorClassifier, bboxOverlap
from oracle import iou_oracle as IO, refshim

VR = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
abev = IO.anchor_bevs(IO.create_anchors(176, 200, VR, [3.9, 1.6, 1.56]))
rng = np.random.default_rng(0)
G = 16 * 12
b = np.zeros((G, 7), np.float32)
b[:, 0] = rng.uniform(2, 68, G); b[:, 1] = rng.uniform(-38, 38, G); b[:, 3] = rng.uniform(3.2, 4.6, G); b[:, 4] = rng.uniform(1.4, 1.9, G)
b[:, 6] = np.where(rng.uniform(size=G) < 0.7, rng.normal(0, 0.1, G) + rng.integers(0, 2, G) * np.pi / 2, rng.uniform(-3.1, 3.1, G))
b3 = torch.from_numpy(b)
bev = IO.bbox3d2bev(b3)
nls, nws = IO.start_cells(b3[:, [0, 1]], 176, 200, VR)
clf = AnchorClassifier(abev)
d_bev, d_nl, d_nw = bev.cuda(), nls.cuda(), nws.cuda()


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


pos, neg, gi, _ = clf.classify_device(d_bev, d_nl, d_nw, 0.45, 0.6)
t_gpu = timed(lambda: clf.classify_device(d_bev, d_nl, d_nw, 0.45, 0.6))
out = {'classify': {'ground_truths': G, 'pos': int(gi.numel()), 'not_neg': int(neg.shape[0]), 'gpu_ms_incl_count_readback': round(t_gpu, 4)}}
have_ref = bool(glob.glob(os.path.join(ROOT, 'oracle', '_ref', 'voxelutil*.so')))
if have_ref:
    vu = refshim.load_voxelutil()
    a_np, g_np, nl_np, nw_np = abev.numpy(), bev.numpy(), nls.numpy(), nws.numpy()
    vu._classifyAnchors(g_np, a_np, nl_np, nw_np, 0.45, 0.6)
    t = time.perf_counter()
    for _ in range(5):
        r = vu._classifyAnchors(g_np, a_np, nl_np, nw_np, 0.45, 0.6)
    out['classify']['reference_cpu_ms'] = round((time.perf_counter() - t) / 5 * 1e3, 3)
    assert np.array_equal(r[2], gi.cpu().numpy()) and np.array_equal(np.stack(r[1]).T, neg.cpu().numpy())
t = time.perf_counter()
for _ in range(5):
    IO.classify(bev, abev, nls, nws, 0.45, 0.6)
out['classify']['oracle_c_cpu_ms'] = round((time.perf_counter() - t) / 5 * 1e3, 3)

for n, m in ((1, 12), (2048, 2048)):
    q = np.zeros((max(n, m), 7), np.float32)
    q[:, :2] = rng.uniform(-20, 20, (max(n, m), 2)); q[:, 3] = rng.uniform(2, 6, max(n, m)); q[:, 4] = rng.uniform(1, 3, max(n, m)); q[:, 6] = rng.uniform(-3, 3, max(n, m))
    qb = IO.bbox3d2bev(torch.from_numpy(q))
    d1, d2 = qb[:n].cuda().contiguous(), qb[:m].cuda().contiguous()
    t_g = timed(lambda: bboxOverlap(d1, d2))
    t = time.perf_counter()
    ref = IO.pairwise(qb[:n], qb[:m], 'iou')
    t_c = (time.perf_counter() - t) * 1e3
    assert np.array_equal(ref.view(np.uint32), bboxOverlap(d1, d2).cpu().numpy().view(np.uint32))
    out[f'bboxOverlap_{n}x{m}'] = {'gpu_ms': round(t_g, 4), 'oracle_c_cpu_ms': round(t_c, 3), 'pairs_per_s_gpu': round(n * m / t_g * 1e3)}
print(json.dumps(out))
