"""Value parity AT THE BENCHMARKED CONFIGURATION (BASELINE.json configs[1] and configs[4]).

The other GPU parity tests check values on golden frames of 1.5-4 k points and 13x42 maps; what changes with size - the
Morton-sorted combine over 100 k rows, the fp64-atomic BatchNorm statistics of 470 CTAs, the 256-row tensor-core tiles, the
per-voxel atomicMax - is only exercised here. Inputs are EXACTLY bench.py's: frame id g -> synth.make_points(g, 120 000),
synth.make_fpn_maps(g) (real FPN shapes 104x336 / 52x168 / 26x84), synth.make_weights(0), batch 8, rank 0 (frame ids 0..7).

Checker: oracle.forward_frame. Bars: voxel coordinates and counts bit-exact; voxel features and grid within 1e-4
(max|a-ref| / max|ref|) of the fp64 evaluation of the reference algorithm; vs the fp32 evaluation (the reference's own
arithmetic) within that reference's own distance to fp64 + 1e-4. The fp64 value of one frame costs ~15 s on the host
cores, so frame 0 is evaluated on the CPU in fp32 and fp64 and the other frames' fp64 values come from the same oracle
expressions evaluated with torch's CUDA fp64 kernels, pinned on frame 0 against the CPU evaluation (<= 1e-9)."""
import numpy as np
import pytest
import torch

from mvxnet_makise_b200 import synth
from oracle import pointpath_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
B = 8
P = 120_000


def rel_err(a, ref):
    a, ref = torch.as_tensor(a).double(), torch.as_tensor(ref).double()
    return float((a - ref.to(a.device)).abs().max() / ref.abs().max().clamp_min(1e-30))


def _oracle(pts, maps, sd, grid, dtype, device):
    with torch.no_grad():
        r = O.forward_frame(pts, synth.kitti_calib(), maps, sd, grid, synth.KITTI_IMSIZE_HW, dtype=dtype, device=device, want_grid=False)
    out = dict(vfeat=r['vfeat'].clone(), idx=r['idx'].cpu())
    del r
    if device != 'cpu':
        torch.cuda.empty_cache()
    return out


@pytest.fixture(scope='module')
def bench_batch():
    """bench.py's rank-0 batch and the oracle's values for every frame of it."""
    assert torch.cuda.is_available()
    sd = synth.make_weights(0)
    frames = [synth.make_points(g, P) for g in range(B)]
    maps = [synth.make_fpn_maps(g) for g in range(B)]
    ref64 = [_oracle(frames[g], maps[g], sd, synth.KITTI_GRID, torch.float64, 'cuda') for g in range(B)]
    cpu32 = _oracle(frames[0], maps[0], sd, synth.KITTI_GRID, torch.float32, 'cpu')
    cpu64 = _oracle(frames[0], maps[0], sd, synth.KITTI_GRID, torch.float64, 'cpu')
    return dict(sd=sd, frames=frames, maps=maps, ref64=ref64, cpu32=cpu32, cpu64=cpu64)


def test_cuda_evaluated_oracle_is_pinned_on_the_cpu_oracle(bench_batch):
    """The fp64 oracle values used for frames 1-7 (torch CUDA kernels) equal the CPU evaluation of the same expressions."""
    b = bench_batch
    assert torch.equal(b['ref64'][0]['idx'], b['cpu64']['idx'])
    assert rel_err(b['ref64'][0]['vfeat'], b['cpu64']['vfeat']) < 1e-9


@pytest.mark.parametrize('fusion_mode', [1, 0])
def test_bench_batch_values_match_oracle(bench_batch, fusion_mode):
    """Every frame of the benchmarked batch-8 / P=120k input: voxel features and dense grid vs the fp64 (and fp32) oracle,
    for the default pixel-first fcn1 (1) and the row-first formulation (0)."""
    from mvxnet_makise_b200 import _lib
    from mvxnet_makise_b200.pipeline import PointPath
    from mvxnet_makise_b200.modules import pack_calib
    b = bench_batch
    offsets = np.concatenate([[0], np.cumsum([p.shape[0] for p in b['frames']])]).tolist()
    pts = torch.from_numpy(np.concatenate(b['frames'], 0)).cuda()
    c32 = torch.stack([pack_calib(synth.kitti_calib()) for _ in range(B)]).cuda()
    maps = [torch.from_numpy(np.concatenate([m[l] for m in b['maps']], 0)).cuda() for l in range(3)]
    _lib.set_fusion_mode(fusion_mode)
    try:
        path = PointPath(b['sd'], synth.KITTI_GRID)
        grid, counts = path.forward_device(pts, offsets, c32, maps)
        torch.cuda.synchronize()
    finally:
        _lib.set_fusion_mode(1)
    counts = counts.cpu().numpy()
    worst = 0.0
    for f in range(B):
        ref = b['ref64'][f]
        n = ref['idx'].shape[0]
        assert counts[f, 0] == n and counts[f, 2] == 0
        vfeat, idx = path.voxel_features(f)
        assert torch.equal(idx.cpu()[:, 1:], ref['idx'][:, 1:]), f'frame {f}: voxel coordinates differ'
        e64 = rel_err(vfeat, ref['vfeat'])
        worst = max(worst, e64)
        assert e64 < TOL, f'frame {f}: voxel features rel err vs fp64 oracle {e64}'
        # dense grid: exact placement of exactly these values, zero elsewhere
        i = ref['idx'].cuda()
        want = torch.zeros_like(grid[f])
        want[:, i[:, 3], i[:, 1], i[:, 2]] = ref['vfeat'].to(torch.float32).T.cuda()
        assert rel_err(grid[f], want) < TOL, f'frame {f}: grid differs'
        assert torch.equal(grid[f][:, i[:, 3], i[:, 1], i[:, 2]].T, vfeat)      # a copy of the voxel features ...
        assert int((grid[f] != 0).sum()) == int((vfeat != 0).sum())             # ... and nothing else
        del want
    # frame 0 against the reference's own fp32 arithmetic: within that reference's distance to fp64 + TOL
    vfeat0, _ = path.voxel_features(0)
    noise = rel_err(b['cpu32']['vfeat'], b['cpu64']['vfeat'])
    e32 = rel_err(vfeat0, b['cpu32']['vfeat'])
    assert e32 <= noise + TOL, f'vs fp32 oracle {e32}, fp32 reference noise {noise}'
    print(f'fusion_mode={fusion_mode}: worst rel err vs fp64 over {B} frames {worst:.3e}; frame 0 vs fp32 {e32:.3e} (fp32 reference noise {noise:.3e})')


def test_dense_config_frame_values_match_oracle():
    """BASELINE.json configs[4]: one dense 128-beam-like frame (P = 250 000, grid 512x512x10) - values, not only placement."""
    from mvxnet_makise_b200.pipeline import PointPath
    DG = synth.DENSE_GRID
    sd = synth.make_weights(0)
    pts = synth.make_points(0, 250_000, grid=DG, beams=128)     # bench.py --workload dense, frame id 0
    maps = synth.make_fpn_maps(0)
    ref = _oracle(pts, maps, sd, DG, torch.float64, 'cuda')
    path = PointPath(sd, DG)
    grid, counts = path([pts], [synth.kitti_calib()], [torch.from_numpy(m) for m in maps])
    torch.cuda.synchronize()
    c = counts.cpu().numpy()[0]
    assert c[0] == ref['idx'].shape[0] and c[2] == 0
    vfeat, idx = path.voxel_features(0)
    assert torch.equal(idx.cpu()[:, 1:], ref['idx'][:, 1:])
    e64 = rel_err(vfeat, ref['vfeat'])
    assert e64 < TOL, f'dense frame: voxel features rel err vs fp64 oracle {e64}'
    i = ref['idx'].cuda()
    assert tuple(grid[0].shape) == (128, 10, 512, 512)
    assert torch.equal(grid[0][:, i[:, 3], i[:, 1], i[:, 2]].T, vfeat)
    assert int((grid[0] != 0).sum()) == int((vfeat != 0).sum())
    print(f'dense frame: rel err vs fp64 {e64:.3e}')
