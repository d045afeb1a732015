"""BASELINE configs[2] as a parity case: full MVXNet inference, random-init weights, batch 8, with the hot path swapped in.

Image branch = torchvision's Faster R-CNN ResNet50-FPN-v2 transform + backbone exactly as modules/imhead/Pipe.py:8-21 builds it
(random init: no network), run on the GPU in eval mode (Head.py:10); its FPN levels '0','1','2' feed `PointPath` (the path under
test) next to raw points and calibration. After the path, the reference's middle layers and RPN (modules/voxelnet/Pipe.py:31-80,
VoxelNet.py:34-38) are restated below as plain torch modules (test scaffolding — they are consumers, not part of the product).
Checker: the oracle port on the SAME FPN maps (copied to the host), frame by frame like the batch-1 reference."""
import numpy as np
import pytest
import torch

from _util import same_occupancy
from torch import nn
from torch.nn import functional as F

pytestmark = pytest.mark.gpu
EPS = 1e-6   # config.yml eps


class CRB(nn.Module):   # Blocks.py:20-53: relu(conv) then batch-statistic, affine-free BatchNorm
    def __init__(self, conv, bn):
        super().__init__()
        self.conv, self.bn = conv, bn

    def forward(self, x):
        return self.bn(F.relu(self.conv(x)))


def crb3d(cin, cout, k, s, p):
    return CRB(nn.Conv3d(cin, cout, k, s, p), nn.BatchNorm3d(cout, eps=EPS, affine=False, track_running_stats=False))


def crb2d(cin, cout, k, s, p):
    return CRB(nn.Conv2d(cin, cout, k, s, p), nn.BatchNorm2d(cout, eps=EPS, affine=False, track_running_stats=False))


def decrb2d(cin, cout, k, s, p):
    return CRB(nn.ConvTranspose2d(cin, cout, k, s, p), nn.BatchNorm2d(cout, eps=EPS, affine=False, track_running_stats=False))


class MiddleAndRPN(nn.Module):   # voxelnet/Pipe.py:31-80 + VoxelNet.py:34-38
    def __init__(self, nx, ny):
        super().__init__()
        self.nx, self.ny = nx, ny
        self.cml = nn.Sequential(crb3d(128, 64, 3, (2, 1, 1), (1, 1, 1)), crb3d(64, 64, 3, 1, (0, 1, 1)), crb3d(64, 64, 3, (2, 1, 1), 1))
        self.blk1 = nn.Sequential(crb2d(128, 128, 3, 2, 1), *[crb2d(128, 128, 3, 1, 1) for _ in range(3)])
        self.blk2 = nn.Sequential(crb2d(128, 128, 3, 2, 1), *[crb2d(128, 128, 3, 1, 1) for _ in range(5)])
        self.blk3 = nn.Sequential(crb2d(128, 256, 3, 2, 1), *[crb2d(256, 256, 3, 1, 1) for _ in range(5)])
        self.deconv1, self.deconv2, self.deconv3 = decrb2d(128, 256, 3, 1, 1), decrb2d(128, 256, 2, 2, 0), decrb2d(256, 256, 4, 4, 0)
        self.cls, self.reg = nn.Conv2d(768, 2, 1, 1, 0), nn.Conv2d(768, 14, 1, 1, 0)
        for m in self.modules():   # MVXNet.py:8-11
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight.data)
                m.bias.data.zero_()

    def forward(self, grid):   # grid (1,128,nz,nx,ny)
        x = self.cml(grid).reshape((1, -1, self.nx, self.ny))
        x1 = self.blk1(x)
        x2 = self.blk2(x1)
        x3 = self.blk3(x2)
        x = torch.concat([self.deconv1(x1), self.deconv2(x2), self.deconv3(x3)], dim=1)
        return torch.sigmoid(self.cls(x)), self.reg(x)


def rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(b.double().abs().max().item(), 1e-30)


def test_full_mvxnet_inference_batch8_hot_path_swapped_in():
    from torchvision.models.detection import fasterrcnn_resnet50_fpn_v2
    from mvxnet_makise_b200 import synth
    from mvxnet_makise_b200.pipeline import PointPath
    from oracle import pointpath_oracle as O

    dev = torch.device('cuda')
    B = 8
    torch.manual_seed(0)
    frcnn = fasterrcnn_resnet50_fpn_v2(weights=None, weights_backbone=None)   # Pipe.py:8 with random init
    transform, backbone = frcnn.transform.to(dev).eval(), frcnn.backbone.to(dev).eval()
    tail = MiddleAndRPN(synth.KITTI_GRID.voxelshape[0], synth.KITTI_GRID.voxelshape[1]).to(dev)
    sd = synth.make_weights(0)
    H, W = synth.KITTI_IMSIZE_HW
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)                  # Load.py:61-62: 375x1242 cropped to 370x1224
    with torch.no_grad():
        x = torch.from_numpy(imgs).to(dev).float().permute(0, 3, 1, 2) / 255   # train.py:128
        feats = backbone(transform(x)[0].tensors)
        maps = [feats[k].contiguous() for k in ('0', '1', '2')]
    assert [tuple(m.shape[1:]) for m in maps] == [(256, h, w) for h, w in synth.fpn_shapes()]
    # frames 0 and 1 are thinned so that the CPU checker finishes in seconds; the other six are full size
    frames = [synth.make_points(200 + f, 30_000 if f < 2 else 120_000) for f in range(B)]
    calib = synth.kitti_calib()
    path = PointPath(sd, synth.KITTI_GRID)
    grids, counts = path(frames, [calib] * B, maps)
    torch.cuda.synchronize()
    counts = counts.cpu().numpy()
    assert np.all(counts[:, 2] == 0) and np.all(counts[2:, 0] > 15_000)

    for f in range(2):
        host_maps = [m[f:f + 1].cpu().numpy() for m in maps]
        with torch.no_grad():
            ref64 = O.forward_frame(frames[f], calib, host_maps, sd, synth.KITTI_GRID, synth.KITTI_IMSIZE_HW, dtype=torch.float64)
            ref32 = O.forward_frame(frames[f], calib, host_maps, sd, synth.KITTI_GRID, synth.KITTI_IMSIZE_HW, dtype=torch.float32)
        n = ref64['idx'].shape[0]
        assert counts[f, 0] == n
        vfeat, idx = path.voxel_features(f)
        assert np.array_equal(idx.cpu().numpy()[:, 1:], ref64['idx'].numpy()[:, 1:])             # bit-exact voxel coordinates
        e_ours, e_ref = rel(vfeat.cpu(), ref64['vfeat']), rel(ref32['vfeat'], ref64['vfeat'])
        print(f'frame {f}: N={n} voxel features vs fp64: ours {e_ours:.2e}, fp32 reference {e_ref:.2e}')
        assert e_ours < 1e-4                                                                       # the bar (north_star)
        g = grids[f]
        assert same_occupancy(g.cpu(), ref64['grid'][0]) and rel(g.cpu(), ref64['grid'][0]) < 1e-4
        if f == 0:   # downstream: middle layers + RPN on our grid and on the checker's grids (same torch modules, same GPU)
            with torch.no_grad():
                s_ours, r_ours = tail(g[None])
                s_64, r_64 = tail(ref64['grid'].float().to(dev))
                s_32, r_32 = tail(ref32['grid'].float().to(dev))
            assert s_ours.shape == (1, 2, 176, 200) and r_ours.shape == (1, 14, 176, 200)
            d_ours = max(rel(s_ours, s_64), rel(r_ours, r_64))
            d_ref = max(rel(s_32, s_64), rel(r_32, r_64))
            print(f'score/reg after CML+RPN vs the fp64-grid run: ours {d_ours:.2e}, fp32 reference grid {d_ref:.2e}')
            # 29 batch-statistic BatchNorm layers amplify input differences; the hot path must not add more than the
            # fp32 reference's own rounding does (plus the bar)
            assert d_ours <= d_ref + 1e-3

    # the six full-size frames: size-independent properties (occupancy = voxel list, features finite, BN-bounded)
    for f in range(2, B):
        vfeat, idx = path.voxel_features(f)
        g = grids[f]
        assert torch.isfinite(vfeat).all() and int((g.abs().sum(0) != 0).sum()) <= counts[f, 0]
        lin = (idx[:, 3] * 352 + idx[:, 1]) * 400 + idx[:, 2]
        assert torch.equal(g.reshape(128, -1)[:, lin].t(), vfeat)
