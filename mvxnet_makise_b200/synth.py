"""Synthetic KITTI-shaped frames.

Pure numpy, no GPU. The frames honour the invariants the reference's offline
``cropdata.py`` guarantees (cropdata.py:26-67; modules/data/Preprocessing.py:12-55):
every point is inside ``velorange``, in front of the camera and inside the image
(with the reference's ``-1e-3`` image-bound fudge), so ``reindex`` and the gather
bounds assert (modules/imhead/Pipe.py:71) hold.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Sequence, Tuple

import numpy as np

# config.yml:3-13,21,24-26 of the reference (the KITTI config) --------------------------
KITTI_VELORANGE = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
KITTI_VOXELSHAPE = (352, 400, 10)          # (x, y, z) cells
KITTI_IMSIZE_HW = (370, 1224)              # config.yml imsize is (h, w)
KITTI_T = 35
# BASELINE.json configs[4] (builder-defined, SURVEY.md §8d "config 5")
DENSE_VELORANGE = (0.0, -51.2, -3.0, 102.4, 51.2, 1.0)
DENSE_VOXELSHAPE = (512, 512, 10)

FPN_CHANNELS = 256


@dataclasses.dataclass(frozen=True)
class GridSpec:
    """Voxel grid description; ``voxelsize`` follows modules/config/Config.py:7 (python doubles)."""
    velorange: Tuple[float, ...] = KITTI_VELORANGE
    voxelshape: Tuple[int, ...] = KITTI_VOXELSHAPE
    T: int = KITTI_T

    @property
    def voxelsize(self) -> Tuple[float, float, float]:
        r, s = self.velorange, self.voxelshape
        return tuple((r[i + 3] - r[i]) / s[i] for i in range(3))

    @property
    def cells(self) -> int:
        s = self.voxelshape
        return s[0] * s[1] * s[2]


KITTI_GRID = GridSpec()
DENSE_GRID = GridSpec(DENSE_VELORANGE, DENSE_VOXELSHAPE, KITTI_T)


def kitti_calib() -> Dict[str, np.ndarray]:
    """KITTI-000000-like calibration, padded to 4x4 like modules/data/Load.py:24-41, fp32."""
    p2 = np.array([[721.5377, 0.0, 609.5593, 44.85728],
                   [0.0, 721.5377, 172.854, 0.2163791],
                   [0.0, 0.0, 1.0, 0.002745884],
                   [0.0, 0.0, 0.0, 1.0]], dtype=np.float32)
    r0 = np.eye(4, dtype=np.float32)
    tr = np.array([[7.533745e-03, -9.999714e-01, -6.166020e-04, -4.069766e-03],
                   [1.480249e-02, 7.280733e-04, -9.998902e-01, -7.631618e-02],
                   [9.998621e-01, 7.523790e-03, 1.480755e-02, -2.717806e-01],
                   [0.0, 0.0, 0.0, 1.0]], dtype=np.float32)
    return {'P2': p2, 'R0_rect': r0, 'Tr_velo_to_cam': tr}


def kitti_calib_f64() -> Dict[str, np.ndarray]:
    """The same calibration as `readCalib` leaves it in memory (modules/data/Load.py:24-41): the fp32 file values inside
    FLOAT64 matrices (`np.concatenate([fp32 (3,4), [[0, 0, 0, 1]]])` and `np.zeros((4, 4))` are float64). The reference's numpy
    branches (cropToSight at Load.py:73, lidar2Img of the pasted sets at train.py:36-39) then compute in fp64."""
    c = kitti_calib()
    v2c = np.concatenate([c['Tr_velo_to_cam'][:3], [[0, 0, 0, 1]]], axis=0)
    p2 = np.concatenate([c['P2'][:3], [[0, 0, 0, 1]]], axis=0)
    r0 = np.zeros((4, 4))
    r0[:3, :3] = c['R0_rect'][:3, :3]
    r0[3, 3] = 1
    assert v2c.dtype == np.float64 and p2.dtype == np.float64 and r0.dtype == np.float64
    return {'P2': p2, 'R0_rect': r0, 'Tr_velo_to_cam': v2c}


def fpn_shapes(imsize_hw: Sequence[int] = KITTI_IMSIZE_HW):
    """Shapes of FPN levels '0','1','2' for a KITTI image (probed):
    the torchvision transform resizes to min side 800 and pads to /32 -> 416x1344."""
    if tuple(imsize_hw) == KITTI_IMSIZE_HW or tuple(imsize_hw) == (375, 1242):
        return [(104, 336), (52, 168), (26, 84)]
    h, w = imsize_hw
    scale = min(800.0 / min(h, w), 1333.0 / max(h, w))
    hp = int(np.ceil(int(h * scale) / 32.0) * 32)
    wp = int(np.ceil(int(w * scale) / 32.0) * 32)
    return [(hp // 4, wp // 4), (hp // 8, wp // 8), (hp // 16, wp // 16)]


def _crop(pcd, velorange):
    low = np.array(velorange[0:3])
    high = np.array(velorange[3:6])
    roi = pcd[:, :3]
    return pcd[np.all((low <= roi) & (roi < high), axis=1)]


def _crop_to_sight(pcd, calib, imsize_wh):
    """Same filter as the numpy branch of cropToSight (Preprocessing.py:28-55)."""
    lim = np.array(imsize_wh) - 1e-3
    pts = np.empty((4, pcd.shape[0]), dtype='float32')
    pts[:3] = pcd.T[:3]
    pts[3] = 1
    pts = calib['R0_rect'] @ calib['Tr_velo_to_cam'] @ pts
    f = pts[2] > 0
    pcd = pcd[f]
    pts = pts[:, f]
    pts = calib['P2'] @ pts
    pts[:2] = pts[:2] / pts[2]
    uv = pts[:2].T
    f = np.all(uv >= 0, axis=1) & np.all(uv < lim, axis=1)
    return pcd[f]


def make_points(frame_id: int, P: int = 120_000, grid: GridSpec = KITTI_GRID, beams: int = 64,
                imsize_hw: Sequence[int] = KITTI_IMSIZE_HW, calib=None) -> np.ndarray:
    """(P,4) fp32 [x,y,z,reflectance] in a fixed, already-permuted order (the order the
    in-function shuffle of ``group`` would have produced; SURVEY.md trap 3)."""
    calib = calib or kitti_calib()
    rng = np.random.default_rng(1000 * frame_id + 7)
    out = np.zeros((0, 4), dtype=np.float32)
    draw = max(3 * P, 1024)
    for _ in range(8):
        az = np.deg2rad(rng.uniform(-42.0, 42.0, draw))
        el = np.deg2rad(np.linspace(-24.8, 2.0, beams))[rng.integers(0, beams, draw)]
        with np.errstate(divide='ignore'):
            rg = np.where(el < 0, 1.73 / np.sin(-el), 200.0)
        obst = rng.uniform(5.0, 70.0, draw)
        hit = rng.uniform(0.0, 1.0, draw) < 0.35
        rg = np.where(hit, np.minimum(rg, obst), rg)
        rg = rg + rng.normal(0.0, 0.02, draw)
        x = rg * np.cos(el) * np.cos(az)
        y = rg * np.cos(el) * np.sin(az)
        z = rg * np.sin(el)
        refl = rng.uniform(0.0, 1.0, draw)
        pcd = np.stack([x, y, z, refl], axis=1).astype(np.float32)
        pcd = _crop(pcd, grid.velorange)
        pcd = _crop_to_sight(pcd, calib, (imsize_hw[1], imsize_hw[0]))
        out = np.concatenate([out, pcd], axis=0)
        if out.shape[0] >= P:
            break
    out = out[:P]
    out = out[rng.permutation(out.shape[0])]
    return np.ascontiguousarray(out, dtype=np.float32)


def make_raw_sweep(frame_id: int, P: int = 120_000, beams: int = 64) -> np.ndarray:
    """(P,4) fp32 UNcropped 360-degree sweep (what a raw KITTI .bin holds, Load.py:57): returns behind the camera, beyond the
    velo range and outside the image are all present - the input of `crop` / `cropToSight`."""
    rng = np.random.default_rng(1000 * frame_id + 13)
    az = np.deg2rad(rng.uniform(-180.0, 180.0, P))
    el = np.deg2rad(np.linspace(-24.8, 4.0, beams))[rng.integers(0, beams, P)]
    with np.errstate(divide='ignore'):
        rg = np.where(el < 0, 1.73 / np.sin(-el), 150.0)
    rg = np.where(rng.uniform(0.0, 1.0, P) < 0.4, np.minimum(rg, rng.uniform(2.0, 120.0, P)), rg)
    rg = rg + rng.normal(0.0, 0.02, P)
    pcd = np.stack([rg * np.cos(el) * np.cos(az), rg * np.cos(el) * np.sin(az), rg * np.sin(el), rng.uniform(0.0, 1.0, P)], axis=1)
    return np.ascontiguousarray(pcd, dtype=np.float32)


def make_fpn_maps(frame_id: int, imsize_hw: Sequence[int] = KITTI_IMSIZE_HW, channels: int = FPN_CHANNELS):
    """Random stand-ins for FPN levels '0','1','2' (NCHW fp32, batch 1) for stage-isolated runs."""
    rng = np.random.default_rng(1000 * frame_id + 11)
    return [rng.standard_normal((1, channels, h, w), dtype=np.float32) for (h, w) in fpn_shapes(imsize_hw)]


HOT_LAYERS = (  # (state-dict prefix, Cin, Cout, is_conv) — SURVEY.md §8b checkpoint table
    ('head.fusion.fcn1.fc', 768, 768, False),
    ('head.fusion.conv1.conv', 768, 128, True),
    ('head.fusion.fcn2.fc', 128, 128, False),
    ('head.fusion.conv2.conv', 128, 16, True),
    ('head.fusion.fcn3.fc', 16, 16, False),
    ('backbone.svfe.vfe1.fcn.fc', 23, 16, False),
    ('backbone.svfe.vfe2.fcn.fc', 32, 64, False),
    ('backbone.fcn.fc', 128, 128, False),
)


def make_weights(seed: int = 0) -> Dict[str, np.ndarray]:
    """Random hot-path weights with the reference's names/shapes (uniform(-1/sqrt(Cin), 1/sqrt(Cin)),
    the nn.Linear / nn.Conv2d default-init range). Not bit-identical to torch's init stream —
    parity tests copy the *same* arrays into both sides."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, cin, cout, is_conv in HOT_LAYERS:
        k = 1.0 / np.sqrt(cin)
        w = rng.uniform(-k, k, (cout, cin)).astype(np.float32)
        b = rng.uniform(-k, k, (cout,)).astype(np.float32)
        sd[name + '.weight'] = w.reshape(cout, cin, 1, 1) if is_conv else w
        sd[name + '.bias'] = b
    return sd
