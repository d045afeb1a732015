"""TEST INFRASTRUCTURE (oracle side) - fp64 restatement of the COMPACT training-mode formulas the CUDA backward
implements (mvxnet_makise_b200/csrc/backward.cu), in plain torch on the CPU.

Why it exists: the gradient of this path is only piecewise continuous in the activations (ReLU masks, the argmax of the
max over T). Two forwards that agree to 1e-5 disagree on a handful of those decisions, and each flipped decision moves a
whole O(1) gradient entry between rows; the fp32 reference's own autograd is 1e-2 .. 6e-2 (max-norm, per tensor) away
from its fp64 evaluation for exactly this reason. So gradient parity is pinned in two steps:
  (1) `compact_forward` + `compact_backward` == autograd through the dense reference chain (oracle/pointpath_oracle.py:
      Pipe.py:84-104, voxelnet/Pipe.py:5-29, VoxelNet.py:27-32) in fp64, to 1e-9   [tests/test_oracle.py, CPU];
  (2) the CUDA backward == `compact_backward` evaluated in fp64 on the CUDA forward's OWN saved activations (same
      decisions), to 1e-4                                                         [tests/test_gpu_backward.py, GPU].

Compact formulation (same as the forward): rows = the K kept points, voxel-major in slot order, plus
weighted pad rows - ONE per frame (multiplicity N*T-K) for the fusion stack and VFE1, one per voxel (multiplicity
T-cnt) for VFE2 and the last FCN."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

LAYERS = ('head.fusion.fcn1.fc', 'head.fusion.conv1.conv', 'head.fusion.fcn2.fc', 'head.fusion.conv2.conv',
          'head.fusion.fcn3.fc', 'backbone.svfe.vfe1.fcn.fc', 'backbone.svfe.vfe2.fcn.fc', 'backbone.fcn.fc')
F64 = torch.float64


def _vmax(z_real, z_pad_per_v, has_pad, row_v, N):
    """per-voxel max over the real rows and (where the voxel has pad slots) the pad value"""
    M = torch.full((N, z_real.shape[1]), -1e300, dtype=F64)
    M.index_reduce_(0, row_v, z_real, 'amax', include_self=True)
    return torch.where(has_pad[:, None], torch.maximum(M, z_pad_per_v), M)


def compact_forward(A1, vox7c, sd, cnt, row_v, T, eps=1e-6, y_given: Optional[Dict[int, torch.Tensor]] = None):
    """A1 (K+1,768) gathered image features with the all-zero frame pad row last; vox7c (K+1,7) voxel columns (pad row
    zero); cnt (N,) kept points per voxel; row_v (K,) voxel of each row. Returns (vfeat (N,128), per-layer state, aux).
    y_given[l] replaces layer l's raw output relu(pre) (e.g. by the CUDA forward's saved activations): everything
    downstream, including statistics, ReLU masks and argmax decisions, then follows those values."""
    cnt = torch.as_tensor(np.asarray(cnt)).long()
    N, K = len(cnt), A1.shape[0] - 1
    R = N * T
    wA = torch.ones(K + 1, dtype=F64)
    wA[K] = R - K
    st = {}

    def layer(i, x, w):
        W = sd[LAYERS[i] + '.weight'].to(F64)
        W = W.reshape(W.shape[0], -1)
        b = sd[LAYERS[i] + '.bias'].to(F64)
        pre = x @ W.t() + b
        y = torch.relu(pre)
        if y_given is not None and i in y_given:
            y = y_given[i].to(F64)
            pre = y                                  # only its sign is used (y > 0 <=> pre > 0)
        mu = (w[:, None] * y).sum(0) / R
        var = ((w[:, None] * y * y).sum(0) / R - mu * mu).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        st[i] = dict(x=x, W=W, pre=pre, y=y, z=(y - mu) * rstd, rstd=rstd, w=w)
        return st[i]['z']

    x = A1.to(F64)
    for i in range(5):
        x = layer(i, x, wA)
    z6 = layer(5, torch.cat([vox7c.to(F64), x], 1), wA)
    has_pad = cnt < T
    M6 = _vmax(z6[:K], z6[K][None].expand(N, -1), has_pad, row_v, N)
    wB = torch.cat([torch.ones(K, dtype=F64), (T - cnt).to(F64)])
    vB = torch.cat([row_v, torch.arange(N)])
    X7 = torch.cat([torch.cat([z6[:K], M6[row_v]], 1), torch.cat([z6[K][None].expand(N, -1), M6], 1)], 0)
    z7 = layer(6, X7, wB)
    M7 = _vmax(z7[:K], z7[K:], has_pad, row_v, N)
    z8 = layer(7, torch.cat([z7, M7[vB]], 1), wB)
    out = _vmax(z8[:K], z8[K:], has_pad, row_v, N)
    return out, st, dict(wA=wA, wB=wB, vB=vB, has_pad=has_pad, R=R, K=K, N=N)


def route_max(dM, z_real, z_pad_per_v, row_v, has_pad, N):
    """dM (N,C) -> (d z_real (K,C), d pad-per-voxel (N,C)): the gradient of a per-voxel max goes to the FIRST real row
    (slot order) attaining it, or to the pad slot when that is strictly larger - torch.max's argmax on the dense
    (N,T,C) tensor, where the real slots precede the pad slots."""
    K, C = z_real.shape
    M_real = torch.full((N, C), -1e300, dtype=F64).index_reduce_(0, row_v, z_real, 'amax', include_self=True)
    pad_wins = has_pad[:, None] & (z_pad_per_v > M_real)
    is_max = z_real == M_real[row_v]
    ridx = torch.arange(K)[:, None].expand(K, C)
    first = torch.full((N, C), K, dtype=torch.long).scatter_reduce_(0, row_v[:, None].expand(K, C),
                                                                    torch.where(is_max, ridx, K), 'amin')
    dz = torch.zeros_like(z_real)
    sel = ~pad_wins
    cols = torch.arange(C)[None].expand(N, C)
    dz[first[sel], cols[sel]] = dM[sel]
    return dz, torch.where(pad_wins, dM, torch.zeros_like(dM))


def compact_backward(st, aux, dOut, row_v):
    """Gradients {state-dict name: tensor} of sum(vfeat * dOut) w.r.t. the weights/biases of the 8 layers."""
    K, N, R = aux['K'], aux['N'], aux['R']
    grads = {}

    def layer_bwd(i, dz_sum):
        s = st[i]
        m1 = dz_sum.sum(0) / R
        m2 = (dz_sum * s['z']).sum(0) / R
        dy = s['rstd'] * (dz_sum - s['w'][:, None] * m1 - s['w'][:, None] * s['z'] * m2)
        dpre = dy * (s['pre'] > 0)
        grads[LAYERS[i] + '.weight'] = dpre.t() @ s['x']
        grads[LAYERS[i] + '.bias'] = dpre.sum(0)
        return dpre @ s['W']

    z8, z7, z6 = st[7]['z'], st[6]['z'], st[5]['z']
    dOut = torch.as_tensor(dOut).to(F64)
    dzr, dzp = route_max(dOut, z8[:K], z8[K:], row_v, aux['has_pad'], N)
    dX8 = layer_bwd(7, torch.cat([dzr, dzp], 0))
    dz7 = dX8[:, :64].clone()
    dM7 = torch.zeros(N, 64, dtype=F64).index_add_(0, aux['vB'], dX8[:, 64:])
    dzr, dzp = route_max(dM7, z7[:K], z7[K:], row_v, aux['has_pad'], N)
    dz7[:K] += dzr
    dz7[K:] += dzp
    dX7 = layer_bwd(6, dz7)
    dz6 = torch.zeros(K + 1, 16, dtype=F64)
    dz6[:K] = dX7[:K, :16]
    dz6[K] = dX7[K:, :16].sum(0)
    dM6 = torch.zeros(N, 16, dtype=F64).index_add_(0, aux['vB'], dX7[:, 16:])
    dzr, dzp = route_max(dM6, z6[:K], z6[K][None].expand(N, -1), row_v, aux['has_pad'], N)
    dz6[:K] += dzr
    dz6[K] += dzp.sum(0)
    d = layer_bwd(5, dz6)[:, 7:23]
    for i in (4, 3, 2, 1, 0):
        d = layer_bwd(i, d)
    return grads


def compact_rows(voxels9: torch.Tensor, T: int):
    """From the dense (N,T,9) voxel tensor: cnt (N,), dense row index of every kept point (K,), voxel of every kept row."""
    cnt = (voxels9[..., :3] != 0).any(-1).sum(1).numpy()
    rows = np.concatenate([v * T + np.arange(c) for v, c in enumerate(cnt)])
    row_v = torch.from_numpy(np.concatenate([np.full(c, v) for v, c in enumerate(cnt)])).long()
    return cnt, rows, row_v
