"""Dev tool (CPU): validates the COMPACT backward formulas (weighted pad rows, argmax routing, BatchNorm batch-stat
backward with multiplicities) against torch autograd through the dense oracle chain. The CUDA backward kernels mirror
`compact_backward` below."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvxnet_makise_b200 import synth
from oracle import pointpath_oracle as O

torch.manual_seed(0)
G = synth.KITTI_GRID
EPS = 1e-6
NAMES = [n for n, *_ in synth.HOT_LAYERS]


def dense_forward_autograd(x768, vox7, sd, Gw):
    """x768 (N,T,768), vox7 (N,T,7) dense fp64; returns loss and grads wrt all weights via autograd."""
    params = {k: v.clone().double().requires_grad_(True) for k, v in sd.items()}
    x = O.fusion(x768[None], params, EPS)
    x23 = torch.concat([vox7[None], x], dim=-1)
    vf = O.voxel_features(x23, params, EPS)           # (N,128)
    loss = (vf * Gw).sum()
    loss.backward()
    return vf.detach(), {k: p.grad for k, p in params.items()}


def compact_forward(A1, vox7c, sd, cnt, row_v, T):
    """Compact forward in fp64 mirroring the CUDA path. A1 (K+1,768) with the pad row last (zeros)."""
    N, K = len(cnt), A1.shape[0] - 1
    R = N * T
    wA = torch.ones(K + 1, dtype=torch.float64); wA[K] = R - K
    st = {}
    def layer(x, name, w):
        W = sd[name + '.weight'].double().reshape(sd[name + '.weight'].shape[0], -1); b = sd[name + '.bias'].double()
        pre = x @ W.t() + b
        y = torch.relu(pre)
        mu = (w[:, None] * y).sum(0) / R
        var = (w[:, None] * y * y).sum(0) / R - mu * mu
        rstd = 1.0 / torch.sqrt(var + EPS)
        z = (y - mu) * rstd
        return dict(x=x, W=W, pre=pre, y=y, z=z, rstd=rstd, w=w)
    x = A1
    for i in range(5):
        st[i] = layer(x, NAMES[i], wA); x = st[i]['z']
    X6 = torch.cat([vox7c, x], 1)                      # pad row: vox part zero
    st[5] = layer(X6, NAMES[5], wA)
    z6 = st[5]['z']
    has_pad = torch.tensor(cnt < T)
    def vmax(z_real, z_pad_per_v, has):                # per-voxel max over real rows and (optionally) the pad value
        M = torch.full((N, z_real.shape[1]), -1e300, dtype=torch.float64)
        M.index_reduce_(0, row_v, z_real, 'amax', include_self=True)
        return torch.where(has[:, None], torch.maximum(M, z_pad_per_v), M)
    M6 = vmax(z6[:K], z6[K][None].expand(N, -1), has_pad)
    wB = torch.cat([torch.ones(K, dtype=torch.float64), torch.tensor(T - cnt, dtype=torch.float64)])
    X7 = torch.cat([torch.cat([z6[:K], M6[row_v]], 1), torch.cat([z6[K][None].expand(N, -1), M6], 1)], 0)
    st[6] = layer(X7, NAMES[6], wB)
    z7 = st[6]['z']
    M7 = vmax(z7[:K], z7[K:], has_pad)
    vB = torch.cat([row_v, torch.arange(N)])
    X8 = torch.cat([z7, M7[vB]], 1)
    st[7] = layer(X8, NAMES[7], wB)
    z8 = st[7]['z']
    out = vmax(z8[:K], z8[K:], has_pad)
    return out, st, dict(M6=M6, M7=M7, wA=wA, wB=wB, vB=vB, has_pad=has_pad, R=R, K=K, N=N)


def route_max(dM, z_real, z_pad_per_v, row_v, has_pad, N):
    """dM (N,C) -> (d z_real (K,C), d pad-per-voxel (N,C)): gradient goes to the FIRST real row attaining the max
    (slot order) if it is >= the pad value, else to the pad slot."""
    K, C = z_real.shape
    M_real = torch.full((N, C), -1e300, dtype=torch.float64).index_reduce_(0, row_v, z_real, 'amax', include_self=True)
    pad_wins = has_pad[:, None] & (z_pad_per_v > M_real)
    # first real row with z == M_real: rows are voxel-major/slot order, so the smallest row index wins
    is_max = z_real == M_real[row_v]
    ridx = torch.arange(K)[:, None].expand(K, C)
    first = torch.full((N, C), K, dtype=torch.long).scatter_reduce_(0, row_v[:, None].expand(K, C), torch.where(is_max, ridx, K), 'amin')
    dz = torch.zeros_like(z_real)
    sel = (~pad_wins)
    cols = torch.arange(C)[None].expand(N, C)
    dz[first[sel], cols[sel]] = dM[sel]
    dpad = torch.where(pad_wins, dM, torch.zeros_like(dM))
    return dz, dpad


def compact_backward(st, aux, dOut, row_v):
    K, N, R = aux['K'], aux['N'], aux['R']
    grads = {}
    def layer_bwd(i, dz_sum):
        s = st[i]
        m1 = dz_sum.sum(0) / R
        m2 = (dz_sum * s['z']).sum(0) / R
        dy = s['rstd'] * (dz_sum - s['w'][:, None] * m1 - s['w'][:, None] * s['z'] * m2)
        dpre = dy * (s['pre'] > 0)
        grads[NAMES[i] + '.weight'] = dpre.t() @ s['x']
        grads[NAMES[i] + '.bias'] = dpre.sum(0)
        return dpre @ s['W']
    z8, z7, z6 = st[7]['z'], st[6]['z'], st[5]['z']
    dzr, dzp = route_max(dOut, z8[:K], z8[K:], row_v, aux['has_pad'], N)
    dX8 = layer_bwd(7, torch.cat([dzr, dzp], 0))
    dz7 = dX8[:, :64].clone()
    dM7 = torch.zeros(N, 64, dtype=torch.float64).index_add_(0, aux['vB'], dX8[:, 64:])
    dzr, dzp = route_max(dM7, z7[:K], z7[K:], row_v, aux['has_pad'], N)
    dz7[:K] += dzr; dz7[K:] += dzp
    dX7 = layer_bwd(6, dz7)
    dz6 = torch.zeros(K + 1, 16, dtype=torch.float64)
    dz6[:K] = dX7[:K, :16]
    dz6[K] = dX7[K:, :16].sum(0)
    dM6 = torch.zeros(N, 16, dtype=torch.float64).index_add_(0, aux['vB'], dX7[:, 16:])
    dzr, dzp = route_max(dM6, z6[:K], z6[K][None].expand(N, -1), row_v, aux['has_pad'], N)
    dz6[:K] += dzr; dz6[K] += dzp.sum(0)
    dX6 = layer_bwd(5, dz6)
    d = dX6[:, 7:23]
    for i in (4, 3, 2, 1, 0):
        d = layer_bwd(i, d)
    return grads


def main():
    P = 1200
    pts = synth.make_points(11, P)
    maps = [np.random.default_rng(1).standard_normal((1, 256, h, w), dtype=np.float32) for (h, w) in ((13, 42), (7, 21), (4, 11))]
    sd_np = synth.make_weights(4)
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    pcd6 = O.points_with_proj(pts, synth.kitti_calib())
    vox9, uidx = O.group(pcd6, G.velorange, G.voxelsize, G.T)
    voxels = torch.Tensor(vox9)
    im768 = O.feature_mapping(voxels, [torch.from_numpy(m) for m in maps], torch.Tensor(list(synth.KITTI_IMSIZE_HW))).double()
    vox7 = voxels[..., :7].double()
    N, T = voxels.shape[0], G.T
    Gw = torch.randn(N, 128, dtype=torch.float64)
    vf, gref = dense_forward_autograd(im768, vox7, sd, Gw)
    cnt = (voxels[..., :3] != 0).any(-1).sum(1).numpy()
    rows = np.concatenate([v * T + np.arange(c) for v, c in enumerate(cnt)])
    row_v = torch.from_numpy(np.concatenate([np.full(c, v) for v, c in enumerate(cnt)]))
    A1 = torch.cat([im768.reshape(-1, 768)[rows], torch.zeros(1, 768, dtype=torch.float64)], 0)
    vox7c = torch.cat([vox7.reshape(-1, 7)[rows], torch.zeros(1, 7, dtype=torch.float64)], 0)
    out, st, aux = compact_forward(A1, vox7c, sd, cnt, row_v, T)
    print('forward max diff', (out - vf).abs().max().item())
    grads = compact_backward(st, aux, Gw, row_v)
    for k in sorted(gref):
        g = grads[k].reshape(gref[k].shape)
        print(f'{k:34s} rel err {((g - gref[k]).abs().max() / gref[k].abs().max()).item():.3e}')


if __name__ == '__main__':
    main()
