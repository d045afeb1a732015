import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from mvxnet_makise_b200 import synth
from mvxnet_makise_b200.pipeline import PointPath
G=synth.KITTI_GRID
sd=synth.make_weights(1); calib=synth.kitti_calib()
rng=np.random.default_rng(0)
maps=[torch.from_numpy(rng.standard_normal((3,256,h,w),dtype=np.float32)) for (h,w) in ((13,42),(7,21),(4,11))]
frames=[synth.make_points(1,700), np.zeros((0,4),np.float32), synth.make_points(2,1)]
path=PointPath(sd,G)
grid,counts=path(frames,[calib]*3,maps)
torch.cuda.synchronize()
print('counts',counts.cpu().numpy().tolist())
for f in range(3):
    vf,idx=path.voxel_features(f)
    print(f, vf.shape, bool(torch.isfinite(vf).all()), float(grid[f].abs().sum()), bool(torch.isfinite(grid[f]).all()))
# all-empty batch
path2=PointPath(sd,G)
g2,c2=path2([np.zeros((0,4),np.float32)],[calib],[m[:1] for m in maps])
torch.cuda.synchronize(); print('empty batch', c2.cpu().numpy().tolist(), float(g2.abs().sum()))
