"""mvxnet_makise_b200 — B200-native (sm_100a) point-side hot path of MVXNet behind the reference's own API.

Submodules:
  synth      synthetic KITTI-shaped frames / weights (pure numpy, no GPU)
  voxelize   `cpp._group`, `group`, `group_` drop-ins (stage 1)
  modules    `lidar2Img`, `featureMaping`, `FCN`, `CRB2d`, `ImageFeatureFusion`, `VFE`, `SVFE`, `reindex` drop-ins
  pipeline   `PointPath`: the fused, batched stages 1-4
  dist       frame sharding across GPUs (one process per GPU)
Everything except `synth` needs the in-tree CUDA library (`libmvx_b200.so`) and fails loudly without it.
"""
__version__ = '0.1.0'
