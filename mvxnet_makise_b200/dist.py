"""Multi-GPU plumbing: one process per GPU, frames sharded across ranks, NO collective in the forward path.

Every frame is independent end to end (the reference is batch-1 with per-frame BatchNorm statistics:
config.yml:18, modules/voxelnet/VoxelNet.py:19), so the forward path is batch-partitioned: frame b runs on rank
b % world. The only exchange step the path has is the training-mode sum of the hot-path layer
gradients (726 880 parameters, 2.9 MB): one all-reduce over a single flat fp32 bucket (NCCL over NVLink on the
GPUs, gloo in the CPU tests) - `training.FlatAdamW.reduce_and_step`, whose bucket the CUDA backward fills directly.
Callers: bench.py (rank / world / frame ids of every workload), training.HotPathTrainer.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR should be 127.0.0.1 on one node)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Global frame ids owned by `rank`: frame b -> rank b % world (round-robin keeps ragged batches balanced)."""
    return list(range(rank, n_frames, world))


def owner_of(frame: int, world: int) -> int:
    return frame % world


def gather_frame_results(local: Sequence, n_frames: int, rank: int, world: int) -> List:
    """All-gather small per-frame python results (counts, checksums) back into global frame order."""
    if world == 1 or not dist.is_initialized():
        return list(local)
    parts = [None] * world
    dist.all_gather_object(parts, list(local))
    out = [None] * n_frames
    for r, part in enumerate(parts):
        for b, item in zip(shard_frames(n_frames, r, world), part):
            out[b] = item
    return out
