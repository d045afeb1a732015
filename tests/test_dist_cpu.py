"""world_size-2 gloo tests (CPU) of the N>1 host logic: frame sharding and the single-bucket gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvxnet_makise_b200 import dist as mdist


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    r, w, _ = mdist.init('gloo')
    assert (r, w) == (rank, world)
    mine = mdist.shard_frames(n_frames, rank, world)
    # stand-in for the per-frame forward: a deterministic function of the global frame id
    local = [(b, b * b + 1) for b in mine]
    allres = mdist.gather_frame_results(local, n_frames, rank, world)
    # gradients: each rank contributes the sum over ITS frames of a frame-dependent gradient; the flat bucket carries
    # the frame count in its last slot, so the ragged case (5 frames = 3 + 2) divides by 5 on both ranks
    from mvxnet_makise_b200.training import FlatAdamW
    params = torch.zeros(18)
    opt = FlatAdamW(params, lr=1.0, eps=1e-12, betas=(0.0, 0.0), weight_decay=0.0)   # update = -sign-ish(g): g is what we check
    bucket = torch.zeros(19)
    for b in mine:
        bucket[:15] += float(b + 1)
        bucket[15:18] += float(2 * b)
    opt.reduce_and_step(bucket, len(mine))
    q.put((rank, mine, allres, bucket[:18].clone(), float(bucket[18])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [8, 5])
def test_frame_sharding_and_grad_allreduce_gloo(n_frames):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(b for _, mine, *_ in res for b in mine)
    assert owned == list(range(n_frames))                                  # every frame exactly once
    for rank, mine, allres, flat, nframes in res:
        assert all(mdist.owner_of(b, world) == rank for b in mine)
        assert allres == [(b, b * b + 1) for b in range(n_frames)]         # global order restored on every rank
        exp_w = sum(b + 1 for b in range(n_frames)) / n_frames
        exp_b = sum(2 * b for b in range(n_frames)) / n_frames
        assert nframes == n_frames                                         # the count travelled with the gradients
        assert torch.allclose(flat[:15], torch.full((15,), exp_w)) and torch.allclose(flat[15:], torch.full((3,), exp_b))
    assert torch.equal(res[0][3], res[1][3])                               # identical on both ranks


def test_single_process_paths():
    assert mdist.shard_frames(8, 0, 1) == list(range(8))
    assert mdist.shard_frames(8, 3, 8) == [3] and mdist.shard_frames(16, 1, 8) == [1, 9]
    assert mdist.gather_frame_results([1, 2, 3], 3, 0, 1) == [1, 2, 3]
    from mvxnet_makise_b200.training import FlatAdamW
    opt = FlatAdamW(torch.zeros(4), lr=1.0, weight_decay=0.0)
    g = torch.ones(5)
    opt.reduce_and_step(g, 2)                         # single process: the count slot still normalises the gradient
    assert torch.allclose(g[:4], torch.full((4,), 0.5)) and float(g[4]) == 2.0


def _opt_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    mdist.init('gloo')
    from mvxnet_makise_b200.training import FlatAdamW
    torch.manual_seed(0)
    params = torch.randn(1000)
    p0 = params.clone()
    opt = FlatAdamW(params)
    frames = 3 if rank == 0 else 2                          # ragged shards: 5 global frames
    for step in range(2):
        g = torch.full((1001,), float(rank + 1 + step))     # this rank's bucket (+ the frame-count slot)
        opt.reduce_and_step(g, frames)
    try:                                                    # a bucket without the slot must not guess the global count
        opt.reduce_and_step(torch.zeros(1000), frames)
        raise AssertionError('expected ValueError')
    except ValueError:
        pass
    q.put((rank, p0, params.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_adamw_allreduce_gloo():
    """The training step's exchange: ONE all-reduce of the flat gradient bucket, averaged over the global frame count,
    then AdamW on the flat parameter vector - identical parameters on every rank, equal to torch.optim.AdamW fed the
    averaged gradient."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_opt_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0][2], res[1][2])
    ref = torch.nn.Parameter(res[0][1].clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, eps=1e-6)
    for step in range(2):
        ref.grad = torch.full((1000,), ((1 + step) + (2 + step)) / 5.0)     # (rank0 + rank1 buckets) / 5 global frames (3 + 2)
        opt.step()
    assert torch.allclose(res[0][2], ref.detach(), rtol=1e-6, atol=1e-7)
