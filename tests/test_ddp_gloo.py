"""Data-parallel host logic with world_size 2 on the gloo backend (CPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_surrogate_b200 import ddp
    idx = ddp.shard_permutation(100, world, rank, seed=1, epoch=3, batch_per_rank=8)
    assert idx.shape == (6, 8)
    gathered = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(gathered, idx)
    allidx = torch.cat([g.reshape(-1) for g in gathered])
    assert allidx.unique().numel() == allidx.numel() == 96      # disjoint shards, ragged tail dropped
    g = torch.full((1000,), float(rank + 1))
    ddp.allreduce_mean_(g)
    assert torch.allclose(g, torch.full((1000,), 1.5))
    p = torch.full((10,), float(rank))
    ddp.broadcast_state_([p], src=0)
    assert float(p.abs().sum()) == 0.0
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").close()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
