"""Host-side data-parallel plumbing: one process per GPU, minibatch sharded across ranks, ONE
all-reduce of the flat gradient bucket per step (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  BatchNorm statistics stay per-rank (each rank runs the reference's batch-32 semantics).
The reference has no distributed code; this is the new launcher's logic."""
import torch
import torch.distributed as dist


def shard_permutation(n, world_size, rank, seed, epoch, batch_per_rank):
    """Seed-synchronised shuffling: every rank draws the SAME permutation of the dataset, takes the
    r-th slice of each global batch and drops the ragged tail (DataLoader(shuffle=True,
    drop_last=True) semantics of utils/load.py:34-35 upstream, sharded)."""
    g = torch.Generator().manual_seed(int(seed) * 100003 + int(epoch))
    perm = torch.randperm(n, generator=g)
    global_batch = batch_per_rank * world_size
    n_batches = n // global_batch
    perm = perm[: n_batches * global_batch].view(n_batches, world_size, batch_per_rank)
    return perm[:, rank, :]  # (n_batches, batch_per_rank) indices for this rank


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of the flat gradient bucket (the 1/world of the mean is folded into the
    consumer: the fused Adam's grad_scale).  Capturable into a CUDA graph on NCCL."""
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def allreduce_mean_(flat, group=None, world_size=None):
    """In-place average of a flat gradient bucket over the group (sum all-reduce, then scale)."""
    ws = world_size if world_size is not None else dist.get_world_size(group)
    if ws == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / ws)
    return flat


def broadcast_state_(tensors, src=0, group=None):
    """Rank `src`'s parameters / BatchNorm buffers to every rank (start of training)."""
    for t in tensors:
        dist.broadcast(t, src, group=group)
