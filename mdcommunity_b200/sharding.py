"""Graph-sharded data parallelism: the only multi-GPU structure of the rollout path.

Every graph's rollout is independent (no cross-graph term in segment sums, max degree, arg-max or LMCC), so graph i
goes to rank i mod W, each rank keeps its graphs resident and runs the device loop to completion, and ONE collective
at the end gathers fixed-width per-graph records (AUDC f64, n_steps, removal sequence, LMCC sequence).
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Round-robin: item i -> rank i % world."""
    return list(range(rank, n_items, world))


def pack_records(score, n_steps, actions, ranks, width):
    """[B, 2 + 2*width] float64 records: score, n_steps, actions (pad -1), LMCC ranks (pad 0).  Integers < 2^53 are
    exact in float64, the score is carried bit-exactly."""
    B = len(score)
    rec = torch.full((B, 2 + 2 * width), -1.0, dtype=torch.float64)
    rec[:, 0] = torch.as_tensor(np.asarray(score, dtype=np.float64))
    rec[:, 1] = torch.as_tensor(np.asarray(n_steps, dtype=np.float64))
    a = torch.as_tensor(np.asarray(actions, dtype=np.float64))
    r = torch.as_tensor(np.asarray(ranks, dtype=np.float64))
    w = min(width, a.shape[1])
    rec[:, 2:2 + w] = a[:, :w]
    rec[:, 2 + width:2 + width + w] = r[:, :w]
    return rec


def gather_records(rec, n_items, per_rank_max, device=None):
    """all_gather the per-rank record blocks and undo the round-robin sharding.  `rec` is this rank's [b, W] block
    (b <= per_rank_max rows; padded here).  Returns a [n_items, W] tensor on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = device if device is not None else rec.device
    block = torch.zeros(per_rank_max, rec.shape[1], dtype=rec.dtype, device=dev)
    block[:rec.shape[0]] = rec.to(dev)
    if world == 1:
        return block[:n_items].clone()
    out = torch.empty(world * per_rank_max, rec.shape[1], dtype=rec.dtype, device=dev)
    dist.all_gather_into_tensor(out, block)
    out = out.view(world, per_rank_max, -1)
    full = torch.empty(n_items, rec.shape[1], dtype=rec.dtype, device=dev)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        full[idx] = out[r, :len(idx)]
    return full


def unpack_records(full, width):
    full = full.cpu()
    score = full[:, 0].numpy()
    n_steps = full[:, 1].numpy().astype(np.int64)
    actions = full[:, 2:2 + width].numpy().astype(np.int64)
    ranks = full[:, 2 + width:2 + 2 * width].numpy().astype(np.int64)
    return score, n_steps, actions, ranks
