"""Multi-GPU sharding of the hot path: one process per GPU, contiguous ranges of pairs / haystacks balanced by
bytes, no data-path collective.  torch.distributed is plumbing only (gather of the u32 results, broadcast of the
search needle); on the CPU box the same logic runs under the gloo backend (tests/test_dist_cpu.py).
"""
import numpy as np


def shard_bounds(a_off, b_off, world):
    """Split n pairs into `world` contiguous ranges with ~equal sum(|a_i| + |b_i|).  Returns world+1 boundaries."""
    n = len(a_off) - 1
    if b_off is None:
        cost = (a_off[1:] - a_off[:-1]).astype(np.float64)
    else:
        cost = ((a_off[1:] - a_off[:-1]) + (b_off[1:] - b_off[:-1])).astype(np.float64)
    cost += 4.0  # per-pair result / launch work so empty strings still count
    csum = np.concatenate([[0.0], np.cumsum(cost)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r / world, side="left")))
    bounds.append(n)
    for i in range(1, len(bounds)):
        bounds[i] = min(max(bounds[i], bounds[i - 1]), n)
    return bounds


def take_shard(buf, off, lo, hi):
    """CSR slice [lo, hi): returns (bytes view, offsets rebased to 0)."""
    o = off[lo:hi + 1]
    return buf[int(o[0]):int(o[-1])], (o - o[0]).astype(np.uint64)


def _gather_u32(local, counts, dist):
    """all_gather of variable-length uint32 arrays (padded to the max count)."""
    import torch
    world = dist.get_world_size()
    m = max(counts) if counts else 0
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    pad = torch.zeros(m, dtype=torch.int32, device=dev)
    pad[:len(local)] = torch.from_numpy(local.view(np.int32)).to(dev)
    outs = [torch.zeros(m, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(outs, pad)
    return np.concatenate([o.cpu().numpy().view(np.uint32)[:c] for o, c in zip(outs, counts)])


def sharded_pairs(fn, a, a_off, b, b_off, gather=True):
    """Run fn(a_shard, a_off_shard, b_shard, b_off_shard) -> uint32[n_shard] on this rank's shard.
    With gather=True every rank returns the full result vector (all_gather); otherwise (result, lo, hi)."""
    import torch.distributed as dist
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    bounds = shard_bounds(a_off, b_off, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    sa, sao = take_shard(a, a_off, lo, hi)
    sb, sbo = take_shard(b, b_off, lo, hi)
    res = fn(sa, sao, sb, sbo) if hi > lo else np.zeros(0, np.uint32)
    if not gather or world == 1:
        return res if gather else (res, lo, hi)
    return _gather_u32(res, [bounds[r + 1] - bounds[r] for r in range(world)], dist)


def broadcast_needle(needle, src=0):
    """Every rank searches for rank `src`'s needle (a few dozen bytes over NCCL / gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(needle, np.uint8)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n = torch.tensor([len(needle) if dist.get_rank() == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.from_numpy(np.asarray(needle, np.uint8)))
    dist.broadcast(buf, src)
    return buf.cpu().numpy()


def sharded_search(fn, needle, hay, hay_off):
    """fn(needle, hay_shard, off_shard) -> (matches[m,3] u64, match_off[n_shard+1]).  Returns this rank's matches
    with their global haystack range (lo, hi); match lists stay sharded (they are sparse and rank-local)."""
    import torch.distributed as dist
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    needle = broadcast_needle(needle)
    bounds = shard_bounds(hay_off, None, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    sh, sho = take_shard(hay, hay_off, lo, hi)
    m, moff = fn(needle, sh, sho)
    return m, moff, lo, hi
