"""Multi-GPU sharding of independent utterances / streams.

The path has no cross-utterance dependency: weights are replicated, every rank (one process per GPU)
runs its contiguous shard of the batch through its own handles, and no collective touches the data
path.  The only optional communication is a result gather (code indices are 64 B per frame) when a
single consumer needs everything; it is done with ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

import torch


def shard_bounds(n_items: int, world: int):
    """Contiguous split; the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(n_items, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def shard_for_rank(n_items: int, rank: int, world: int):
    return shard_bounds(n_items, world)[rank]


def run_sharded(x_all: torch.Tensor, codec_fn, rank: int, world: int, group=None, gather=True):
    """Run ``codec_fn(x_shard) -> (idx (Nq,b,F), y (b,1,T'))`` on this rank's shard of ``x_all`` (B,1,T).

    With ``gather`` every rank receives the full (idx, y) in utterance order (all_gather of padded shards);
    otherwise only the local results and the shard bounds are returned."""
    lo, hi = shard_for_rank(x_all.shape[0], rank, world)
    idx, y = codec_fn(x_all[lo:hi]) if hi > lo else (None, None)
    if not gather or world == 1:
        return idx, y, (lo, hi)
    import torch.distributed as dist
    bounds = shard_bounds(x_all.shape[0], world)
    bmax = max(e - s for s, e in bounds)
    # shapes are identical across ranks except for the shard size: pad to bmax
    meta = torch.zeros(3, dtype=torch.int64, device=x_all.device)
    if idx is not None:
        meta[:] = torch.tensor([idx.shape[0], idx.shape[2], y.shape[2]])
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)
    nq, F, Ty = (int(v) for v in meta)
    idx_pad = torch.zeros(nq, bmax, F, dtype=torch.int64, device=x_all.device)
    y_pad = torch.zeros(bmax, 1, Ty, dtype=torch.float32, device=x_all.device)
    if idx is not None:
        idx_pad[:, : hi - lo] = idx
        y_pad[: hi - lo] = y
    idx_all = [torch.empty_like(idx_pad) for _ in range(world)]
    y_all = [torch.empty_like(y_pad) for _ in range(world)]
    dist.all_gather(idx_all, idx_pad, group=group)
    dist.all_gather(y_all, y_pad, group=group)
    idx_full = torch.cat([t[:, : e - s] for t, (s, e) in zip(idx_all, bounds)], dim=1)
    y_full = torch.cat([t[: e - s] for t, (s, e) in zip(y_all, bounds)], dim=0)
    return idx_full, y_full, (lo, hi)
