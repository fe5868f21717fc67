"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed).

The path shards without any data-path collective: images are independent, and the
points of one image are independent of each other.  The reference has no multi-GPU inference at all
(`device='cuda:0'` defaults, recon/recon_fit_base.py:50; sequences are split by hand with -fs/-fe,
recon/recon_fit_behave.py:385-386).  Here:

  * `shard_images`   batch axis across ranks (config 4: 32 images -> 4 per GPU),
  * `shard_range`    contiguous point shards of one huge query (config 2: the 256^3 grid),
  * `gather_results` the only collective: an all-gather of per-image / per-shard results
                     (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                align: int = 128) -> Tuple[int, int]:
    """Contiguous [start, start+count) share of `total` points for `rank`; shard boundaries are multiples
    of `align` (the query kernel's tile size) so no tile straddles two ranks.  Ragged totals are fine:
    the last ranks may get fewer points, or none."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    tiles = (total + align - 1) // align
    per, extra = divmod(tiles, world_size)
    t0 = rank * per + min(rank, extra)
    t1 = t0 + per + (1 if rank < extra else 0)
    start, end = min(t0 * align, total), min(t1 * align, total)
    return start, end - start


def shard_images(batch: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Indices of the images of a batch that `rank` processes (contiguous blocks, like DistributedSampler
    without shuffling)."""
    start, count = shard_range(batch, rank, world_size, align=1)
    return list(range(start, start + count))


def gather_results(local: torch.Tensor, counts: Optional[Sequence[int]] = None, dim: int = -1) -> torch.Tensor:
    """All-gather per-rank results along `dim`.  `counts[r]` is rank r's extent along `dim` (ragged shards
    are padded to the maximum for the collective and trimmed afterwards).  Identity when not distributed."""
    rank, w = world()
    if w == 1:
        return local
    dim = dim % local.dim()
    if counts is None:
        n = torch.tensor([local.shape[dim]], device=local.device)
        all_n = [torch.zeros_like(n) for _ in range(w)]
        dist.all_gather(all_n, n)
        counts = [int(x.item()) for x in all_n]
    m = max(counts)
    pad_shape = list(local.shape)
    pad_shape[dim] = m
    buf = local.new_zeros(pad_shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    out = [torch.empty_like(buf) for _ in range(w)]
    dist.all_gather(out, buf.contiguous())
    return torch.cat([o.narrow(dim, 0, c) for o, c in zip(out, counts)], dim)


def query_grid_sharded(net, res, b_min, b_max, crop_center, batch_index: int = 0, head_mask: int = 1,
                       chunk: int = 1 << 22):
    """The dense grid of model/sdf.py:4-48 for ONE image, point-sharded over the ranks: every rank holds the
    feature maps (encode redundantly: cheaper than a broadcast), evaluates its contiguous slab and the slabs are
    all-gathered.  Returns per-head (n_out, X*Y*Z) tensors on every rank."""
    from . import _lib
    total = int(res[0]) * int(res[1]) * int(res[2])
    start, count = shard_range(total)
    feat, skip = net._maps()
    cc = crop_center.detach().to(feat.device, torch.float32).contiguous()
    outs = [torch.empty(c, total, device=feat.device) if head_mask & (1 << i) else None
            for i, c in enumerate(_lib.HEAD_OUT)]
    for s in range(start, start + count, chunk):
        net.handle.query_grid(feat, skip, cc, batch_index, res, b_min, b_max, s, min(chunk, start + count - s),
                              head_mask, outs)
    rank, w = world()
    if w == 1:
        return outs
    counts = [shard_range(total, r, w)[1] for r in range(w)]
    return [None if o is None else gather_results(o[:, start:start + count].contiguous(), counts, dim=1) for o in outs]
