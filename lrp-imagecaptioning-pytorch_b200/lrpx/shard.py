"""Request sharding for multi-GPU explanation jobs (one process per GPU, torch.distributed).

An explanation request is (image b, target word t).  Requests of one image share that image's forward state
(activations, gains, decoder states), so the unit of partitioning is the IMAGE: rank r owns a contiguous block
of images and all their words.  The data path needs no collective; ``gather_results`` is the optional final
all-gather of the per-request outputs (heat-maps / linguistic relevance) back to every rank in request order.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def image_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of images owned by ``rank``; blocks differ by at most one image."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_requests(words_per_image: Sequence[int], rank: int, world: int):
    """-> (lo, hi, req_img_local, req_t, req_global) for this rank.

    ``req_img_local`` indexes into the rank's own image block, ``req_global`` is the position of each request
    in the global (image-major, word-minor) request order."""
    lo, hi = image_range(len(words_per_image), rank, world)
    offsets = [0]
    for w in words_per_image:
        offsets.append(offsets[-1] + int(w))
    req_img, req_t, req_global = [], [], []
    for b in range(lo, hi):
        for t in range(int(words_per_image[b])):
            req_img.append(b - lo)
            req_t.append(t)
            req_global.append(offsets[b] + t)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    return lo, hi, i32(req_img), i32(req_t), torch.tensor(req_global, dtype=torch.int64)


def gather_results(local: torch.Tensor, counts: List[int]) -> torch.Tensor:
    """All-gathers per-request rows (ragged over ranks: ``counts[r]`` rows on rank r) into request order.
    Works with NCCL (CUDA tensors) and gloo (CPU tensors, used by the tests)."""
    world = dist.get_world_size()
    assert len(counts) == world and local.shape[0] == counts[dist.get_rank()]
    m = max(counts)
    pad = local.new_zeros((m,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
