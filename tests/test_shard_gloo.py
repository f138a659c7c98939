"""CPU, world_size 2 over gloo: the request sharder covers every request exactly once, in order, and the final
gather reassembles ragged per-rank results (the N>1 host logic of bench.py / lrpx.shard)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lrpx import shard


def test_image_ranges_partition():
    for n in (0, 1, 5, 64, 513):
        for world in (1, 2, 3, 8):
            blocks = [shard.image_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_requests_cover_all():
    words = [3, 1, 4, 2, 5]
    seen = []
    for r in range(2):
        lo, hi, ri, rt, rg = shard.shard_requests(words, r, 2)
        assert int(ri.max()) == hi - lo - 1
        seen += rg.tolist()
    assert seen == list(range(sum(words)))


def _worker(rank, world, port, words):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi, ri, rt, rg = shard.shard_requests(words, rank, world)
        # the "explanation" of request g is a row filled with g; word index goes to column 0
        local = rg.float()[:, None].repeat(1, 4)
        local[:, 0] = rt.float()
        counts = [len(shard.shard_requests(words, r, world)[4]) for r in range(world)]
        full = shard.gather_results(local, counts)
        assert full.shape == (sum(words), 4)
        assert torch.equal(full[:, 1], torch.arange(sum(words)).float())
        exp_t = torch.tensor([t for w in words for t in range(w)]).float()
        assert torch.equal(full[:, 0], exp_t)
    finally:
        dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, [3, 1, 4, 2, 5]), nprocs=2, join=True)
