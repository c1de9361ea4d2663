"""world_size-2 gloo test of the sharding / gather logic used by bench.py --gpus N (CPU only)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infinite_video_b200.dist import gather_videos, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 1024, 1027):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_videos, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = shard_range(n_videos, rank, world)
        # stand-in for the per-video context vectors: a function of the global video id
        local = torch.stack([torch.full((4, 6), float(v)) + torch.arange(6.) for v in range(s, e)]) \
            if e > s else torch.zeros(0, 4, 6)
        full = gather_videos(local, n_videos)
        want = torch.stack([torch.full((4, 6), float(v)) + torch.arange(6.) for v in range(n_videos)])
        ok = torch.equal(full, want)
        t = max_over_ranks(float(rank + 1), "cpu")
        q.put((rank, ok, t))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [5, 8])
def test_gather_videos_world2(n_videos):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_videos
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_videos, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert all(t == 2.0 for _, _, t in res)
