"""Multi-GPU host logic on CPU: world_size-2 gloo run of the clip sharding + max-over-ranks timing reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lass_b200 import sharding


def test_shard_bounds_cover_all_clips_once():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            b = sharding.all_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _worker(rank, world, port, n_clips, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n_clips, rank, world)
    # "separate" the local slab: identity * (clip index + 1) so the gather order is checkable
    clips = torch.arange(n_clips, dtype=torch.float32)[:, None, None].expand(n_clips, 1, 4)
    local = clips[lo:hi] + 1.0
    t = sharding.max_over_ranks(10.0 + rank)
    s = sharding.sum_over_ranks(float(hi - lo))
    full = sharding.gather_waveforms(local.contiguous(), n_clips)
    dist.barrier()
    if rank == 0:
        ret["t"], ret["s"], ret["full"] = t, s, full.clone()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7])
def test_two_rank_gloo_shard_reduce_gather(n_clips):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n_clips, ret), nprocs=2, join=True)
    assert ret["t"] == 11.0                      # slowest rank wins
    assert ret["s"] == float(n_clips)            # every clip processed exactly once
    expect = torch.arange(n_clips, dtype=torch.float32)[:, None, None].expand(n_clips, 1, 4) + 1.0
    assert torch.equal(ret["full"], expect)
