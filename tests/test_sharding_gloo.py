"""CPU suite, part 3: the N>1 plumbing with world_size-2 gloo processes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_char_alignment_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_indices(n_items, rank, world)
        rng = np.random.default_rng(0)
        table = {i: (np.sort(rng.random(1 + i % 5)), np.sort(rng.random(1 + i % 5))) for i in range(n_items)}
        merged = sharding.gather_alignments({i: table[i] for i in mine})
        assert sorted(merged) == list(range(n_items))
        for i in range(n_items):
            np.testing.assert_array_equal(merged[i][0], table[i][0])
            np.testing.assert_array_equal(merged[i][1], table[i][1])
        tot = sharding.gather_counters(len(mine), 10 * rank + 1, 7)
        assert tot == (n_items, sum(10 * r + 1 for r in range(world)), 7 * world)
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    world, n_items = 2, 11
    mp.spawn(_worker, args=(world, _free_port(), n_items, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shards_partition_the_list():
    for n in (0, 1, 7, 1680):
        for world in (1, 2, 4, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_single_process_gather_is_identity():
    got = sharding.gather_alignments({3: ([0.1, 0.2], [0.2, 0.4])})
    assert list(got) == [3] and got[3][1].tolist() == [0.2, 0.4]
    assert sharding.gather_counters(1, 2, 3) == (1, 2, 3)
