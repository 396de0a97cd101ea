"""Host logic of the pair batcher, including a world_size-2 gloo run on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosnet_b200.pair_batcher import SyntheticPairBatcher, query_reference_groups, shard_range


@pytest.mark.parametrize("units,world", [(32, 1), (32, 8), (16, 8), (5, 4), (3, 8), (0, 2), (153, 8)])
def test_shard_range_partitions_exactly(units, world):
    seen = []
    for r in range(world):
        s, c = shard_range(units, world, r)
        seen.extend(range(s, s + c))
    assert seen == list(range(units))
    counts = [shard_range(units, world, r)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_sharded_batches_equal_single_process():
    full = SyntheticPairBatcher(6, (3, 4), world_size=1, rank=0)
    ref = next(full.batches(6))
    got = []
    for r in range(4):
        b = SyntheticPairBatcher(6, (3, 4), world_size=4, rank=r)
        for batch in b.batches(2):
            got.append(batch)
    ids = [i for b in got for i in b.pair_ids]
    assert ids == list(range(6))
    assert torch.equal(torch.cat([b.v_a for b in got]), ref.v_a)
    assert torch.equal(torch.cat([b.d_b for b in got]), ref.d_b)


def test_query_reference_groups_follow_the_reference_pairing():
    groups = query_reference_groups(10, 5, frames_per_sequence=7, seed=1)
    assert len(groups) == 10
    for q, refs in groups:
        assert len(refs) == 5 and len(set(refs)) == 5 and all(0 <= r < 7 for r in refs)
    # fewer frames than sample_range: every frame once (random.sample semantics of sbm_rgbd_loader.py:556-574)
    (_, refs), = query_reference_groups(1, 5, frames_per_sequence=3)
    assert sorted(refs) == [0, 1, 2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = SyntheticPairBatcher(total, (2, 3), world_size=world, rank=rank)
        ids = torch.full((total,), -1, dtype=torch.int64)
        mine = b.local_pair_ids()
        ids[: len(mine)] = torch.tensor(mine, dtype=torch.int64)
        gathered = [torch.empty_like(ids) for _ in range(world)]
        dist.all_gather(gathered, ids)
        flat = sorted(int(i) for g in gathered for i in g if i >= 0)
        assert flat == list(range(total))
        # checksum of checksums: sum over ranks of local feature sums == single-process sum
        local = torch.zeros(1, dtype=torch.float64)
        for batch in b.batches(2):
            local += batch.v_a.double().sum() + batch.d_b.double().sum()
        dist.all_reduce(local)
        ref = next(SyntheticPairBatcher(total, (2, 3)).batches(total))
        expect = ref.v_a.double().sum() + ref.d_b.double().sum()
        assert abs(float(local) - float(expect)) < 1e-6 * max(1.0, abs(float(expect)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_partition():
    mp.spawn(_worker, args=(2, _free_port(), 5), nprocs=2, join=True)
