"""The N > 1 host path on the CPU: two gloo ranks shard a photon range exactly like the GPU ranks
of bench.py do (tiny_mc_b200/shards.py), each walks its shard, ONE integer all-reduce of the
2*SHELLS+4 tally words combines them, and the result is bit-identical to a single rank's.
The oracle's replay of the product's stream stands in for the CUDA kernel (checker only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, str(ROOT))
from tiny_mc_b200.shards import shard_range, tally_words  # noqa: E402


def test_shard_range_partitions_exactly():
    for first, n, world in ((0, 10, 3), (5, 0, 2), (1 << 40, (1 << 32) + 7, 8), (3, 2, 4)):
        pieces = [shard_range(first, n, r, world) for r in range(world)]
        assert pieces[0][0] == first and sum(c for _, c in pieces) == n
        for (lo, c), (lo2, _) in zip(pieces, pieces[1:]):
            assert lo + c == lo2
        assert max(c for _, c in pieces) - min(c for _, c in pieces) <= 1
    with pytest.raises(ValueError):
        shard_range(0, 10, 2, 2)
    assert tally_words(101) == 206


def _rank_main(rank, world, port, first, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle as orc

    dist.init_process_group("gloo", rank=rank, world_size=world)
    shells = 101
    lo, cnt = shard_range(first, n, rank, world)
    heat, heat2, events = orc.replay("default", 0x5EED, lo, cnt)
    words = torch.zeros(tally_words(shells), dtype=torch.int64)
    words[:shells] = torch.from_numpy(heat.astype(np.int64))
    words[shells:2 * shells] = torch.from_numpy(heat2.astype(np.int64))
    words[2 * shells] = events
    words[2 * shells + 1] = cnt
    dist.all_reduce(words)                      # the single collective of the path
    if rank == 0:
        np.save(out_path, words.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_bit_for_bit(orc, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    first, n = 1000, 3001                       # odd count: ragged shards
    out = tmp_path / "words.npy"
    mp.spawn(_rank_main, args=(2, port, first, n, str(out)), nprocs=2, join=True)
    words = np.load(out)
    heat, heat2, events = orc.replay("default", 0x5EED, first, n)
    assert np.array_equal(words[:101].astype(np.uint64), heat)
    assert np.array_equal(words[101:202].astype(np.uint64), heat2)
    assert words[202] == events and words[203] == n
