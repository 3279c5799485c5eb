"""Sequence sharding: partition policies on CPU and the N>1 result/timing gather with world_size-2 gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from busca_b200 import sharding


def test_round_robin_covers_every_sequence_once():
    parts = sharding.partition_sequences([1000] * 64, 8, "round_robin")
    assert [len(p) for p in parts] == [8] * 8
    assert sorted(i for p in parts for i in p) == list(range(64))
    assert parts[3] == [3, 11, 19, 27, 35, 43, 51, 59]


def test_longest_first_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    frames = rng.integers(200, 3000, 64).tolist()
    for world in (1, 2, 4, 8):
        parts = sharding.partition_sequences(frames, world)
        assert sorted(i for p in parts for i in p) == list(range(64))
        assert parts == sharding.partition_sequences(frames, world)
        lpt = sharding.makespan(frames, parts)
        rr = sharding.makespan(frames, sharding.partition_sequences(frames, world, "round_robin"))
        ideal = -(-sum(frames) // world)
        assert ideal <= lpt <= rr
        assert lpt <= ideal * 4 / 3 + max(frames) / 3 + 1          # Graham's LPT bound


def test_fewer_sequences_than_ranks_and_errors():
    parts = sharding.partition_sequences([5, 9], 4)
    assert sorted(i for p in parts for i in p) == [0, 1] and sum(1 for p in parts if not p) == 2
    assert sharding.partition_sequences([], 2) == [[], []]
    with pytest.raises(ValueError):
        sharding.partition_sequences([1], 0)
    with pytest.raises(ValueError):
        sharding.partition_sequences([1], 2, "nope")


def test_sequence_seeds_disjoint_and_world_independent():
    a = [s for r in range(8) for s in sharding.sequence_seeds(8, r, 8)]
    assert len(set(a)) == 64
    assert sharding.sequence_seeds(2, 1, 8) == sharding.sequence_seeds(8, 1, 8)


def test_single_process_gather_sorts():
    rows = sharding.pack_results([(1, 2, 7, 0, 0, 1, 1, .5), (0, 3, 1, 0, 0, 1, 1, .5), (0, 1, 9, 0, 0, 1, 1, .5), (0, 1, 2, 0, 0, 1, 1, .5)])
    out = sharding.gather_results(rows)
    assert out[:, :3].tolist() == [[0, 1, 2], [0, 1, 9], [0, 3, 1], [1, 2, 7]]
    assert sharding.gather_results(np.zeros((0, 8))).shape == (0, 8)


def test_mot_txt_format():
    rows = sharding.pack_results([(0, 1, 3, 10.04, 20.06, 30.0, 40.25, 0.456), (0, 1, -1, 0, 0, 1, 1, .5), (1, 1, 4, 1, 2, 3, 4, .9)])
    assert sharding.write_mot_txt(rows, 0) == "1,3,10.0,20.1,30.0,40.2,0.46,-1,-1,-1\n"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rows_for(seq_ids, n_frames):
    rows = []
    for s in seq_ids:
        rng = np.random.default_rng(1000 + s)
        for f in range(1, n_frames[s] + 1):
            for tid in rng.choice(50, size=int(rng.integers(0, 4)), replace=False):
                rows.append((s, f, int(tid), *rng.uniform(0, 1000, 4).tolist(), float(rng.uniform())))
    return sharding.pack_results(rows)


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = sharding.partition_sequences(n_frames, world)
        local = _rows_for(parts[rank], n_frames)
        merged = sharding.gather_results(local, dist)
        ms = sharding.reduce_max([10.0 + rank, 5.0 - rank], dist)
        q.put((rank, None if merged is None else merged, ms, parts[rank]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [[7, 3, 5, 2, 6], [4], []])
def test_gloo_world2_gather(n_frames):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        r, merged, ms, owned = q.get(timeout=120)
        got[r] = (merged, ms, owned)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[1][0] is None                                     # only rank 0 holds the merged table
    assert got[0][1] == got[1][1] == [11.0, 5.0]                 # max over ranks
    assert sorted(got[0][2] + got[1][2]) == list(range(len(n_frames)))
    expect = sharding.sort_results(_rows_for(range(len(n_frames)), n_frames))
    assert np.array_equal(got[0][0], expect)                     # identical to the single-process run, bit for bit
