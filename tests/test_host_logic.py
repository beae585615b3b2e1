"""CPU tests of the host-side logic: synthetic scenes, sharding arithmetic, and the multi-rank
winner reduction over gloo (world_size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tv5 import dist as tdist
from tv5 import synth


def test_synthetic_pair_is_consistent():
    sc = synth.make_pair(2000, seed=5, f32_origin=False, noise_px=0.0, outlier_frac=0.25)
    x1h = np.c_[sc["x1"], np.ones(2000)]
    x2h = np.c_[sc["x2"], np.ones(2000)]
    res = np.abs(np.einsum("ni,ij,nj->n", x2h, sc["E_gt"] / np.linalg.norm(sc["E_gt"]), x1h))
    assert res[sc["inlier_gt"]].max() < 1e-12
    assert (res[~sc["inlier_gt"]] > 1e-6).mean() > 0.9
    assert abs(sc["inlier_gt"].mean() - 0.75) < 0.05
    d = synth.make_pair(dense=True, seed=1)
    assert d["x1"].shape == (370 * 1226, 2)


def test_pair_shard_covers_everything_once():
    for n in (1, 7, 256, 1000):
        for w in (1, 2, 4, 8):
            spans = [tdist.pair_shard(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_hypothesis_shard_and_key_order():
    for w in (1, 2, 4, 8):
        ids = []
        for r in range(w):
            t0, tpr, h0 = tdist.hypothesis_shard(32, w, r)
            assert h0 == t0 * 32
            ids += list(range(h0, h0 + tpr * 32))
        assert ids == list(range(512 * 32))
    k = tdist.pack_key
    assert k(10, 5, 0) > k(9, 0, 0)          # more inliers wins
    assert k(10, 4, 3) > k(10, 5, 0)         # tie: smaller set id wins
    assert k(10, 4, 1) > k(10, 4, 2)         # tie: smaller root wins
    assert tdist.unpack_key(k(7615, 3350, 2)) == (7615, 3350, 2)
    off, a, b = tdist.shard_offsets([0, 10, 25, 40], 1, 3)
    assert off.tolist() == [0, 15, 30] and (a, b) == (10, 40)


def test_local_hypothesis_table_cuts_the_thread_dimension():
    """Rank g of G owns reference threads [g*512/G, (g+1)*512/G), i.e. a contiguous range of hypothesis ids
    h = thread*iters + it: the concatenation of the ranks' tables is the full table, in id order."""
    iters = 32
    full = torch.arange(512 * iters * 5, dtype=torch.int32).view(512 * iters, 5)
    for world in (1, 2, 4, 8):
        rows = []
        for rank in range(world):
            table, h0, iters_local = tdist.local_hypothesis_table(None, 1000, iters, world, rank, full)
            assert table.shape == (512 * iters // world, 5) and table.is_contiguous()
            assert h0 == rank * 512 * iters // world and iters_local * 512 == table.shape[0]
            assert int(table[0, 0]) == h0 * 5
            rows.append(table)
        assert torch.equal(torch.cat(rows), full)
    with pytest.raises(ValueError):
        tdist.local_hypothesis_table(None, 1000, 3, 8, 0, torch.zeros(512 * 3, 5, dtype=torch.int32))   # 3 iterations do not split 8 ways
    with pytest.raises(ValueError):
        tdist.hypothesis_shard(8, 3, 0)                                                                  # 512 % 3


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 1 has the same count but a later hypothesis id -> rank 0's hypothesis must win
        E = torch.full((3, 3), float(rank + 1), dtype=torch.float64)
        P = torch.full((3, 4), float(10 * (rank + 1)), dtype=torch.float64)
        c, s, r, Eg, Pg = tdist.reduce_winner(100, 40 + 1000 * rank, 1, E, P)
        ok = (c, s, r) == (100, 40, 1) and float(Eg[0, 0]) == 1.0 and float(Pg[0, 0]) == 10.0
        # now rank 1 has more inliers
        c, s, r, Eg, Pg = tdist.reduce_winner(100 + 5 * rank, 40 + 1000 * rank, rank, E, P)
        ok = ok and (c, s, r) == (105, 1040, 1) and float(Eg[0, 0]) == 2.0
        # nobody has anything
        c, s, r, Eg, Pg = tdist.reduce_winner(0, -1, -1, E, P)
        ok = ok and (c, s) == (0, -1) and float(Eg.abs().sum()) == 0.0
        # pair-sharded gather (5 pairs over 2 ranks -> 3 + 2)
        a, b = tdist.pair_shard(5, world, rank)
        Em = torch.arange(a, b, dtype=torch.float64).view(-1, 1, 1).expand(-1, 3, 3).contiguous()
        Pm = torch.zeros(b - a, 3, 4, dtype=torch.float64)
        st = torch.arange(a, b, dtype=torch.int32).view(-1, 1).expand(-1, 8).contiguous()
        Eg, Pg, sg = tdist.gather_pair_results(Em, Pm, st, 5)
        ok = ok and Eg[:, 0, 0].tolist() == [0, 1, 2, 3, 4] and sg[:, 0].tolist() == [0, 1, 2, 3, 4]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_winner_reduction_and_gather_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
