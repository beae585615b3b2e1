"""Multi-GPU execution: one process per GPU, torch.distributed for the plumbing.

Two ways the path shards:

* pair-sharded (config 3): image pairs are independent; rank g takes a contiguous block of
  pairs.  No data-path collective; an optional all_gather returns every pair's 176-byte result
  to all ranks.
* hypothesis-sharded (config 4): one large pair, correspondences replicated, rank g scores the
  hypothesis ids [g*H/G, (g+1)*H/G) — i.e. reference threads [g*512/G, (g+1)*512/G) — and the
  global winner is taken from ONE all_gather of the ranks' 192-byte (key, E, P) records by a
  device-side first-maximum over the packed (count, ~id) key; no host synchronisation.  The key
  order reproduces the reference's first-maximum-over-(thread, iteration, root) rule
  (essential_matrix.cu:252).

The reference has no counterpart: it runs under torch.nn.DataParallel (main.py:219), whose
replicas serialise on the extension's blocking call.
"""
import numpy as np
import torch
import torch.distributed as dist

REF_THREADS = 512
_ID_BITS = 31  # key = count << 31 | (2^31 - 1 - (set * 16 + root))


def pair_shard(n_pairs, world, rank):
    """[start, stop) of the contiguous block of pairs owned by `rank` (balanced to +-1)."""
    base, rem = divmod(int(n_pairs), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def hypothesis_shard(iters, world, rank):
    """(first reference thread, number of threads, first hypothesis id) of `rank`'s share of the
    512*iters hypotheses.  512 must be divisible by world (1, 2, 4, 8 ... GPUs)."""
    if REF_THREADS % world:
        raise ValueError("world size must divide 512")
    tpr = REF_THREADS // world
    return rank * tpr, tpr, rank * tpr * int(iters)


def pack_key(count, set_id, root):
    """int64 ordering key: larger count wins, then smaller (set, root)."""
    ident = (int(set_id) << 4) | int(root)
    return (int(count) << _ID_BITS) | ((1 << _ID_BITS) - 1 - ident)


def unpack_key(key):
    key = int(key)
    count = key >> _ID_BITS
    ident = (1 << _ID_BITS) - 1 - (key & ((1 << _ID_BITS) - 1))
    return count, ident >> 4, ident & 15


def reduce_winner(count, set_id, root, E, P, group=None):
    """Host-level form of the winner reduction (any backend; the gloo tests of the sharding logic
    use it with CPU tensors).  All ranks call it with their local winner (global set id): ONE
    all_gather of the 23-double record (key bits, E, P), then the first maximum by key.  Returns the
    global (count, set, root, E, P).  On CUDA the engine's device-side form is used instead
    (compute_pose_hypothesis_sharded: tv5_winner_record -> all_gather -> tv5_winner_pick)."""
    dev = E.device
    has = count > 0 and set_id >= 0
    key = torch.tensor([pack_key(count, set_id, root) if has else 0], dtype=torch.int64, device=dev)
    rec = torch.cat([key.view(torch.float64), E.reshape(-1).double(), P.reshape(-1).double()])
    world = dist.get_world_size(group)
    parts = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec, group=group)
    allrec = torch.stack(parts)
    keys = allrec[:, 0].contiguous().view(torch.int64)
    owner = int(torch.argmax(keys))
    gcount, gset, groot = unpack_key(int(keys[owner]))
    if gcount == 0:
        return 0, -1, -1, torch.zeros_like(E), torch.zeros_like(P)
    n_e = E.numel()
    return (gcount, gset, groot, allrec[owner, 1:1 + n_e].view_as(E).to(E.dtype).clone(),
            allrec[owner, 1 + n_e:].view_as(P).to(P.dtype).clone())


def local_hypothesis_table(engine, n_points, iters, world, rank, sets=None):
    """This rank's rows of the [512*iters, 5] minimal-set table (the reference RNG table when sets is
    None), its first global set id, and the engine's `iters` for that many rows."""
    t0, tpr, h0 = hypothesis_shard(iters, world, rank)
    if (tpr * int(iters)) % REF_THREADS:
        raise ValueError("512*iters/world must be a multiple of 512 (iters divisible by world)")
    if sets is None:
        sets = engine.ref_rng_sets(n_points, iters)
    local = sets.view(REF_THREADS * int(iters), 5)[h0:h0 + tpr * int(iters)].contiguous()
    # the engine's hypothesis budget is 512 * iters_local: the id layout is kept by cutting the
    # thread dimension, not the iteration dimension
    return local, h0, tpr * int(iters) // REF_THREADS


def compute_pose_hypothesis_sharded(engine, x1, x2, iters, thr, sets=None, with_cheirality=True,
                                    group=None, local=None):
    """One large pair on all ranks of `group`; x1/x2 must already be replicated.  `sets` is the
    full [512*iters, 5] table (or None for the reference RNG table); `local` = a cached
    local_hypothesis_table(...) skips the slicing.  Everything is stream-ordered on the device —
    solve + score of this rank's hypotheses, tv5_winner_record, one all_gather of the 192-byte
    records over NCCL, tv5_winner_pick — with no host synchronisation; returns a PoseResult whose
    best_set is the GLOBAL hypothesis id."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if local is None:
        local = local_hypothesis_table(engine, x1.shape[0], iters, world, rank, sets)
    table, h0, iters_local = local
    r = engine.compute_pose(x1, x2, iters_local, thr, sets=table, with_cheirality=with_cheirality)
    rec = engine.winner_record(r, h0)
    allrec = torch.empty(world * engine.RECORD_DOUBLES, dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(allrec, rec, group=group)
    return engine.winner_pick(allrec)


_REC = 25   # doubles per pair: E 9 + P 12 + the eight int32 of tv5_result as 4 doubles (bit container)


def gather_pair_results(E, P, stats, n_pairs, group=None):
    """all_gather of pair-sharded results into [n_pairs, ...] tensors: ONE collective on a packed
    [pairs, 25] float64 record (E, P and the tv5_result bits), ragged shards padded to the largest."""
    world = dist.get_world_size(group)
    spans = [pair_shard(n_pairs, world, r) for r in range(world)]
    per = max(b - a for a, b in spans)
    dev = E.device
    n = E.shape[0]
    rec = torch.zeros((per, _REC), dtype=torch.float64, device=dev)
    rec[:n, :9] = E.reshape(n, 9)
    rec[:n, 9:21] = P.reshape(n, 12)
    rec[:n, 21:] = stats.reshape(n, 8).contiguous().view(torch.float64)
    out = torch.empty((world * per, _REC), dtype=torch.float64, device=dev)
    try:
        dist.all_gather_into_tensor(out, rec, group=group)
    except (RuntimeError, NotImplementedError):      # a backend without the flat form
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec, group=group)
        out = torch.cat(parts, 0)
    if per * world != n_pairs:                       # ragged: drop the padding rows
        keep = torch.cat([torch.arange(r * per, r * per + (b - a), device=dev) for r, (a, b) in enumerate(spans)])
        out = out.index_select(0, keep)
    Eg = out[:, :9].reshape(n_pairs, 3, 3)
    Pg = out[:, 9:21].reshape(n_pairs, 3, 4)
    sg = out[:, 21:].contiguous().view(torch.int32).reshape(n_pairs, 8)
    return Eg, Pg, sg


def shard_offsets(offsets, start, stop):
    """Offsets of a contiguous block of pairs, rebased to 0, plus the point range."""
    off = np.asarray(offsets, dtype=np.int64)
    return off[start:stop + 1] - off[start], int(off[start]), int(off[stop])
