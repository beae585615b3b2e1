"""Memory-safety checks that stand in for compute-sanitizer (memcheck / initcheck), which the B200
pool keeps closed (profiles/r2_compute_sanitizer_closed.txt): every kernel family is run on a
context whose workspace buffers carry guard zones and poisoned payloads (include/tv5.h:
tv5_debug_guard).  Pass = (i) no guard byte was overwritten by any kernel, (ii) every output is
bit-identical under three different poisons and identical to an unguarded context, i.e. no kernel
reads workspace that the same submission did not write, (iii) repeated runs are bit-identical
(racecheck's job: atomics-ordered slot allocation, the TMA tile reused across loop trips and the grid
barrier of irls_polish must not leak into results)."""
import numpy as np
import pytest
import torch

import tv5
from tv5 import synth

pytestmark = pytest.mark.gpu
THR = 1e-4


def _scenarios(eng):
    """One call of every kernel family; returns {name: numpy array} of everything that comes back."""
    dev = eng.device
    out = {}

    def put(name, *tensors):
        for i, t in enumerate(tensors):
            out[f"{name}.{i}"] = t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t).copy()

    B, N, IT = 4, 6000, 8
    pairs = [synth.make_pair(N - 37 * i, **synth.pair_variation(i)) for i in range(B)]   # ragged, one odd-sized
    ns = [p["x1"].shape[0] for p in pairs]
    x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).to(dev)
    x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).to(dev)
    sets = torch.from_numpy(np.stack([synth.make_sets(n, 512 * IT, 7000 + i) for i, n in enumerate(ns)])).to(dev)
    off = np.r_[0, np.cumsum(ns)]
    r = eng.compute_pose_batch(x1, x2, off, IT, THR, sets=sets, want_mask=True)
    put("batch", r.E, r.P, r.stats[:, :5], r.mask)
    eng.set_early_exit(True)
    big = [synth.make_pair(10000, **synth.pair_variation(10 + i)) for i in range(4)]      # large enough to be staged
    bx1 = torch.from_numpy(np.concatenate([p["x1"] for p in big])).to(dev)
    bx2 = torch.from_numpy(np.concatenate([p["x2"] for p in big])).to(dev)
    bsets = torch.from_numpy(np.stack([synth.make_sets(10000, 512 * IT, 7100 + i) for i in range(4)])).to(dev)
    boff = np.arange(5) * 10000
    r = eng.compute_pose_batch(bx1, bx2, boff, IT, THR, sets=bsets, want_mask=True)
    eng.set_early_exit(False)
    put("early_exit", r.E, r.P, r.stats[:, :3], r.mask)
    r = eng.compute_pose_batch(bx1, bx2, boff, IT, THR, sets=bsets, want_mask=True)
    put("full_same_batch", r.E, r.P, r.stats[:, :3], r.mask)
    a, b = x1[:2000].contiguous(), x2[:2000].contiguous()
    s0 = sets[0, :512].contiguous()
    for split in (True, False):
        eng.set_split_solver(split)
        r = eng.compute_pose(a, b, 1, THR, n_pre=100, n_full=1999, sets=s0, want_mask=True)   # two-stage route
        put(f"two_stage{int(split)}", r.E, r.P, r.stats[:5], r.mask)
        r = eng.compute_pose(a, b, 1, THR, sets=s0, with_cheirality=False)
        put(f"initialise{int(split)}", r.E, r.stats[:5])
    eng.set_split_solver(True)
    for k in range(4):                                   # reference RNG table; the 4th call replays a CUDA graph
        r = eng.compute_pose(a, b, 2, THR)
        put(f"ref_rng{k}", r.E, r.P, r.stats[:5])
    r = eng.compute_pose_batch(x1, x2, off, 2, THR)
    put("ref_rng_batch", r.E, r.P, r.stats[:, :5])
    eng.set_force_exact(True)
    r = eng.compute_pose(a, b, 1, THR, sets=s0, want_mask=True)
    eng.set_force_exact(False)
    put("force_exact", r.E, r.P, r.stats[:3], r.mask)
    s = eng.solve5(a, b, s0)
    put("solve5", s["E"], s["P"], s["n_roots"], s["n_valid"])
    El = s["E"].reshape(-1, 9)[:300].contiguous()
    put("score", eng.score(a, b, El, THR))
    El_ok = El[El.abs().sum(1) > 0]
    put("score_bounds", *eng.score_bounds(a, b, El_ok, THR))
    E0 = eng.compute_pose_batch(x1, x2, off, IT, THR, sets=sets).E
    put("optimise_batch", *eng.optimise_batch(x1, x2, off, E0, THR, 1.0, 10))
    put("optimise", eng.optimise(a, b, E0[0], THR, 1.0, 10))
    put("decompose", *eng.decompose_batch(E0).values())
    Eh, Ph, sh = eng.compute_pose_batch_host(x1.cpu().numpy(), x2.cpu().numpy(), off, IT, THR, sets=sets.cpu().numpy())
    put("host_batch", Eh, Ph, sh[:, :5])
    fl = synth.make_flow(hw=(96, 320), seed=3)
    flow = torch.from_numpy(fl["flow"])[None].to(dev)
    Kinv = torch.from_numpy(fl["Kinv"])[None].to(dev)
    P32, E32, rr = eng.pose_from_flow(flow, Kinv, 2, THR, margin=10)
    put("pose_from_flow", P32, E32, rr.stats[:, :5])
    ra = eng.compute_pose(a, b, 1, THR, sets=s0)
    rb = eng.compute_pose(a, b, 1, THR, sets=sets[0, 512:1024].contiguous())
    w = eng.winner_pick(torch.cat([eng.winner_record(ra, 0), eng.winner_record(rb, 512)]))
    put("winner", w.E, w.P, w.stats[:5])
    return out


def _assert_same(a, b, what):
    assert a.keys() == b.keys()
    for k in a:
        x, y = a[k], b[k]
        assert x.shape == y.shape and x.dtype == y.dtype, (what, k)
        assert x.tobytes() == y.tobytes(), f"{what}: {k} differs"     # bit for bit (NaN-safe)


def test_no_out_of_bounds_writes_and_no_reads_of_unwritten_workspace(engine):
    ref = _scenarios(engine)                              # the ordinary, unguarded context
    g = tv5.Engine(engine.device)
    try:
        g.debug_guard(0xFF)                               # float NaN / int -1 everywhere
        runs = [_scenarios(g)]
        for poison in (0x00, 0x5A):
            g.debug_poison(poison)
            runs.append(_scenarios(g))
        bad, n_buf = g.debug_check_guards()
        assert n_buf >= 20                                # every workspace buffer is guarded
        assert bad == 0, f"{bad} guard bytes overwritten"
        for i, r in enumerate(runs):
            _assert_same(ref, r, f"poison run {i} vs unguarded context")
    finally:
        g.close()


def test_guard_zones_do_detect_a_stray_write(engine):
    """The detector itself: one byte written just outside a payload is reported."""
    g = tv5.Engine(engine.device)
    try:
        g.debug_guard(0x00)
        sc = synth.make_pair(500, 3)
        g.compute_pose(torch.from_numpy(sc["x1"]).cuda(), torch.from_numpy(sc["x2"]).cuda(), 1, THR)
        assert g.debug_check_guards()[0] == 0
        g.debug_stray_write(back=True)
        assert g.debug_check_guards()[0] == 1
        g.debug_stray_write(back=False)
        assert g.debug_check_guards()[0] == 2
    finally:
        g.close()


def test_differential_fuzz_of_the_pose_entry_points():
    """tools/fuzz_paths.py: random shapes, budgets, thresholds and modes (early exit, graph replay, solver form,
    host-supplied vs reference minimal sets), non-finite / degenerate / huge coordinates — the float32 guard-band
    route equals the all-float64 route bit for bit, batches equal their pairs solved one by one, and the guard zones
    of the fuzzed context stay intact."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pr = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_paths.py"), "250", "7"], capture_output=True,
                        text=True, timeout=600)
    assert pr.returncode == 0, (pr.stdout + pr.stderr)[-800:]
    assert "fuzz ok" in pr.stdout
