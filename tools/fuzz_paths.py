"""Dev: differential fuzzing of the pose entry points.  For random shapes / budgets / thresholds / modes the float32
guard-band route must equal the all-float64 route bit for bit, a batch must equal its pairs solved one by one, early exit
and CUDA-graph replay must not change anything, and a guarded context must keep its guard zones intact.
    python tools/fuzz_paths.py [n_cases] [seed]
"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
eng = tv5.Engine(torch.device("cuda", 0))
eng.debug_guard(0xFF)
ref = tv5.get_engine()            # ordinary context: the float64 route
dev = "cuda"


def same(a, b, what, case):
    ok = torch.equal(a.E, b.E) and torch.equal(a.P, b.P) and torch.equal(a.stats[..., :3], b.stats[..., :3])
    if a.mask is not None and b.mask is not None:
        ok = ok and torch.equal(a.mask, b.mask)
    if not ok:
        print("MISMATCH", what, case, a.stats.tolist(), b.stats.tolist(), flush=True)
        sys.exit(1)


t0 = time.time()
for case in range(n_cases):
    B = int(rng.choice([1, 1, 1, 2, 3, 7]))
    ns = [int(rng.choice([rng.integers(5, 40), rng.integers(40, 1500), rng.integers(1500, 12000)])) for _ in range(B)]
    iters = int(rng.choice([1, 2, 3, 5, 8]))
    thr = float(10 ** rng.uniform(-5, -2.5))
    cheir = bool(rng.integers(2))
    use_sets = bool(rng.integers(2))
    noise = float(rng.choice([0.0, 0.05, 0.5]))
    pairs = [synth.make_pair(n, seed=int(rng.integers(1 << 30)), noise_px=noise, outlier_frac=float(rng.uniform(0, 0.6))) for n in ns]
    a1 = np.concatenate([p["x1"] for p in pairs]); a2 = np.concatenate([p["x2"] for p in pairs])
    quirk = rng.uniform()
    if quirk < 0.08:                                   # a non-finite coordinate: that pair takes the float64 route
        a1[int(rng.integers(a1.shape[0])), int(rng.integers(2))] = float(rng.choice([np.nan, np.inf, -np.inf]))
    elif quirk < 0.14:                                 # all correspondences identical (every minimal set degenerate)
        a1[:] = a1[0]; a2[:] = a2[0]
    elif quirk < 0.20:                                 # huge coordinates (beyond the float32 band's range)
        a1 *= 5000.0
    x1 = torch.from_numpy(a1).to(dev)
    x2 = torch.from_numpy(a2).to(dev)
    off = np.r_[0, np.cumsum(ns)]
    sets = torch.from_numpy(np.stack([synth.make_sets(n, 512 * iters, int(rng.integers(1 << 30))) for n in ns])).to(dev) if use_sets else None
    desc = dict(case=case, ns=ns, iters=iters, thr=thr, cheir=cheir, sets=use_sets)
    eng.set_early_exit(bool(rng.integers(2))); eng.set_graphs(bool(rng.integers(2))); eng.set_split_solver(bool(rng.integers(4) > 0))
    fast = eng.compute_pose_batch(x1, x2, off, iters, thr, sets=sets, with_cheirality=cheir, want_mask=True)
    ref.set_force_exact(True)
    exact = ref.compute_pose_batch(x1, x2, off, iters, thr, sets=sets, with_cheirality=cheir, want_mask=True)
    ref.set_force_exact(False)
    same(fast, exact, "fast vs float64", desc)
    for rep in range(2 if B == 1 else 1):          # singles (repeat: graph capture / replay for B = 1)
        for i in range(B):
            a, b = off[i], off[i + 1]
            one = eng.compute_pose(x1[a:b].contiguous(), x2[a:b].contiguous(), iters, thr, sets=None if sets is None else sets[i].contiguous(),
                                   with_cheirality=cheir, want_mask=True)
            ok = (torch.equal(one.E, fast.E[i]) and torch.equal(one.P, fast.P[i]) and torch.equal(one.stats[:3], fast.stats[i, :3])
                  and torch.equal(one.mask, fast.mask[a:b]))
            if not ok:
                print("MISMATCH single vs batch", desc, i, one.stats.tolist(), fast.stats[i].tolist(), flush=True); sys.exit(1)
    if case % 25 == 24:
        eng.debug_poison(int(rng.choice([0x00, 0xFF, 0x5A])))
        bad, nb = eng.debug_check_guards()
        if bad:
            print("GUARD BYTES OVERWRITTEN", bad, desc, flush=True); sys.exit(1)
bad, nb = eng.debug_check_guards()
print(f"fuzz ok: {n_cases} cases (seed {seed}) in {time.time() - t0:.1f} s, guard bytes overwritten: {bad} over {nb} buffers")
sys.exit(1 if bad else 0)
