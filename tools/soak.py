"""Dev: soak test — many submissions of changing shape and mode; results must not depend on history
and device memory must stop growing once the workspace has reached its high-water mark."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
rng = np.random.default_rng(0)
cases = []
for k in range(6):
    B = int(rng.choice([1, 1, 2, 5, 17, 40]))
    ns = [int(n) for n in rng.integers(200, 6000, B)]
    iters = int(rng.choice([1, 2, 4, 8]))
    pairs = [synth.make_pair(n, seed=1000 + 10 * k + i) for i, n in enumerate(ns)]
    x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).cuda()
    x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).cuda()
    sets = torch.from_numpy(np.stack([synth.make_sets(n, 512 * iters, 2000 + 10 * k + i) for i, n in enumerate(ns)])).cuda()
    cases.append((x1, x2, np.r_[0, np.cumsum(ns)], iters, sets))
ref = {}
mem = []
for rep in range(120):
    k = int(rng.integers(len(cases)))
    early = bool(rng.integers(2)); split = int(rng.integers(2)); graphs = bool(rng.integers(2))
    eng.set_early_exit(early); eng.set_split_solver(split); eng.set_graphs(graphs)
    x1, x2, off, iters, sets = cases[k]
    r = eng.compute_pose_batch(x1, x2, off, iters, 1e-4, sets=sets, want_mask=True)
    key = (r.E.cpu().numpy().tobytes(), r.P.cpu().numpy().tobytes(), r.mask.cpu().numpy().tobytes(),
           r.stats[:, :4].cpu().numpy().tobytes())
    if k in ref:
        assert ref[k] == key, f"case {k} differs at repetition {rep} (early={early}, split={split}, graphs={graphs})"
    else:
        ref[k] = key
    mem.append(torch.cuda.mem_get_info()[0])
eng.set_early_exit(False); eng.set_split_solver(1); eng.set_graphs(True)
print("free MiB series:", [int(m / 2**20) for m in mem[::4]])
print("soak ok:", len(ref), "cases x 120 submissions identical across modes; free memory last 20 reps constant:",
      len(set(mem[-20:])) == 1, f"({(mem[0] - mem[-1]) / 2**20:.0f} MiB workspace growth in total)")
