"""Dev: how much exact branch-and-bound pruning of hypotheses would save on the bench workload."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
N = 10000
for i in range(3):
    p = synth.make_pair(N, **synth.pair_variation(i))
    x1 = torch.from_numpy(p["x1"]).cuda(); x2 = torch.from_numpy(p["x2"]).cuda()
    sets = torch.from_numpy(synth.make_sets(N, 4096, 7000 + i)).cuda()
    s = eng.solve5(x1, x2, sets, True)
    nv = s["n_valid"].cpu().numpy()
    E = torch.cat([s["E"][h, :nv[h]].reshape(-1, 9) for h in range(4096) if nv[h] > 0])
    full = eng.score(x1, x2, E, 1e-4).cpu().numpy()
    L = full.max()
    M = len(full)
    for stages in ([4096], [3072, 5120], [2048, 3072, 4096, 6144], [1024, 2048, 3072, 4096, 5120, 6144, 8192]):
        alive = np.ones(M, bool); work = 0.0; prev = 0
        for n1 in stages + [N]:
            work += alive.sum() * (n1 - prev)
            if n1 < N:
                c1 = eng.score(x1, x2, E, 1e-4, n_test=n1).cpu().numpy()
                alive &= (c1 + (N - n1) >= L)
            prev = n1
        print(f"pair {i}: M={M} L={L} stages={stages}: work fraction {work / (M * N):.3f}, survivors {alive.sum()}")
