"""Dev: one 10k x 4096 pair (and the SFMnet default 10k x 2560), plain launches, for an ncu launch list."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine(); eng.set_graphs(False)
sc = synth.make_pair(10000, 1234)
x1 = torch.from_numpy(sc["x1"]).cuda(); x2 = torch.from_numpy(sc["x2"]).cuda()
for it in (8, 5):
    for _ in range(4):
        r = eng.compute_pose(x1, x2, it, 1e-4)
    torch.cuda.synchronize()
    print(it, r.count, r.n_hypotheses)
