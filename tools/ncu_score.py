"""Short driver for ncu: a few batched solves (8 pairs, config-2 shape)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = tv5.get_engine()
pairs = [synth.make_pair(10000, **synth.pair_variation(i)) for i in range(B)]
x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).cuda()
x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).cuda()
sets = torch.from_numpy(np.stack([synth.make_sets(10000, 4096, 7000 + i) for i in range(B)])).cuda()
off = np.arange(B + 1) * 10000
for _ in range(reps):
    r = eng.compute_pose_batch(x1, x2, off, 8, 1e-4, sets=sets)
torch.cuda.synchronize()
print("counts", r.count[:4], "M", r.n_hypotheses[:4])
