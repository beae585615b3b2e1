"""Dev: one rank's share of configs[3] on a single GPU (453,620 correspondences x 16,384/G minimal sets)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
sc = synth.make_pair(dense=True, seed=4)
x1 = torch.from_numpy(sc["x1"]).cuda(); x2 = torch.from_numpy(sc["x2"]).cuda()
table = torch.from_numpy(synth.make_sets(sc["x1"].shape[0], 16384, 5)).cuda()
for G in (1, 2, 4, 8):
    local = table[: 16384 // G].contiguous()
    it = 32 // G
    for _ in range(3): r = eng.compute_pose(x1, x2, it, 1e-4, sets=local)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): r = eng.compute_pose(x1, x2, it, 1e-4, sets=local)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    eng.profile_enable(True)
    for _ in range(5): eng.compute_pose(x1, x2, it, 1e-4, sets=local)
    prof = eng.profile_read(); eng.profile_enable(False)
    st = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}
    print(f"G={G}: {ms:.4f} ms, hyps {r.n_hypotheses}, count {r.count}, stages {st}", flush=True)
