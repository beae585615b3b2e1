"""Dev: step time of compute_pose_batch for several batch sizes (strong-scaling shares of the 256-pair batch)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
Bmax = 256
pairs = [synth.make_pair(10000, **synth.pair_variation(i)) for i in range(Bmax)]
X1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).cuda()
X2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).cuda()
S = torch.from_numpy(np.stack([synth.make_sets(10000, 4096, 7000 + i) for i in range(Bmax)])).cuda()
for B in (256, 128, 64, 32, 16, 8, 4, 2):
    x1, x2, s = X1[:B * 10000], X2[:B * 10000], S[:B]
    off = np.arange(B + 1) * 10000
    for _ in range(3): eng.compute_pose_batch(x1, x2, off, 8, 1e-4, sets=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): eng.compute_pose_batch(x1, x2, off, 8, 1e-4, sets=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    eng.profile_enable(True)
    for _ in range(5): eng.compute_pose_batch(x1, x2, off, 8, 1e-4, sets=s)
    prof = eng.profile_read(); eng.profile_enable(False)
    st = {k: round(v[0] / max(v[1], 1), 3) for k, v in prof.items()}
    print(f"B={B:3d}: {ms:.4f} ms  ({ms / B * 256:.2f} ms per 256)  {st}", flush=True)
