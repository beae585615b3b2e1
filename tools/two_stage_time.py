"""Dev: latency of the general two-stage route (n_pre != n_full) and of initialise, one 10k pair."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
sc = synth.make_pair(10000, 1234)
x1 = torch.from_numpy(sc["x1"]).cuda(); x2 = torch.from_numpy(sc["x2"]).cuda()
def t(fn, n=30):
    for _ in range(5): r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / n
for name, kw in (("computeP N,N", {}), ("n_pre 10, n_full 1000", dict(n_pre=10, n_full=1000)), ("n_pre 10, n_full 10000", dict(n_pre=10, n_full=10000)),
                 ("n_pre 1000, n_full 10000", dict(n_pre=1000, n_full=10000)), ("initialise N,N", dict(with_cheirality=False)),
                 ("initialise 10/1000", dict(with_cheirality=False, n_pre=10, n_full=1000))):
    r, ms = t(lambda: eng.compute_pose(x1, x2, 8, 1e-4, **kw))
    print(f"{name:28s} {ms:.4f} ms  count {r.count} hyps {r.n_hypotheses} fast {r.fast_path}", flush=True)
