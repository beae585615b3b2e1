"""Dev: quick batch timing + stage profile (B pairs of config-2 shape)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = tv5.get_engine()
eng.set_overlap(os.environ.get('TV5_OVERLAP', '0') != '0')
eng.set_early_exit(os.environ.get('TV5_EARLY', '0') == '1')
pairs = [synth.make_pair(10000, **synth.pair_variation(i)) for i in range(B)]
x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).cuda()
x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).cuda()
sets = torch.from_numpy(np.stack([synth.make_sets(10000, 4096, 7000 + i) for i in range(B)])).cuda()
off = np.arange(B + 1) * 10000
def run(): return eng.compute_pose_batch(x1, x2, off, 8, 1e-4, sets=sets)
for _ in range(3): r = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): r = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
eng.profile_enable(True)
for _ in range(5): run()
prof = eng.profile_read(); eng.profile_enable(False)
st = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()}
evals = float(r.n_hypotheses.sum()) * 10000
print(f"lib={os.environ.get('TV5_LIB','default')} B={B}: {ms:.3f} ms/batch = {B/ms*1e3:.0f} pairs/s; stages {st}; score {evals/st['score_bounds']*1e3*34e-12:.1f} TFLOP/s; cands {r.n_candidates[:6]} counts {r.count[:4]}")
a, b, s0 = x1[:10000].contiguous(), x2[:10000].contiguous(), sets[0].contiguous()
for _ in range(5): eng.compute_pose(a, b, 8, 1e-4, sets=s0)
torch.cuda.synchronize(); e0.record()
for _ in range(50): eng.compute_pose(a, b, 8, 1e-4, sets=s0)
e1.record(); torch.cuda.synchronize()
print(f"  single pair {e0.elapsed_time(e1)/50:.4f} ms")
