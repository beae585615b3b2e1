"""Target of the compute-sanitizer passes (memcheck / racecheck / synccheck / initcheck):
__graft_entry__.smoke() plus one call of every kernel family at small sizes —
an early-exit batch (staged scoring, prune_compact ping-pong), the two-stage float64 route
(set_winners), the reference-RNG default path, optimise_batch (cooperative grid barrier),
flow -> points -> pose, plane sweep, winner record / pick.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "deep-sfm-revisited_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import __graft_entry__ as ge
import tv5
from tv5 import synth

ge.smoke()
eng = tv5.get_engine()
dev = torch.device("cuda", 0)
THR = 1e-4
B, N, IT = 4, 10000, 8
pairs = [synth.make_pair(N, **synth.pair_variation(i)) for i in range(B)]
x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).to(dev)
x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).to(dev)
sets = torch.from_numpy(np.stack([synth.make_sets(N, 512 * IT, 7000 + i) for i in range(B)])).to(dev)
off = np.arange(B + 1) * N
r0 = eng.compute_pose_batch(x1, x2, off, IT, THR, sets=sets, want_mask=True)
eng.set_early_exit(True)
r1 = eng.compute_pose_batch(x1, x2, off, IT, THR, sets=sets, want_mask=True)
eng.set_early_exit(False)
assert torch.equal(r0.E, r1.E) and torch.equal(r0.stats[:, :3], r1.stats[:, :3]) and torch.equal(r0.mask, r1.mask)
print("early-exit batch ok", r1.count.tolist())
# two-stage (n_pre != n_full), both solver forms; reference RNG default (sets=None), single + batch
a, b = x1[:2000].contiguous(), x2[:2000].contiguous()
for split in (True, False):
    eng.set_split_solver(split)
    r2 = eng.compute_pose(a, b, 1, THR, n_pre=100, n_full=2000, sets=sets[0, :512].contiguous())
    print("two-stage", split, r2.count, r2.best_set, r2.best_root)
eng.set_split_solver(True)
for _ in range(4):                      # the fourth call replays the captured graph
    r3 = eng.compute_pose(a, b, 2, THR)
r3b = eng.compute_pose_batch(x1, x2, off, 2, THR)
print("reference RNG default", r3.count, r3b.count.tolist())
# refinement: cooperative kernel with a grid barrier per iteration
Eo, it = eng.optimise_batch(x1, x2, off, r0.E, THR, 1.0, 10)
E1 = eng.optimise(a, b, r0.E[0], THR, 1.0, 10)
print("optimise ok", it.tolist(), bool(torch.isfinite(Eo).all() and torch.isfinite(E1).all()))
# flow -> points -> pose; plane sweep
fl = synth.make_flow(hw=(96, 320), seed=3)
flow = torch.from_numpy(fl["flow"])[None].to(dev)
Kinv = torch.from_numpy(fl["Kinv"])[None].to(dev)
P32, E32, rr = eng.pose_from_flow(flow, Kinv, 2, THR, margin=10)
feat = torch.randn(1, 8, 24, 80, device=dev)
K4 = torch.from_numpy(synth.KITTI_K.astype(np.float32))[None].to(dev)
K4[:, :2] /= 4
vol = eng.plane_sweep(feat, feat.flip(3), P32, K4, torch.inverse(K4), 16, 1.0)
print("flow/sweep ok", rr.count.tolist(), float(vol.abs().sum()) > 0)
# hypothesis-sharded winner record / pick (two local records standing in for two GPUs)
ra = eng.compute_pose(a, b, 1, THR, sets=sets[0, :512].contiguous())
rb = eng.compute_pose(a, b, 1, THR, sets=sets[0, 512:1024].contiguous())
rec = torch.cat([eng.winner_record(ra, 0), eng.winner_record(rb, 512)])
w = eng.winner_pick(rec)
assert w.count == max(ra.count, rb.count)
print("winner pick ok", w.count, w.best_set)
torch.cuda.synchronize()
print("SANITIZE_TARGET_DONE")
