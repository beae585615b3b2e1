"""Dev: single-pair latency with and without CUDA-graph replay, small and large shapes."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
for dense, iters in ((False, 8), (True, 16)):
    sc = synth.make_pair(10000, seed=4) if not dense else synth.make_pair(dense=True, seed=4)
    x1 = torch.from_numpy(sc["x1"]).cuda(); x2 = torch.from_numpy(sc["x2"]).cuda()
    sets = torch.from_numpy(synth.make_sets(x1.shape[0], 512 * iters, 5)).cuda()
    for g in (False, True):
        eng.set_graphs(g)
        for _ in range(5): eng.compute_pose(x1, x2, iters, 1e-4, sets=sets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(20): eng.compute_pose(x1, x2, iters, 1e-4, sets=sets)
        e1.record(); torch.cuda.synchronize()
        print(f"N={x1.shape[0]} iters={iters} graphs={g}: {e0.elapsed_time(e1)/20:.4f} ms/call (host {1e3*(time.perf_counter()-t0)/20:.4f})")
