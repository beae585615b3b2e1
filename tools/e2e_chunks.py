"""Dev: end-to-end (host buffers) rate of the 256-pair batch for different copy/compute chunkings."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
B = 256
eng = tv5.get_engine()
pairs = [synth.make_pair(10000, **synth.pair_variation(i)) for i in range(B)]
x1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).pin_memory()
x2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).pin_memory()
sets = torch.from_numpy(np.stack([synth.make_sets(10000, 4096, 7000 + i) for i in range(B)])).pin_memory()
off = np.arange(B + 1) * 10000
def run(): return eng.compute_pose_batch_host(x1.numpy(), x2.numpy(), off, 8, 1e-4, sets=sets.numpy())
for _ in range(3): run()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(15): run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 15
print(f"TV5_HOST_CHUNKS={os.environ.get('TV5_HOST_CHUNKS','default')}: {dt*1e3:.3f} ms/step = {B/dt:.0f} pairs/s", flush=True)
