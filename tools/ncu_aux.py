"""Short driver for ncu: the auxiliary kernels (flow_points, plane_sweep, irls_polish)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine()
dev = "cuda"
Hh, Ww = synth.KITTI_HW
flow = torch.randn(16, 2, Hh, Ww, device=dev) * 3
Kinv = torch.from_numpy(np.linalg.inv(synth.KITTI_K).astype(np.float32)).to(dev).repeat(16, 1, 1)
Cc, hq, wq, L = 32, 93, 307, 128
rf = torch.randn(1, Cc, hq, wq, device=dev); tg = torch.randn(1, Cc, hq, wq, device=dev)
Kq = synth.KITTI_K.copy(); Kq[:2] /= 4
K4 = torch.from_numpy(Kq.astype(np.float32)).to(dev)[None]; Ki4 = torch.from_numpy(np.linalg.inv(Kq).astype(np.float32)).to(dev)[None]
sc = synth.make_pair(10000, seed=1)
P = torch.from_numpy(np.concatenate([sc["R"], sc["t"][:, None]], 1)[None].astype(np.float32)).to(dev)
vol = torch.empty(1, 2 * Cc, L, hq, wq, device=dev)
x1 = torch.from_numpy(sc["x1"]).to(dev); x2 = torch.from_numpy(sc["x2"]).to(dev)
E0 = torch.from_numpy(sc["E_gt"] + 1e-3).to(dev)
for _ in range(3):
    eng.flow_to_points(flow, Kinv, 10)
    eng.plane_sweep(rf, tg, P, K4, Ki4, L, 1.0, out=vol)
    eng.optimise(x1, x2, E0, 1e-4, 1.0, 10)
torch.cuda.synchronize()
print("ok", float(vol.abs().mean()))
