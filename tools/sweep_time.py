import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine(); dev = "cuda"
Cc, hq, wq, L = 32, 93, 307, 128
rf = torch.randn(1, Cc, hq, wq, device=dev); tg = torch.randn(1, Cc, hq, wq, device=dev)
Kq = synth.KITTI_K.copy(); Kq[:2] /= 4
K4 = torch.from_numpy(Kq.astype(np.float32)).to(dev)[None]; Ki4 = torch.from_numpy(np.linalg.inv(Kq).astype(np.float32)).to(dev)[None]
sc = synth.make_pair(10, seed=1)
P = torch.from_numpy(np.concatenate([sc["R"], sc["t"][:, None]], 1)[None].astype(np.float32)).to(dev)
vol = torch.empty(1, 2 * Cc, L, hq, wq, device=dev)
for _ in range(3): eng.plane_sweep(rf, tg, P, K4, Ki4, L, 1.0, out=vol)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.plane_sweep(rf, tg, P, K4, Ki4, L, 1.0, out=vol)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"G<={os.environ.get('TV5_SWEEP_G','4')}: {ms:.4f} ms  {(vol.numel()*4 + 2*rf.numel()*4)/ms*1e-6:.0f} GB/s")
for name, fn in (("fill_", lambda: vol.fill_(1.0)), ("zero_ (memset)", lambda: vol.zero_()), ("copy_", None)):
    if fn is None:
        src = torch.empty_like(vol)
        fn = lambda: vol.copy_(src)
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nb = vol.numel() * 4 * (2 if name == "copy_" else 1)
    print(f"{name}: {ms:.4f} ms  {nb/ms*1e-6:.0f} GB/s")
