"""Dev: flow_points bandwidth (64 dense KITTI frames) and pose_from_flow latency of one dense frame."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5
from tv5 import synth
eng = tv5.get_engine(); dev = "cuda"
Hh, Ww = synth.KITTI_HW
for Bf in (64, 16, 1):
    flow = torch.randn(Bf, 2, Hh, Ww, device=dev) * 3
    Kinv = torch.from_numpy(np.linalg.inv(synth.KITTI_K).astype(np.float32)).to(dev).repeat(Bf, 1, 1)
    for _ in range(3): eng.flow_to_points(flow, Kinv, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): eng.flow_to_points(flow, Kinv, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    n = Bf * (Hh - 20) * (Ww - 20)
    print(f"lib={os.environ.get('TV5_LIB','default')} frames={Bf}: {ms:.4f} ms = {n*40/ms*1e-6:.0f} GB/s", flush=True)
flow = torch.randn(1, 2, Hh, Ww, device=dev) * 3
for _ in range(3): eng.pose_from_flow(flow, Kinv[:1], 8, 1e-4)
torch.cuda.synchronize(); e0.record()
for _ in range(10): eng.pose_from_flow(flow, Kinv[:1], 8, 1e-4)
e1.record(); torch.cuda.synchronize()
print(f"  pose_from_flow dense frame {e0.elapsed_time(e1)/10:.4f} ms")
