"""Multi-GPU validation under torchrun (NCCL): hypothesis-sharded solve of one dense pair must
equal the single-GPU solve; pair-sharded batch gathers every pair's result; timings are the max
over ranks of device-event times."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5  # noqa: E402
from tv5 import dist as tdist  # noqa: E402
from tv5 import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = tv5.get_engine(dev)
THR = 1e-4
iters = 32                                   # 512 x 32 = 16,384 hypotheses (config 4)
sc = synth.make_pair(dense=True, seed=4)     # 453,620 correspondences
x1 = torch.from_numpy(sc["x1"]).to(dev)      # "replicated": every rank builds the same pair
x2 = torch.from_numpy(sc["x2"]).to(dev)
sets = torch.from_numpy(synth.make_sets(sc["x1"].shape[0], 512 * iters, 5)).to(dev)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        out = fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


local = tdist.local_hypothesis_table(eng, x1.shape[0], iters, world, rank, sets)
rs, ms_sharded = timed(lambda: tdist.compute_pose_hypothesis_sharded(eng, x1, x2, iters, THR, local=local), n=20, warm=3)
cnt, gset, groot, E, P = rs.count, rs.best_set, rs.best_root, rs.E, rs.P
res = {"world": world, "config4_hypothesis_sharded_ms": ms_sharded, "count": cnt, "set": gset, "root": groot}
if rank == 0:
    r, ms_single = None, None
    for _ in range(2):
        r = eng.compute_pose(x1, x2, iters, THR, sets=sets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r = eng.compute_pose(x1, x2, iters, THR, sets=sets)
    e1.record(); torch.cuda.synchronize()
    ms_single = e0.elapsed_time(e1) / 3
    same = (r.count, r.best_set, r.best_root) == (cnt, gset, groot) and torch.equal(r.E, E) and torch.equal(r.P, P)
    Pn = P.cpu().numpy()
    res.update({"config4_single_gpu_ms": ms_single, "sharded_equals_single": bool(same),
                "rot_err_deg": synth.rotation_error_deg(Pn[:, :3], sc["R"]),
                "trans_err_deg": synth.translation_error_deg(Pn[:, 3], sc["t"]),
                "n_hypotheses_single": r.n_hypotheses})
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"dist_check_n{world}.json"), "w"))
dist.barrier()
dist.destroy_process_group()
