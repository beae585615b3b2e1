"""Runs the unmodified reference extension (oracle/_ref/refext) on a saved pair; prints JSON."""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ext_dir = os.path.join(ROOT, "oracle", "_ref", "refext")
so = [f for f in os.listdir(ext_dir) if f.endswith(".so")][0]
spec = importlib.util.spec_from_file_location("essential_matrix", os.path.join(ext_dir, so))
refext = importlib.util.module_from_spec(spec)
spec.loader.exec_module(refext)
d = torch.load(sys.argv[1])
iters = int(sys.argv[2]); thr = float(sys.argv[3])
x1 = d["x1"].cuda(); x2 = d["x2"].cuda()
N = x1.shape[0]
print("calling reference computeP", flush=True)
E, P, c = refext.computeP(x1, x2, N, N, iters, thr)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.time()
    refext.computeP(x1, x2, N, N, iters, thr)
    torch.cuda.synchronize(); ts.append(time.time() - t0)
print(json.dumps(dict(count=int(c), E=E.cpu().numpy().ravel().tolist(), P=P.cpu().numpy().ravel().tolist(),
                      ms=1e3 * float(np.median(ts)))))
