"""Generates tests/golden/gpu_reference.npz ON THE GPU BOX from compiled copies of the reference:
  * oracle/_ref/libref_twin_cuda.so  (reference sources + thin dump entry points, nvcc sm_100a)
  * oracle/_ref/refext/              (the unmodified reference extension), in a subprocess
Usage (from the repo root, on a B200):  python tests/golden/make_golden_gpu.py
Writes gpurun_out/gpu_reference.npz; copy it to tests/golden/ and commit it.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
from tv5 import synth  # noqa: E402

dev = torch.device("cuda:0")
T = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_twin_cuda.so"))
vp = C.c_void_p
T.ref_rng_sets.argtypes = [C.c_int, C.c_int, vp]
T.ref_score.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_double, vp, vp]
T.ref_solve_sets.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
out = {}


def rng(N, iters):
    s = torch.empty(512 * iters, 5, dtype=torch.int32, device=dev)
    assert T.ref_rng_sets(N, iters, s.data_ptr()) == 0
    return s


for N, iters in ((10000, 8), (1000, 2), (2000, 2), (453620, 1)):
    out[f"rng_{N}_{iters}"] = rng(N, iters).cpu().numpy()

g = np.load(os.path.join(ROOT, "tests", "golden", "solver_ref_host.npz"))
for name in ("kitti", "sideways"):
    x1 = torch.from_numpy(g[f"{name}_x1"]).to(dev)
    x2 = torch.from_numpy(g[f"{name}_x2"]).to(dev)
    sets = torch.from_numpy(g[f"{name}_sets"]).to(dev)
    H = sets.shape[0]
    E_all = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    E_val = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    P_val = torch.zeros(H, 10, 12, dtype=torch.float64, device=dev)
    nr = torch.zeros(H, dtype=torch.int32, device=dev)
    nv = torch.zeros(H, dtype=torch.int32, device=dev)
    assert T.ref_solve_sets(x1.data_ptr(), x2.data_ptr(), x1.shape[0], sets.data_ptr(), H, E_all.data_ptr(),
                            nr.data_ptr(), E_val.data_ptr(), P_val.data_ptr(), nv.data_ptr()) == 0
    out[f"{name}_twin_E_all"] = E_all.cpu().numpy()
    out[f"{name}_twin_n_roots"] = nr.cpu().numpy()
    out[f"{name}_twin_E"] = E_val.cpu().numpy()
    out[f"{name}_twin_P"] = P_val.cpu().numpy()
    out[f"{name}_twin_n_valid"] = nv.cpu().numpy()
    idx = (torch.arange(10, device=dev)[None, :] < nv[:, None])
    E_list = E_val[idx].contiguous()
    M = E_list.shape[0]
    n = x1.shape[0]
    for thr in (1e-4, 1e-3):
        cnt = torch.zeros(M, dtype=torch.int32, device=dev)
        err = torch.zeros(M, n, dtype=torch.float64, device=dev)
        assert T.ref_score(x1.data_ptr(), x2.data_ptr(), n, E_list.data_ptr(), M, thr, cnt.data_ptr(), err.data_ptr()) == 0
        out[f"{name}_twin_counts_{thr:g}"] = cnt.cpu().numpy()
    out[f"{name}_twin_E_list"] = E_list.cpu().numpy()
    out[f"{name}_twin_err_sample"] = err[:48, ::4].cpu().numpy()  # raw Sampson values, bit pattern matters

# the unmodified extension, end to end
cases = [("kitti", 2, 1e-4), ("std10k", 8, 1e-4), ("std10k", 5, 1e-4)]
for name, iters, thr in cases:
    if name == "std10k":
        sc = synth.make_pair(10000, 1234)
        x1, x2 = sc["x1"], sc["x2"]
    else:
        x1, x2 = g[f"{name}_x1"], g[f"{name}_x2"]
    torch.save(dict(x1=torch.from_numpy(x1), x2=torch.from_numpy(x2)), "/tmp/pair.pt")
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_ext_check.py"), "/tmp/pair.pt", str(iters), str(thr)],
                        capture_output=True, text=True)
    print(name, iters, "refext rc", pr.returncode, pr.stderr[-300:])
    if pr.returncode == 0:
        d = json.loads(pr.stdout.strip().splitlines()[-1])
        out[f"refext_{name}_{iters}_E"] = np.array(d["E"]).reshape(3, 3)
        out[f"refext_{name}_{iters}_P"] = np.array(d["P"]).reshape(3, 4)
        out[f"refext_{name}_{iters}_count"] = np.array(d["count"])
        out[f"refext_{name}_{iters}_ms"] = np.array(d["ms"])
        print("   count", d["count"], "ms", d["ms"])

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "gpu_reference.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
