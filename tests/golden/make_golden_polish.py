"""Generates tests/golden/polish_ref.npz from the REFERENCE extension's host functions
(`decompose`, `decomposeUV`, `optimise`; CPU code of polish_E.cu), imported from
oracle/_ref/refext (built by oracle/build_ref.sh ext).  Runs in the build container, no GPU."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
from tv5 import synth  # noqa: E402

d = os.path.join(ROOT, "oracle", "_ref", "refext")
so = [f for f in os.listdir(d) if f.endswith(".so")][0]
spec = importlib.util.spec_from_file_location("essential_matrix", os.path.join(d, so))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(42)
out = {}
Es = []
for i in range(24):
    sc = synth.make_pair(50, seed=100 + i, rvec=rng.normal(0, 0.2, 3), t=rng.normal(0, 1, 3))
    E = sc["E_gt"] * rng.uniform(0.1, 10.0) * rng.choice([-1.0, 1.0]) + rng.normal(0, 1e-3, (3, 3))
    Es.append(E)
Es = np.array(Es)
out["E"] = Es
out["U"] = np.array([ref.decomposeUV(torch.from_numpy(E.copy()))[0].numpy() for E in Es])
out["V"] = np.array([ref.decomposeUV(torch.from_numpy(E.copy()))[1].numpy() for E in Es])
out["angles"] = np.array([ref.decompose(torch.from_numpy(E.copy())).numpy() for E in Es])
cases = []
for i, (n, noise, outl) in enumerate([(2000, 0.05, 0.2), (500, 0.3, 0.0), (10000, 0.05, 0.3), (64, 0.0, 0.0)]):
    sc = synth.make_pair(n, seed=300 + i, noise_px=noise, outlier_frac=outl)
    E0 = sc["E_gt"] / np.linalg.norm(sc["E_gt"]) + rng.normal(0, 3e-3, (3, 3))
    out[f"opt{i}_x1"], out[f"opt{i}_x2"], out[f"opt{i}_E0"], out[f"opt{i}_Egt"] = sc["x1"], sc["x2"], E0, sc["E_gt"]
    for delta, alpha, reps in ((1e-4, 1.0, 0), (1e-4, 1.0, 1), (1e-4, 1.0, 10), (1e-4, 0.0, 10), (5e-4, 0.5, 200)):
        Er = ref.optimise(torch.from_numpy(sc["x1"]), torch.from_numpy(sc["x2"]), torch.from_numpy(E0.copy()),
                          delta, alpha, reps).numpy()
        out[f"opt{i}_{delta:g}_{alpha:g}_{reps}"] = Er
path = os.path.join(ROOT, "tests", "golden", "polish_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
