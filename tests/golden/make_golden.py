"""Generates tests/golden/solver_ref_host.npz from the REFERENCE's own solver + cheirality
sources, host-compiled by oracle/build_ref.sh into oracle/_ref/libref_host.so.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixture pins the oracle (tests/test_oracle.py) and the CUDA solver (tests/test_gpu_solver.py)
on machines where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import oracle  # noqa: E402
from tv5 import synth  # noqa: E402


def main():
    assert oracle.ref_host_available(), "run oracle/build_ref.sh host first"
    out = {}
    cases = [("kitti", dict(n=2000, seed=11)),
             ("noisefree", dict(n=500, seed=12, noise_px=0.0, outlier_frac=0.0)),
             ("sideways", dict(n=1000, seed=13, t=(0.8, 0.05, 0.1), rvec=(0.05, -0.1, 0.02))),
             ("f64coords", dict(n=1000, seed=14, f32_origin=False))]
    for name, kw in cases:
        sc = synth.make_pair(**kw)
        sets = synth.make_sets(sc["x1"].shape[0], 192, seed=100 + kw["seed"])
        # a few degenerate sets: repeated indices (the reference samples with replacement)
        sets[0] = sets[0, 0]
        sets[1, 1] = sets[1, 0]
        ref = oracle.ref_solve_sets(sc["x1"], sc["x2"], sets)
        out[f"{name}_x1"] = sc["x1"]
        out[f"{name}_x2"] = sc["x2"]
        out[f"{name}_sets"] = sets
        out[f"{name}_R"] = sc["R"]
        out[f"{name}_t"] = sc["t"]
        for k in ("E_all", "n_roots", "E", "P", "n_valid"):
            out[f"{name}_{k}"] = ref[k]
    path = os.path.join(ROOT, "tests", "golden", "solver_ref_host.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
