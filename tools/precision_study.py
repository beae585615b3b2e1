"""CPU study behind DESIGN.md section 4.1: how stable are the inlier decisions when the inputs of the
Sampson products are rounded to tensor-core input formats?  (numpy; runs anywhere)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import oracle
from tv5 import synth

sc = synth.make_pair(10000, 1234)
sets = synth.make_sets(10000, 512, 5)
d = oracle.solve_sets(sc["x1"], sc["x2"], sets, True)
nv = d["n_valid"]
E = np.concatenate([d["E"][h, :nv[h]] for h in range(len(sets))]).reshape(-1, 9)
E = E / np.linalg.norm(E, axis=1, keepdims=True)
x1 = np.c_[sc["x1"], np.ones(10000)]; x2 = np.c_[sc["x2"], np.ones(10000)]
thr = 1e-4


def decisions(Em, a, b):
    M = Em.reshape(-1, 3, 3)
    Ex = np.einsum("mij,nj->mni", M, a)
    Etx = np.einsum("mji,nj->mni", M, b)
    num = np.einsum("mni,ni->mn", Ex, b)
    den = Ex[..., 0] ** 2 + Ex[..., 1] ** 2 + Etx[..., 0] ** 2 + Etx[..., 1] ** 2
    return num ** 2 <= thr ** 2 * den, num / np.sqrt(den)


def round_mant(a, bits):
    m, e = np.frexp(np.asarray(a, np.float64))
    return np.ldexp(np.round(m * 2 ** bits) / 2 ** bits, e)


ref, err_ref = decisions(E, x1, x2)
print(f"{E.shape[0]} hypotheses x 10000 points, exact inlier fraction {ref.mean():.4f}")
for name, bits in (("bf16 inputs (8-bit significand)", 8), ("tf32 / fp16 inputs (11-bit)", 11), ("fp32 inputs (24-bit)", 24)):
    dec, err = decisions(round_mant(E, bits), round_mant(x1, bits), round_mant(x2, bits))
    cnt = np.abs(dec.sum(1) - ref.sum(1))
    print(f"{name}: {np.mean(dec != ref):.4%} of the decisions flip; per-hypothesis count error median {np.median(cnt):.0f}, "
          f"max {cnt.max()}; Sampson distance off by up to {np.nanmax(np.abs(err - err_ref)) / thr:.2f} thresholds")
