"""Dev: distribution of GPU-solver vs golden differences."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import tv5, oracle
from tv5 import synth
eng = tv5.get_engine()
g = np.load(os.path.join(ROOT, "tests/golden/solver_ref_host.npz"))
gg = np.load(os.path.join(ROOT, "tests/golden/gpu_reference.npz"))
def dev(a, dt=torch.float64): return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dt)
def report(tag, mine, E_ref, P_ref, nr_ref, nv_ref, skip):
    H = len(nv_ref); ok = np.ones(H, bool); ok[:skip] = False
    nr, nv = mine["n_roots"].cpu().numpy(), mine["n_valid"].cpu().numpy()
    same = (nr == nr_ref) & (nv == nv_ref)
    E = mine["E"].cpu().numpy().reshape(H, 10, 9); P = mine["P"].cpu().numpy().reshape(H, 10, 12)
    sel = ok & same
    dE = np.abs(E - E_ref).reshape(H, -1).max(1)[sel] / (np.abs(E_ref).reshape(H, -1).max(1)[sel] + 1)
    dP = np.abs(P - P_ref).reshape(H, -1).max(1)[sel]
    print(f"{tag}: same {same[ok].mean():.4f} (nr {(nr==nr_ref)[ok].mean():.4f} nv {(nv==nv_ref)[ok].mean():.4f}) dE q50/90/99/max "
          + " ".join(f"{np.quantile(dE,q):.2e}" for q in (.5,.9,.99,1)) + " | dP " + " ".join(f"{np.quantile(dP,q):.2e}" for q in (.5,.9,.99,1))
          + f" | frac dE<1e-6 {(dE<1e-6).mean():.3f}")
    bad = np.where(ok & ~same)[0][:6]
    if len(bad): print("   mismatches", bad, "mine nr/nv", nr[bad], nv[bad], "ref", nr_ref[bad], nv_ref[bad])
for name in ("kitti", "noisefree", "sideways", "f64coords"):
    mine = eng.solve5(dev(g[f"{name}_x1"]), dev(g[f"{name}_x2"]), dev(g[f"{name}_sets"], torch.int32))
    report(name + " vs ref_host", mine, g[f"{name}_E"], g[f"{name}_P"], g[f"{name}_n_roots"], g[f"{name}_n_valid"], 2)
    if f"{name}_twin_E" in gg.files:
        report(name + " vs ref_gpu ", mine, gg[f"{name}_twin_E"], gg[f"{name}_twin_P"], gg[f"{name}_twin_n_roots"], gg[f"{name}_twin_n_valid"], 2)
        # reference-vs-reference (host compile vs GPU compile) for scale
        class W:  # wrap numpy as tensors
            pass
        ref_as_mine = dict(n_roots=torch.from_numpy(g[f"{name}_n_roots"]), n_valid=torch.from_numpy(g[f"{name}_n_valid"]),
                           E=torch.from_numpy(g[f"{name}_E"]), P=torch.from_numpy(g[f"{name}_P"]))
        report(name + " ref_host vs ref_gpu", ref_as_mine, gg[f"{name}_twin_E"], gg[f"{name}_twin_P"], gg[f"{name}_twin_n_roots"], gg[f"{name}_twin_n_valid"], 2)
sc = synth.make_pair(10000, 1234)
sets = synth.make_sets(10000, 4096, 17)
mine = eng.solve5(dev(sc["x1"]), dev(sc["x2"]), dev(sets, torch.int32))
orc = oracle.solve_sets(sc["x1"], sc["x2"], sets, True)
report("std10k vs oracle", mine, orc["E"], orc["P"], orc["n_roots"], orc["n_valid"], 0)
if oracle.ref_host_available():
    rh = oracle.ref_solve_sets(sc["x1"], sc["x2"], sets)
    report("std10k vs ref_host", mine, rh["E"], rh["P"], rh["n_roots"], rh["n_valid"], 0)
    o2 = dict(n_roots=torch.from_numpy(orc["n_roots"]), n_valid=torch.from_numpy(orc["n_valid"]), E=torch.from_numpy(orc["E"]), P=torch.from_numpy(orc["P"]))
    report("std10k oracle vs ref_host", o2, rh["E"], rh["P"], rh["n_roots"], rh["n_valid"], 0)
