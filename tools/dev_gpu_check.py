"""Development smoke/measurement script run on the GPU box (not part of the test suite)."""
import ctypes as C
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-sfm-revisited_b200"))
import oracle  # noqa: E402
import tv5  # noqa: E402
from tv5 import synth  # noqa: E402

out = {}
dev = torch.device("cuda:0")
eng = tv5.get_engine(dev)
print("SMs", eng.sm_count, torch.cuda.get_device_name(0))

# 1. FP32 peak
for mode in (0, 1):
    out[f"fp32_peak_mode{mode}_tflops"] = eng.measure_fp32_peak(mode)
print({k: v for k, v in out.items()})

sc = synth.make_pair(10000, 1234)
x1 = torch.from_numpy(sc["x1"]).to(dev)
x2 = torch.from_numpy(sc["x2"]).to(dev)
thr = 1e-4
iters = 8
H = 512 * iters

# 2. reference twin
twin_path = os.path.join(ROOT, "oracle", "_ref", "libref_twin_cuda.so")
sets_ref = None
if os.path.exists(twin_path):
    T = C.CDLL(twin_path)
    vp = C.c_void_p
    T.ref_rng_sets.argtypes = [C.c_int, C.c_int, vp]
    T.ref_score.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_double, vp, vp]
    T.ref_solve_sets.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
    sets_ref = torch.empty(H, 5, dtype=torch.int32, device=dev)
    rc = T.ref_rng_sets(10000, iters, sets_ref.data_ptr())
    mine = eng.ref_rng_sets(10000, iters)
    torch.cuda.synchronize()
    print("twin rng rc", rc, "equal to tv5_ref_rng_sets:", bool((mine == sets_ref).all().item()),
          "max idx", int(sets_ref.max()))
    out["rng_equal"] = bool((mine == sets_ref).all().item())
    # twin solve on GPU vs oracle (host)
    E_all = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    E_val = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    P_val = torch.zeros(H, 10, 12, dtype=torch.float64, device=dev)
    nr = torch.zeros(H, dtype=torch.int32, device=dev)
    nv = torch.zeros(H, dtype=torch.int32, device=dev)
    t0 = time.time()
    rc = T.ref_solve_sets(x1.data_ptr(), x2.data_ptr(), 10000, sets_ref.data_ptr(), H, E_all.data_ptr(),
                          nr.data_ptr(), E_val.data_ptr(), P_val.data_ptr(), nv.data_ptr())
    print("twin solve rc", rc, "time %.3fs" % (time.time() - t0))
    # twin scoring vs oracle bits
    Ms = int(nv.sum().item())
    idx = (torch.arange(10, device=dev)[None, :] < nv[:, None])
    E_list = E_val[idx].contiguous()
    cnt_twin = torch.zeros(Ms, dtype=torch.int32, device=dev)
    sub = min(Ms, 256)
    err_twin = torch.zeros(sub, 10000, dtype=torch.float64, device=dev)
    rc = T.ref_score(x1.data_ptr(), x2.data_ptr(), 10000, E_list.data_ptr(), sub, thr, cnt_twin.data_ptr(),
                     err_twin.data_ptr())
    El = E_list[:sub].cpu().numpy()
    err_or = np.empty((sub, 10000))
    for m in range(sub):
        for k in range(0, 10000, 97):
            err_or[m, k] = oracle.sampson_err(El[m], sc["x1"][k, 0], sc["x1"][k, 1], sc["x2"][k, 0], sc["x2"][k, 1])
    et = err_twin.cpu().numpy()[:, ::97]
    eo = err_or[:, ::97]
    print("twin ComputeError vs oracle: bit-equal fraction", float((et == eo).mean()),
          "max rel diff", float(np.nanmax(np.abs(et - eo) / np.maximum(np.abs(eo), 1e-300))))
    out["sampson_bits_equal_fraction"] = float((et == eo).mean())
    t0 = time.time()
    rc = T.ref_score(x1.data_ptr(), x2.data_ptr(), 10000, E_list.data_ptr(), Ms, thr, cnt_twin.data_ptr(), None)
    print("twin score all M=%d rc %d time %.3fs" % (Ms, rc, time.time() - t0))
    cnt_mine = eng.score(x1, x2, E_list, thr)
    print("tv5_score == twin counts:", bool((cnt_mine == cnt_twin).all().item()))
    out["score_equal_twin"] = bool((cnt_mine == cnt_twin).all().item())
    # my solver vs twin solver
    ms = eng.solve5(x1, x2, sets_ref, with_cheirality=True)
    torch.cuda.synchronize()
    eq_nr = (ms["n_roots"] == nr).float().mean().item()
    eq_nv = (ms["n_valid"] == nv).float().mean().item()
    ok = (ms["n_valid"] == nv) & (ms["n_roots"] == nr)
    dE = (ms["E"].view(H, 10, 9) - E_val).abs().amax(dim=(1, 2))[ok]
    dP = (ms["P"].view(H, 10, 12) - P_val).abs().amax(dim=(1, 2))[ok]
    print("solver vs twin: n_roots eq %.4f n_valid eq %.4f; dE median %.3g p99 %.3g max %.3g; dP median %.3g max %.3g" % (
        eq_nr, eq_nv, dE.median().item(), dE.quantile(0.99).item(), dE.max().item(), dP.median().item(), dP.max().item()))
    out["solver_vs_twin"] = dict(n_roots_eq=eq_nr, n_valid_eq=eq_nv, dE_median=dE.median().item(), dE_max=dE.max().item())
else:
    print("no twin")

sets = sets_ref if sets_ref is not None else torch.from_numpy(synth.make_sets(10000, H, 7)).to(dev)
sets_h = sets.cpu().numpy()

# 3. solver vs oracle
ms = eng.solve5(x1, x2, sets, with_cheirality=True)
torch.cuda.synchronize()
orc = oracle.solve_sets(sc["x1"], sc["x2"], sets_h, True)
nv_m = ms["n_valid"].cpu().numpy()
nr_m = ms["n_roots"].cpu().numpy()
print("solver vs oracle: n_roots eq %.4f n_valid eq %.4f" % ((nr_m == orc["n_roots"]).mean(), (nv_m == orc["n_valid"]).mean()))
ok = (nv_m == orc["n_valid"]) & (nr_m == orc["n_roots"])
dE = np.abs(ms["E"].cpu().numpy().reshape(H, 10, 9) - orc["E"]).reshape(H, -1).max(1)[ok]
print("  dE median %.3g p99 %.3g max %.3g" % (np.median(dE), np.quantile(dE, .99), dE.max()))
bad = np.where(~ok)[0][:5]
print("  mismatching sets", bad, nr_m[bad], orc["n_roots"][bad], nv_m[bad], orc["n_valid"][bad])

# 4. exact scoring vs oracle
idx = (torch.arange(10, device=dev)[None, :] < ms["n_valid"][:, None])
E_list = ms["E"].view(H, 10, 9)[idx].contiguous()
M = E_list.shape[0]
cnt, masks = eng.score(x1, x2, E_list, thr, want_mask=True)
torch.cuda.synchronize()
sub = 300
c_or, m_or = oracle.score(sc["x1"], sc["x2"], E_list[:sub].cpu().numpy(), thr, want_mask=True)
print("exact score vs oracle counts equal:", bool((cnt[:sub].cpu().numpy() == c_or).all()))
mk = masks[:sub].cpu().numpy().view(np.uint32)
bits = ((mk[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(sub, -1)[:, :10000].astype(np.uint8)
print("exact masks vs oracle equal:", bool((bits == m_or).all()))
out["score_equal_oracle"] = bool((cnt[:sub].cpu().numpy() == c_or).all()) and bool((bits == m_or).all())

# 5. bounds
lo, hi = eng.score_bounds(x1, x2, E_list, thr)
torch.cuda.synchronize()
viol = int(((lo > cnt) | (hi < cnt)).sum().item())
w = (hi - lo).float()
top = cnt.argmax()
print("bounds: M=%d violations=%d  width mean %.2f max %d ; best count %d lo %d hi %d" % (
    M, viol, w.mean().item(), int(w.max().item()), int(cnt[top]), int(lo[top]), int(hi[top])))
ncand = int((hi >= lo.max()).sum().item())
print("  candidates (hi >= max lo):", ncand)
out["bounds"] = dict(M=M, violations=viol, width_mean=w.mean().item(), n_candidates=ncand)

# 6. full pipeline vs oracle ransac
r = eng.compute_pose(x1, x2, iters, thr, sets=sets, want_mask=True)
torch.cuda.synchronize()
t0 = time.time()
o = oracle.ransac(sc["x1"], sc["x2"], sets_h, iters, thr, want_mask=True)
t_or = time.time() - t0
print("pipeline: count %d set %d root %d M %d cand %d fast %d | oracle count %d set %d root %d (%.1fs)" % (
    r.count, r.best_set, r.best_root, r.n_hypotheses, r.n_candidates, r.fast_path, o["count"], o["best_set"], o["best_root"], t_or))
E = r.E.cpu().numpy(); P = r.P.cpu().numpy()
print("  dE %.3g dP %.3g mask equal %s ; rot err %.4f deg, t err %.4f deg" % (
    np.abs(E - o["E"]).max(), np.abs(P - o["P"]).max(), bool((r.mask.cpu().numpy() == o["mask"]).all()),
    synth.rotation_error_deg(P[:, :3], sc["R"]), synth.translation_error_deg(P[:, 3], sc["t"])))
out["pipeline"] = dict(count=r.count, oracle_count=o["count"], set=r.best_set, oracle_set=o["best_set"])

# 7. reference extension (in a subprocess: its error path calls exit())
import subprocess
torch.save(dict(x1=x1.cpu(), x2=x2.cpu()), "/tmp/pair.pt")
pr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_ext_check.py"), "/tmp/pair.pt", str(iters), str(thr)],
                    capture_output=True, text=True)
print("ref ext subprocess rc", pr.returncode)
print(pr.stdout[-3000:]); print(pr.stderr[-2000:])
r2 = eng.compute_pose(x1, x2, iters, thr)  # sets=None -> reference RNG table
print("tv5 (ref RNG table): count %d set %d root %d" % (r2.count, r2.best_set, r2.best_root))
print("  E", r2.E.cpu().numpy().ravel())
if sets_ref is not None:
    h = r2.best_set
    print("  twin E for that set", E_val[h, r2.best_root].cpu().numpy().ravel(), "twin nv", int(nv[h]))

# 8. timings
def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_single = timeit(lambda: eng.compute_pose(x1, x2, iters, thr, sets=sets))
print("single pair (10k x 4096): %.3f ms/pair" % t_single)
out["single_pair_ms"] = t_single
eng.profile_enable(True)
for _ in range(20):
    eng.compute_pose(x1, x2, iters, thr, sets=sets)
prof = eng.profile_read()
eng.profile_enable(False)
print("  stages (ms):", {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items()})
out["single_stages_ms"] = {k: v[0] / max(v[1], 1) for k, v in prof.items()}

B = 64
pairs = [synth.make_pair(10000, **synth.pair_variation(i)) for i in range(B)]
X1 = torch.from_numpy(np.concatenate([p["x1"] for p in pairs])).to(dev)
X2 = torch.from_numpy(np.concatenate([p["x2"] for p in pairs])).to(dev)
off = np.arange(B + 1) * 10000
S = torch.from_numpy(np.stack([synth.make_sets(10000, H, 100 + i) for i in range(B)])).to(dev)
rb = eng.compute_pose_batch(X1, X2, off, iters, thr, sets=S)
torch.cuda.synchronize()
print("batch counts[:8]", rb.count[:8], "cands", rb.n_candidates[:8], "M", rb.n_hypotheses[:4])
t_batch = timeit(lambda: eng.compute_pose_batch(X1, X2, off, iters, thr, sets=S), n=10, warm=2)
print("batch of %d: %.3f ms/batch = %.1f pairs/s" % (B, t_batch, B / t_batch * 1e3))
out["batch64_ms"] = t_batch
eng.profile_enable(True)
for _ in range(5):
    eng.compute_pose_batch(X1, X2, off, iters, thr, sets=S)
prof = eng.profile_read()
eng.profile_enable(False)
st = {k: v[0] / max(v[1], 1) for k, v in prof.items()}
print("  stages (ms):", {k: round(v, 4) for k, v in st.items()})
evals = float(rb.n_hypotheses.sum()) * 10000
print("  score kernel: %.3g evals in %.3f ms = %.3g evals/s = %.1f TFLOP/s (34 flop/eval)" % (
    evals, st["score_bounds"], evals / st["score_bounds"] * 1e3, evals / st["score_bounds"] * 1e3 * 34e-12))
out["batch_stages_ms"] = st
out["batch_score_tflops"] = evals / st["score_bounds"] * 1e3 * 34e-12
# accuracy of the batch
errs = []
for i in range(B):
    P = rb.P[i].cpu().numpy()
    errs.append((synth.rotation_error_deg(P[:, :3], pairs[i]["R"]), synth.translation_error_deg(P[:, 3], pairs[i]["t"])))
errs = np.array(errs)
print("  batch pose errors: rot median %.4f max %.4f deg; t median %.4f max %.4f deg" % (
    np.median(errs[:, 0]), errs[:, 0].max(), np.median(errs[:, 1]), errs[:, 1].max()))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "dev_check.json"), "w"), indent=1, default=float)
