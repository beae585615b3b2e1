"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of
libtv5.so (via tv5.Engine) and is checked against the CPU oracle and the committed golden
vectors.  Tolerances are written next to each assertion."""
import os

import numpy as np
import pytest
import torch

import oracle
from tv5 import synth

pytestmark = pytest.mark.gpu
THR = 1e-4


def dev(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def unpack_masks(masks, n):
    mk = masks.cpu().numpy().view(np.uint32)
    bits = (mk[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
    return bits.reshape(mk.shape[0], -1)[:, :n].astype(np.uint8)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "solver_ref_host.npz"))


@pytest.fixture(scope="module")
def gold_gpu(golden_dir):
    p = os.path.join(golden_dir, "gpu_reference.npz")
    if not os.path.exists(p):
        pytest.skip("gpu_reference.npz not generated yet")
    return np.load(p)


@pytest.fixture(scope="module")
def std_pair():
    sc = synth.make_pair(10000, 1234)
    return sc, dev(sc["x1"]), dev(sc["x2"])


# ---------------------------------------------------------------------------------------------
# scoring: bit-exact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 1999])
def test_exact_score_bit_exact_vs_oracle(engine, gold, n):
    x1, x2 = gold["kitti_x1"][:n], gold["kitti_x2"][:n]
    E = gold["kitti_E"].reshape(-1, 9)
    E = E[np.abs(E).sum(1) > 0][:200]
    for thr in (1e-4, 1e-3, 0.5):
        cnt, masks = engine.score(dev(x1), dev(x2), dev(E), thr, want_mask=True)
        c_or, m_or = oracle.score(x1, x2, E, thr, want_mask=True)
        assert (cnt.cpu().numpy() == c_or).all()          # integer: exact
        assert (unpack_masks(masks, n) == m_or).all()     # bit: exact


def test_exact_score_matches_reference_gpu_counts(engine, gold, gold_gpu):
    """Counts of the reference's own ComputeError<double> compiled by nvcc (twin, GPU box)."""
    for name in ("kitti", "sideways"):
        x1, x2 = dev(gold[f"{name}_x1"]), dev(gold[f"{name}_x2"])
        E = dev(gold_gpu[f"{name}_twin_E_list"])
        for thr in (1e-4, 1e-3):
            cnt = engine.score(x1, x2, E, thr).cpu().numpy()
            assert (cnt == gold_gpu[f"{name}_twin_counts_{thr:g}"]).all()


def test_oracle_sampson_bits_match_reference_gpu(gold, gold_gpu):
    """Pins the oracle's operation order: raw error values of the reference kernel, bit for bit."""
    x1, x2 = gold["sideways_x1"], gold["sideways_x2"]
    E = gold_gpu["sideways_twin_E_list"][:48]
    err = gold_gpu["sideways_twin_err_sample"]
    for m in range(0, 48, 5):
        for j, k in enumerate(range(0, x1.shape[0], 4)):
            if j % 7:
                continue
            mine = oracle.sampson_err(E[m], x1[k, 0], x1[k, 1], x2[k, 0], x2[k, 1])
            assert mine == err[m, j] or (np.isnan(mine) and np.isnan(err[m, j]))


def test_score_edge_cases(engine):
    x = np.tile([[0.1, 0.2]], (40, 1))
    xn = x.copy()
    xn[7, 0] = np.nan
    xn[9, 1] = np.inf
    Efwd = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0]]).reshape(1, 9)
    E = np.concatenate([Efwd, np.zeros((1, 9)), np.full((1, 9), np.nan)])
    cnt = engine.score(dev(xn), dev(x), dev(E), 1e-9).cpu().numpy()
    assert cnt.tolist() == oracle.score(xn, x, E, 1e-9).tolist() == [38, 0, 0]
    # empty hypothesis list / zero points
    assert engine.score(dev(x), dev(x), torch.zeros(0, 9, dtype=torch.float64, device="cuda"), 1.0).numel() == 0
    assert engine.score(dev(x), dev(x), dev(Efwd), 1.0, n_test=0).cpu().tolist() == [0]


def test_guard_band_brackets_exact_counts(engine, std_pair):
    sc, x1, x2 = std_pair
    sets = dev(synth.make_sets(10000, 2048, 3), torch.int32)
    s = engine.solve5(x1, x2, sets)
    idx = torch.arange(10, device="cuda")[None, :] < s["n_valid"][:, None]
    E = s["E"].view(-1, 10, 9)[idx].contiguous()
    exact = engine.score(x1, x2, E, THR)
    for n_test in (10000, 9999, 4097, 130):
        lo, hi = engine.score_bounds(x1, x2, E, THR, n_test=n_test)
        ex = engine.score(x1, x2, E, THR, n_test=n_test)
        assert bool(((lo <= ex) & (ex <= hi)).all())     # rigorous bracket, no exceptions
    lo, hi = engine.score_bounds(x1, x2, E, THR)
    assert float((hi - lo).float().mean()) < 0.02 * 10000  # and a useful one
    assert int((hi >= lo.max()).sum()) < 0.05 * E.shape[0]
    assert bool((exact <= hi).all())


def test_guard_band_adversarial_scales(engine):
    """Bounds must hold for unnormalised E of any magnitude, tiny thresholds' neighbours and
    points far from the principal point."""
    rng = np.random.default_rng(5)
    sc = synth.make_pair(3000, seed=9, noise_px=0.3)
    x1, x2 = sc["x1"] * 3.0, sc["x2"] * 3.0
    Eg = sc["E_gt"] / np.linalg.norm(sc["E_gt"])
    E = np.stack([(Eg + rng.normal(0, s, (3, 3))) * sc_ for s in (0, 1e-6, 1e-4, 1e-2) for sc_ in (1e-8, 1.0, 1e6)
                  for _ in range(20)]).reshape(-1, 9)
    for thr in (3e-4, 1e-3, 1e-2):
        lo, hi = engine.score_bounds(dev(x1), dev(x2), dev(E), thr)
        ex = engine.score(dev(x1), dev(x2), dev(E), thr)
        assert bool(((lo <= ex) & (ex <= hi)).all())
        assert (ex.cpu().numpy() == oracle.score(x1, x2, E, thr)).all()


def test_guard_band_refuses_nonfinite_input(engine):
    import tv5
    x = np.random.default_rng(0).normal(size=(100, 2))
    xb = x.copy()
    xb[3, 1] = np.nan
    with pytest.raises(tv5.Tv5Error):
        engine.score_bounds(dev(xb), dev(x), dev(np.eye(3).reshape(1, 9)), 1e-3)


# ---------------------------------------------------------------------------------------------
# solver: tolerance
# ---------------------------------------------------------------------------------------------
def _compare_solver(mine, ref_E, ref_P, ref_nr, ref_nv, skip=2, same_regime=False):
    """same_regime: the reference numbers were produced on the GPU (nvcc FMA contraction, like
    ours).  Otherwise they come from the host build of the reference, and measured on these very
    fixtures the reference-host and reference-GPU builds differ from EACH OTHER by: 0.5-1 % of
    sets with another real-root count, 1-2 % of solutions off by > 1e-6 and single ones by up to
    1e-1 (ill-conditioned minimal sets) - that is the floor any independent solver can reach."""
    H = ref_nv.shape[0]
    ok = np.ones(H, bool)
    ok[:skip] = False
    nr, nv = mine["n_roots"].cpu().numpy(), mine["n_valid"].cpu().numpy()
    same = (nr == ref_nr) & (nv == ref_nv)
    sel = ok & same
    E = mine["E"].cpu().numpy().reshape(H, 10, 9)
    P = mine["P"].cpu().numpy().reshape(H, 10, 12)
    dE = np.abs(E - ref_E).reshape(H, -1).max(1)[sel] / (np.abs(ref_E).reshape(H, -1).max(1)[sel] + 1)
    dP = np.abs(P - ref_P).reshape(H, -1).max(1)[sel]
    stats = dict(same_counts=float(same[ok].mean()), dE_le_1e6=float((dE < 1e-6).mean()), dP_le_1e6=float((dP < 1e-6).mean()),
                 dE_med=float(np.median(dE)), dE_max=float(dE.max()), dP_max=float(dP.max()), dE_q97=float(np.quantile(dE, 0.97)))
    print("SOLVER_STATS", "gpu" if same_regime else "host", stats)
    # thresholds sit just below what was measured on B200: against the
    # reference compiled for the GPU 100 % equal counts and 98.4-99.5 % of solutions within 1e-6
    assert same[ok].mean() >= (0.999 if same_regime else 0.98)   # measured: 1.0 (GPU build) / 0.984-0.999 (host build)
    # unnormalised E equal in scale, sign and order
    assert np.median(dE) < 1e-9 and np.median(dP) < 1e-9
    if same_regime:
        assert (dE < 1e-6).mean() >= 0.98 and (dP < 1e-6).mean() >= 0.975   # measured: E 0.984 / 0.995, P 0.979 / 0.995
        assert dE.max() < 1e-3 and dP.max() < 5e-3
    else:
        assert (dE < 1e-6).mean() >= 0.94 and (dP < 1e-6).mean() >= 0.94   # measured 0.952 (noisefree) .. 1.0
        assert np.quantile(dE, 0.97) < 1e-3


@pytest.mark.parametrize("name", ["kitti", "noisefree", "sideways", "f64coords"])
def test_solver_vs_reference_host_golden(engine, gold, name):
    x1, x2, sets = gold[f"{name}_x1"], gold[f"{name}_x2"], gold[f"{name}_sets"]
    mine = engine.solve5(dev(x1), dev(x2), dev(sets, torch.int32))
    _compare_solver(mine, gold[f"{name}_E"], gold[f"{name}_P"], gold[f"{name}_n_roots"], gold[f"{name}_n_valid"])


@pytest.mark.parametrize("name", ["kitti", "sideways"])
def test_solver_vs_reference_gpu_golden(engine, gold, gold_gpu, name):
    x1, x2, sets = gold[f"{name}_x1"], gold[f"{name}_x2"], gold[f"{name}_sets"]
    mine = engine.solve5(dev(x1), dev(x2), dev(sets, torch.int32))
    _compare_solver(mine, gold_gpu[f"{name}_twin_E"], gold_gpu[f"{name}_twin_P"], gold_gpu[f"{name}_twin_n_roots"],
                    gold_gpu[f"{name}_twin_n_valid"], same_regime=True)


def test_solver_vs_oracle_large(engine, std_pair):
    sc, x1, x2 = std_pair
    sets = synth.make_sets(10000, 4096, 17)
    mine = engine.solve5(x1, x2, dev(sets, torch.int32))
    orc = oracle.solve_sets(sc["x1"], sc["x2"], sets, True)
    _compare_solver(mine, orc["E"], orc["P"], orc["n_roots"], orc["n_valid"], skip=0)
    ma = engine.solve5(x1, x2, dev(sets, torch.int32), with_cheirality=False)
    oa = oracle.solve_sets(sc["x1"], sc["x2"], sets, False)
    assert (ma["n_valid"].cpu().numpy() == oa["n_valid"]).mean() > 0.99
    assert bool((ma["n_valid"] == ma["n_roots"]).all())


def test_solver_outputs_are_valid_poses(engine, std_pair):
    sc, x1, x2 = std_pair
    s = engine.solve5(x1, x2, dev(synth.make_sets(10000, 1024, 23), torch.int32))
    nv = s["n_valid"].cpu().numpy()
    P = s["P"].cpu().numpy()
    E = s["E"].cpu().numpy()
    n_all = n_close = 0
    for h in range(0, 1024, 7):
        for j in range(nv[h]):
            R, t = P[h, j, :, :3], P[h, j, :, 3]
            assert abs(np.linalg.det(R) - 1) < 1e-6 and np.abs(R @ R.T - np.eye(3)).max() < 1e-6
            assert abs(np.linalg.norm(t) - 1) < 1e-9
            # [t]x R reproduces E for well-conditioned solutions (ill-conditioned roots give an E
            # that is only approximately essential, see solve5.cuh)
            n_all += 1
            n_close += synth.essential_distance(synth.essential_from_pose(R, t), E[h, j]) < 1e-6
    assert n_close > 0.97 * n_all


def test_reference_rng_table(engine, gold_gpu):
    for key in gold_gpu.files:
        if key.startswith("rng_"):
            _, N, iters = key.split("_")
            mine = engine.ref_rng_sets(int(N), int(iters)).cpu().numpy()
            assert (mine == np.minimum(gold_gpu[key], int(N) - 1)).all()   # index table: exact


# ---------------------------------------------------------------------------------------------
# whole pipeline
# ---------------------------------------------------------------------------------------------
def _check_self_consistent(engine, x1h, x2h, r, thr, n_full=None):
    """Size-independent properties: the reported count and mask are exactly the reference
    Sampson decisions for the returned E, and the returned E/P are the solver's for that id."""
    E = r.E.cpu().numpy()
    n = x1h.shape[0] if n_full is None else n_full
    c, m = oracle.score(x1h, x2h, E.reshape(1, 9), thr, n=n, want_mask=True)
    assert r.count == int(c[0])
    if r.mask is not None:
        assert (r.mask.cpu().numpy() == m[0]).all() and int(r.mask.sum()) == r.count


@pytest.mark.parametrize("cheir", [True, False])
def test_pipeline_vs_oracle_ransac(engine, gold, cheir):
    x1h, x2h = gold["kitti_x1"], gold["kitti_x2"]
    sets = synth.make_sets(2000, 1024, 31)
    r = engine.compute_pose(dev(x1h), dev(x2h), 2, THR, sets=dev(sets, torch.int32), with_cheirality=cheir,
                            want_mask=True)
    o = oracle.ransac(x1h, x2h, sets, 2, THR, with_cheirality=cheir)
    _check_self_consistent(engine, x1h, x2h, r, THR)
    # independent solvers differ in the last bits of E, which can move a point across the
    # threshold: same winner, or a count within 2 of the oracle's best
    assert (r.best_set, r.best_root) == (o["best_set"], o["best_root"]) or abs(r.count - o["count"]) <= 2
    assert abs(r.count - o["count"]) <= 2
    assert synth.essential_distance(r.E.cpu().numpy(), o["E"]) < 1e-4
    if cheir:
        assert synth.rotation_error_deg(r.P.cpu().numpy()[:, :3], o["P"][:, :3]) < 1e-2
    else:
        assert float(r.P.abs().sum()) == 0.0


def test_pipeline_fast_path_equals_float64_path(engine, std_pair):
    sc, x1, x2 = std_pair
    sets = dev(synth.make_sets(10000, 4096, 41), torch.int32)
    a = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    engine.set_force_exact(True)
    try:
        b = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    finally:
        engine.set_force_exact(False)
    assert a.fast_path == 1 and b.fast_path == 0
    assert (a.count, a.best_set, a.best_root) == (b.count, b.best_set, b.best_root)
    assert torch.equal(a.E, b.E) and torch.equal(a.P, b.P) and torch.equal(a.mask, b.mask)
    assert a.n_candidates < 0.05 * a.n_hypotheses and b.n_candidates == b.n_hypotheses
    _check_self_consistent(engine, sc["x1"], sc["x2"], a, THR)
    # accuracy against ground truth (degrees)
    P = a.P.cpu().numpy()
    assert synth.rotation_error_deg(P[:, :3], sc["R"]) < 0.05
    assert synth.translation_error_deg(P[:, 3], sc["t"]) < 0.5


def test_pipeline_is_deterministic_and_batch_equals_single(engine):
    ns = [10000, 9999, 777, 5, 2048]
    pairs = [synth.make_pair(n, **{**synth.pair_variation(i), "seed": 50 + i}) for i, n in enumerate(ns)]
    X1 = dev(np.concatenate([p["x1"] for p in pairs]))
    X2 = dev(np.concatenate([p["x2"] for p in pairs]))
    off = np.r_[0, np.cumsum(ns)]
    sets = np.stack([synth.make_sets(n, 1024, 60 + i) for i, n in enumerate(ns)])
    rb = engine.compute_pose_batch(X1, X2, off, 2, THR, sets=dev(sets, torch.int32), want_mask=True)
    rb2 = engine.compute_pose_batch(X1, X2, off, 2, THR, sets=dev(sets, torch.int32), want_mask=True)
    assert torch.equal(rb.E, rb2.E) and torch.equal(rb.stats[:, :3], rb2.stats[:, :3]) and torch.equal(rb.mask, rb2.mask)
    for i, n in enumerate(ns):
        a, b = off[i], off[i + 1]
        rs = engine.compute_pose(X1[a:b].contiguous(), X2[a:b].contiguous(), 2, THR, sets=dev(sets[i], torch.int32),
                                 want_mask=True)
        assert torch.equal(rs.E, rb.E[i]) and torch.equal(rs.P, rb.P[i])
        assert rs.count == int(rb.count[i]) and rs.best_set == int(rb.best_set[i])
        assert torch.equal(rs.mask, rb.mask[a:b])
        _check_self_consistent(engine, pairs[i]["x1"], pairs[i]["x2"], rs, THR)


def _reference_rule_two_stage(engine, x1h, x2h, sets, thr, n_pre, n_full):
    """The reference's selection (kernel_functions.cu:184-219 + essential_matrix.cu:252) applied by
    the CPU oracle's bit-exact scorer to the solutions of OUR solver: per set the first maximum
    over roots on n_pre points, that root re-scored on n_full points, first maximum over sets
    (strict >, starting from 0).  Integer work on identical E bits -> the pipeline must agree
    exactly, whatever the solver's rounding."""
    sol = engine.solve5(dev(x1h), dev(x2h), dev(sets, torch.int32))
    E = sol["E"].cpu().numpy().reshape(-1, 10, 9)
    nv = sol["n_valid"].cpu().numpy()
    best = (0, -1, -1)
    for h in range(E.shape[0]):
        if nv[h] == 0:
            continue
        pre = oracle.score(x1h, x2h, E[h, :nv[h]], thr, n=n_pre)
        j, top = 0, 0
        for r in range(nv[h]):
            if pre[r] > top:
                top, j = int(pre[r]), r
        full = int(oracle.score(x1h, x2h, E[h, j:j + 1], thr, n=n_full)[0])
        if full > best[0]:
            best = (full, h, j)
    return best, E


@pytest.mark.parametrize("split", [1, 0])
def test_two_stage_selection_matches_reference_rule(engine, gold, split):
    """n_pre != n_full (kernel_functions.cu:186-215): all-float64 route.  (count, set, root) equal the
    reference rule exactly; repeated calls are identical (slot order is arbitrary, results are not)."""
    x1h, x2h = gold["kitti_x1"], gold["kitti_x2"]
    sets = synth.make_sets(2000, 512, 71)
    engine.set_split_solver(bool(split))
    try:
        for n_pre, n_full in ((100, 2000), (2000, 500), (10, 1000), (37, 38)):
            (cnt, h, j), E = _reference_rule_two_stage(engine, x1h, x2h, sets, THR, n_pre, n_full)
            runs = [engine.compute_pose(dev(x1h), dev(x2h), 1, THR, n_pre=n_pre, n_full=n_full,
                                        sets=dev(sets, torch.int32), want_mask=True) for _ in range(3)]
            for r in runs:
                assert r.fast_path == 0
                assert (r.count, r.best_set, r.best_root) == (cnt, h, j)                 # integer: exact
                assert (r.E.cpu().numpy().reshape(9) == E[h, j]).all()                   # the winner's own bits
                _check_self_consistent(engine, x1h, x2h, r, THR, n_full=n_full)
            o = oracle.ransac(x1h, x2h, sets, 1, THR, n_pre=n_pre, n_full=n_full)
            # against the oracle's own solver the hypotheses differ in the last bits:
            # counts within 2, and the same winner whenever the counts agree
            assert abs(cnt - o["count"]) <= 2
    finally:
        engine.set_split_solver(True)


def test_no_inliers_gives_zero_result(engine):
    """Defined divergence from the reference's uninitialised output: no hypothesis with
    an inlier -> count 0, best_set -1, E = P = 0.  NaN coordinates make every Sampson value NaN,
    i.e. an outlier (kernel_functions.cu:194 `error <= thr`), on the float64 route."""
    x1h = np.full((200, 2), np.nan)
    x2h = np.full((200, 2), np.nan)
    r = engine.compute_pose(dev(x1h), dev(x2h), 1, THR, sets=dev(synth.make_sets(200, 512, 1), torch.int32),
                            want_mask=True)
    assert r.count == 0 and r.best_set == -1 and r.best_root == -1 and r.fast_path == 0
    assert float(r.E.abs().sum()) == 0.0 and float(r.P.abs().sum()) == 0.0 and int(r.mask.sum()) == 0
    # finite points, threshold far below the rounding of the five sample residuals: the winner, if
    # any, is whatever the exact float64 scorer says about its own E
    rng = np.random.default_rng(3)
    x1h, x2h = rng.uniform(-1, 1, (200, 2)), rng.uniform(-1, 1, (200, 2))
    r = engine.compute_pose(dev(x1h), dev(x2h), 1, 1e-300, sets=dev(synth.make_sets(200, 512, 1), torch.int32))
    if r.count == 0:
        assert r.best_set == -1 and float(r.E.abs().sum()) == 0.0
    else:
        _check_self_consistent(engine, x1h, x2h, r, 1e-300)


def test_reference_rng_default_and_reference_extension_golden(engine, std_pair, gold, gold_gpu):
    """sets=None must behave like the reference: same index table, and - when the golden file
    holds outputs of the unmodified extension - same winner within tolerance."""
    sc, x1, x2 = std_pair
    r = engine.compute_pose(x1, x2, 8, THR)
    tab = engine.ref_rng_sets(10000, 8)
    r2 = engine.compute_pose(x1, x2, 8, THR, sets=tab)
    assert torch.equal(r.E, r2.E) and r.count == r2.count
    _check_self_consistent(engine, sc["x1"], sc["x2"], r, THR)
    key = "refext_std10k_8_count"
    if key in gold_gpu.files:
        assert abs(r.count - int(gold_gpu[key])) <= 3
        assert synth.essential_distance(r.E.cpu().numpy(), gold_gpu["refext_std10k_8_E"]) < 1e-4
        assert synth.rotation_error_deg(r.P.cpu().numpy()[:, :3], gold_gpu["refext_std10k_8_P"][:, :3]) < 1e-2
        # and the reference's own count is exactly our exact score of the reference's E
        c = engine.score(x1, x2, dev(gold_gpu["refext_std10k_8_E"].reshape(1, 9)), THR)
        assert int(c[0]) == int(gold_gpu[key])


def test_full_size_dense_pair_properties(engine):
    """Config-4 shape (453,620 correspondences); properties only, the oracle would take minutes."""
    sc = synth.make_pair(dense=True, seed=4)
    x1, x2 = dev(sc["x1"]), dev(sc["x2"])
    sets = dev(synth.make_sets(sc["x1"].shape[0], 1024, 5), torch.int32)
    r = engine.compute_pose(x1, x2, 2, THR, sets=sets, want_mask=True)
    assert r.fast_path == 1 and int(r.mask.sum()) == r.count
    c = engine.score(x1, x2, r.E.reshape(1, 9), THR)
    assert int(c[0]) == r.count
    P = r.P.cpu().numpy()
    assert synth.rotation_error_deg(P[:, :3], sc["R"]) < 0.05 and synth.translation_error_deg(P[:, 3], sc["t"]) < 0.5
    assert r.count > 0.6 * sc["inlier_gt"].sum()


def test_config4_full_size_counts_equal_reference_scorer(engine):
    """configs[3] at its full size — 453,620 dense correspondences x 16,384 minimal sets — checked
    against the REFERENCE's own ComputeError<double> compiled by nvcc (oracle/_ref/libref_twin_cuda.so,
    reference sources included verbatim): every hypothesis of our solver is scored by the reference
    scorer on the GPU box, the reference's first-maximum rule is applied on the host, and the
    pipeline's (count, set, root) must be that winner exactly (guard-band scorer included)."""
    import ref_twin
    T = ref_twin.load()
    if T is None:
        pytest.skip("oracle/_ref/libref_twin_cuda.so not built")
    sc = synth.make_pair(dense=True, seed=4)
    n = sc["x1"].shape[0]
    assert n == 453620
    x1, x2 = dev(sc["x1"]), dev(sc["x2"])
    sets = dev(synth.make_sets(n, 16384, 9), torch.int32)
    r = engine.compute_pose(x1, x2, 32, THR, sets=sets)
    sol = engine.solve5(x1, x2, sets)
    nv = sol["n_valid"]
    valid = torch.arange(10, device="cuda")[None, :] < nv[:, None]
    E_list = sol["E"].reshape(-1, 10, 9)[valid].contiguous()
    ids = (torch.arange(16384, device="cuda")[:, None] * 16 + torch.arange(10, device="cuda")[None, :])[valid]
    assert E_list.shape[0] == r.n_hypotheses
    counts = ref_twin.score(T, x1, x2, n, E_list, THR)            # the reference's scorer, FP64
    best = int(counts.max())
    first = int(ids[counts == best].min())                          # first maximum by (set, root)
    assert (r.count, r.best_set, r.best_root) == (best, first >> 4, first & 15)
    assert r.fast_path == 1 and r.n_candidates < 200               # the float32 bound did the pruning


# ---------------------------------------------------------------------------------------------
# drop-in module
# ---------------------------------------------------------------------------------------------
def test_shim_computeP_contract(std_pair):
    import essential_matrix as em
    sc, x1, x2 = std_pair
    E, P, n = em.computeP(x1, x2, 10000, 10000, 5, THR)
    assert E.shape == (3, 3) and P.shape == (3, 4) and E.dtype == torch.float64 and E.is_cuda and P.is_cuda
    assert int(n) > 5000 and n == int(n) and f"{n}" == str(int(n))
    E2 = em.initialise(x1, x2, 10000, 10000, 5, THR)
    assert E2.shape == (3, 3) and E2.is_cuda
    for bad, msg in ((x1.cpu(), "input1 must be a CUDA tensor"), (x1.float(), "input1 must be a double tensor"),
                     (x1.t().contiguous().t(), "input1 must be contiguous")):
        with pytest.raises(RuntimeError, match=msg):
            em.computeP(bad, x2, 10000, 10000, 5, THR)


def test_shim_in_sfmnet_call_shape(std_pair):
    """The call sequence of epipolar_utils.compute_P_matrix_ransac (epipolar_utils.py:112-135)
    with float32 inputs, as models/SFMnet.py:259-270 produces them."""
    import essential_matrix as em
    sc, x1, x2 = std_pair
    c1, c2 = x1.float(), x2.float()
    K = torch.tensor(sc["K"], dtype=torch.float32, device="cuda")
    Kinv = torch.inverse(K)
    E, P, n = em.computeP(c1.double(), c2.double(), c1.shape[0], c1.shape[0], 5, THR)
    E = E.float()
    F = Kinv.t() @ E @ Kinv
    assert F.shape == (3, 3) and torch.isfinite(F).all() and torch.isfinite(P).all()
    Pm = P.cpu().numpy()
    assert synth.rotation_error_deg(Pm[:, :3], sc["R"]) < 0.05


# ---------------------------------------------------------------------------------------------
# decomposition and refinement (the reference's polish_E.cu host functions, here on the GPU)
# ---------------------------------------------------------------------------------------------
POLISH_KEYS = ("0.0001_1_0", "0.0001_1_1", "0.0001_1_10", "0.0001_0_10", "0.0005_0.5_200")


@pytest.fixture(scope="module")
def gold_polish(golden_dir):
    return np.load(os.path.join(golden_dir, "polish_ref.npz"))


def test_decompose_batch_bit_exact_vs_reference(engine, gold_polish):
    g = gold_polish
    r = engine.decompose_batch(dev(g["E"]))
    assert (r["U"].cpu().numpy() == g["U"]).all()            # no contraction on the device: bit-exact
    assert (r["V"].cpu().numpy() == g["V"]).all()
    # atan2 is not correctly rounded on either side: 2 ulp
    assert np.abs(r["angles"].cpu().numpy() - g["angles"]).max() < 1e-15
    import essential_matrix as em
    U, V = em.decomposeUV(dev(g["E"][3]))
    assert U.is_cuda and (U.cpu().numpy() == g["U"][3]).all() and (V.cpu().numpy() == g["V"][3]).all()
    assert em.decompose(dev(g["E"][3])).shape == (5,)


@pytest.mark.parametrize("case", range(4))
def test_optimise_matches_reference_golden(engine, gold_polish, case):
    """Golden E of the reference extension's own `optimise`.  The GPU sums J^T W e / J^T W J in
    a tree instead of sequentially, so agreement is to rounding: 1e-9 absolute on ||E||_F = sqrt 2
    (200-iteration case: 1e-7, rounding differences are amplified by the weight switch at delta)."""
    import essential_matrix as em
    g = gold_polish
    x1, x2, E0 = g[f"opt{case}_x1"], g[f"opt{case}_x2"], g[f"opt{case}_E0"]
    for key in POLISH_KEYS:
        delta, alpha, reps = key.split("_")
        tol = 1e-7 if int(reps) > 10 else 1e-9
        ref = g[f"opt{case}_{key}"]
        E = engine.optimise(dev(x1), dev(x2), dev(E0), float(delta), float(alpha), int(reps)).cpu().numpy()
        assert np.abs(E - ref).max() < tol, (key, np.abs(E - ref).max())
        # the reference's calling convention: CPU tensors in, CPU tensor out
        Ec = em.optimise(torch.from_numpy(x1), torch.from_numpy(x2), torch.from_numpy(E0.copy()),
                         float(delta), float(alpha), int(reps))
        assert not Ec.is_cuda and np.abs(Ec.numpy() - ref).max() < tol
        assert np.abs(oracle.optimise(x1, x2, E0, float(delta), float(alpha), int(reps)) - ref).max() == 0.0


def test_optimise_edge_cases_and_determinism(engine, gold_polish):
    g = gold_polish
    x1, x2, E0 = dev(g["opt2_x1"]), dev(g["opt2_x2"]), dev(g["opt2_E0"])
    a, ia = engine.optimise(x1, x2, E0, 1e-4, 1.0, 10, want_iters=True)
    b, ib = engine.optimise(x1, x2, E0, 1e-4, 1.0, 10, want_iters=True)
    assert torch.equal(a, b) and int(ia) == int(ib) == 10           # run-to-run identical
    # no points: the gradient is zero and, like the reference (break before Eprod,
    # polish_E.cu:1545), the half-reduced working matrix of the decomposition comes back
    z = torch.zeros(0, 2, dtype=torch.float64, device="cuda")
    E_none = engine.optimise(z, z, E0, 1e-4, 1.0, 5).cpu().numpy()
    assert (E_none == oracle.optimise(np.zeros((0, 2)), np.zeros((0, 2)), g["opt2_E0"], 1e-4, 1.0, 5)).all()
    # mask == inlier subset is the same as passing the subset
    m = (torch.arange(x1.shape[0], device="cuda") % 3 != 0)
    sub = engine.optimise(x1[m].contiguous(), x2[m].contiguous(), E0, 1e-4, 1.0, 6)
    msk = engine.optimise(x1, x2, E0, 1e-4, 1.0, 6, mask=m.to(torch.uint8))
    assert (sub - msk).abs().max() < 1e-10
    # a batch of problems equals the problems one by one
    n = [2000, 500, 10000, 64]
    xs1 = torch.cat([dev(g[f"opt{i}_x1"]) for i in range(4)])
    xs2 = torch.cat([dev(g[f"opt{i}_x2"]) for i in range(4)])
    E0s = torch.stack([dev(g[f"opt{i}_E0"]) for i in range(4)])
    Eb, it = engine.optimise_batch(xs1, xs2, np.concatenate([[0], np.cumsum(n)]), E0s, 1e-4, 1.0, 10)
    for i in range(4):
        ref = g[f"opt{i}_0.0001_1_10"]
        assert np.abs(Eb[i].cpu().numpy() - ref).max() < 1e-9
    assert (it.cpu().numpy() <= 10).all()


def test_local_optimisation_after_ransac_improves_pose(engine, std_pair):
    """LO step: refine the RANSAC winner on its own inlier mask (the use SURVEY 8(f) f2 names)."""
    sc, x1, x2 = std_pair
    r = engine.compute_pose(x1, x2, 8, THR, want_mask=True)
    E = engine.optimise(x1, x2, r.E, THR, 1.0, 10, mask=r.mask).cpu().numpy()
    d0 = synth.essential_distance(r.E.cpu().numpy(), sc["E_gt"])
    d1 = synth.essential_distance(E, sc["E_gt"])
    assert d1 <= d0 * 1.05 and d1 < 2e-3
    cnt = engine.score(x1, x2, dev(E.reshape(1, 9)), THR).cpu().numpy()[0]
    assert cnt >= r.count - 20


# ---------------------------------------------------------------------------------------------
# flow -> correspondences -> pose (front of SFMnet.pose_by_ransac)
# ---------------------------------------------------------------------------------------------
def _flow_batch(hw, B, seed=0):
    scs = [synth.make_flow(hw=hw, seed=seed + i, rvec=(0.002, 0.01 + 0.002 * i, -0.001)) for i in range(B)]
    flow = np.stack([s["flow"] for s in scs])
    Kinv = np.stack([s["Kinv"] for s in scs])
    return scs, flow, Kinv


@pytest.mark.parametrize("hw", [(48, 80), (370, 1226)])
def test_flow_to_points_matches_oracle_and_reference_torch_chain(engine, hw):
    import ref_flow_torch as rf
    H, W = hw
    B = 2
    scs, flow, Kinv = _flow_batch(hw, B)
    tf, tk = dev(flow, torch.float32), dev(Kinv, torch.float32)
    rng = np.random.default_rng(9)
    n = [37, 64]
    pts_i = np.stack([rng.integers(0, W, sum(n)), rng.integers(0, H, sum(n))], 1).astype(np.int32)
    pts_f = np.stack([rng.uniform(0, W - 1, sum(n)), rng.uniform(0, H - 1, sum(n))], 1).astype(np.float32)
    pts_f[0] = (W - 1, H - 1)
    off = np.array([0, n[0], sum(n)])
    ulp = float(np.finfo(np.float32).eps)
    for name, pts in (("crop", None), ("gather", pts_i), ("bilinear", pts_f)):
        x1, x2, o = engine.flow_to_points(tf, tk, 10, None if pts is None else torch.from_numpy(pts),
                                          None if pts is None else off)
        x1, x2 = x1.cpu().numpy(), x2.cpu().numpy()
        for b in range(B):
            p = None if pts is None else pts[off[b]:off[b + 1]]
            oa, oc = oracle.flow_to_points(flow[b], Kinv[b], 10, p)
            ma, mc = x1[o[b]:o[b + 1]], x2[o[b]:o[b + 1]]
            assert ma.shape == oa.shape
            # same float32 operation sequence as the oracle: bit-exact, except the bilinear mode
            # where the oracle's emulated fma can double-round (1 ulp)
            tol = 0.0 if name != "bilinear" else ulp * max(1.0, np.abs(oc).max())
            assert np.abs(ma - oa).max() <= tol and np.abs(mc - oc).max() <= tol, name
            # the reference's own torch ops on this GPU (cuBLAS bmm, cuDNN-free grid_sample): 1 ulp
            kw = {} if pts is None else (dict(pts=p.astype(np.float64)) if name == "gather"
                                         else dict(pts=p.astype(np.float64), sample_sp=True))
            ta, tc = rf.points_of_image(tf, tk, b, 10, **kw)
            tol = ulp * max(1.0, np.abs(oc).max())
            if name == "bilinear":
                # torch's CUDA grid_sample may round its four-tap sums differently (no documented
                # operation order): allow 4 ulp of the pixel coordinate (<= W) scaled by K^-1
                tol = max(tol, 4 * ulp * W * float(np.abs(Kinv[b][0, 0])))
            assert np.abs(ma - ta.cpu().numpy()).max() <= tol, name
            assert np.abs(mc - tc.cpu().numpy()).max() <= tol, name


def test_pose_from_flow_equals_pose_on_reference_points(engine):
    """One fused submission == the reference's chain followed by computeP, image by image."""
    hw = (120, 400)
    B = 3
    scs, flow, Kinv = _flow_batch(hw, B, seed=20)
    tf, tk = dev(flow, torch.float32), dev(Kinv, torch.float32)
    P32, E32, r = engine.pose_from_flow(tf, tk, 2, THR, margin=10)
    assert P32.dtype == torch.float32 and P32.shape == (B, 3, 4) and E32.shape == (B, 3, 3)
    assert torch.equal(E32, r.E.float()) and torch.equal(P32, r.P.float())
    for b in range(B):
        oa, oc = oracle.flow_to_points(flow[b], Kinv[b], 10)
        one = engine.compute_pose(dev(oa), dev(oc), 2, THR)
        assert one.count == r.count[b] and one.best_set == r.best_set[b]
        assert torch.equal(one.E, r.E[b]) and torch.equal(one.P, r.P[b])
        Pn = r.P[b].cpu().numpy()
        assert synth.rotation_error_deg(Pn[:, :3], scs[b]["R"]) < 0.05       # degrees, vs ground truth
        assert synth.translation_error_deg(Pn[:, 3], scs[b]["t"]) < 1.0
    # keypoint list (integer gather), ragged per image
    rng = np.random.default_rng(1)
    n = [500, 2000, 64]
    pts = np.stack([rng.integers(10, hw[1] - 10, sum(n)), rng.integers(10, hw[0] - 10, sum(n))], 1).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(n)])
    P32, E32, r = engine.pose_from_flow(tf, tk, 2, THR, pts=torch.from_numpy(pts), offsets=off)
    for b in range(B):
        oa, oc = oracle.flow_to_points(flow[b], Kinv[b], 10, pts[off[b]:off[b + 1]])
        one = engine.compute_pose(dev(oa), dev(oc), 2, THR)
        assert one.count == r.count[b] and torch.equal(one.E, r.E[b])


# ---------------------------------------------------------------------------------------------
# plane-sweep cost volume (consumer of P, models/PSNet.py:141-157)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("by_depth", [False, True])
def test_plane_sweep_matches_oracle_and_reference_torch_loop(engine, by_depth):
    import ref_planesweep_torch as rp
    from test_oracle import sweep_case, sweep_tolerance
    B, C, h, w, L = 2, 4, 24, 40, 8
    ref, tgt, pose, K4, Kinv4 = sweep_case(B, C, h, w)
    dv = [dev(a, torch.float32) for a in (ref, tgt, pose, K4, Kinv4)]
    cost = engine.plane_sweep(*dv, L, 1.0, by_depth)
    assert cost.shape == (B, 2 * C, L, h, w) and cost.dtype == torch.float32
    mine = cost.cpu().numpy()
    tol = sweep_tolerance(h, w, tgt)
    ct = rp.cost_volume(*dv, L, 1.0, by_depth).cpu().numpy()      # the reference's torch ops on this GPU
    for b in range(B):
        co = oracle.plane_sweep_cost_volume(ref[b], tgt[b], pose[b], K4[b], Kinv4[b], L, 1.0, by_depth)
        assert (mine[b, :C] == co[:C]).all() and (mine[b, :C] == ct[b, :C]).all()   # copies: exact
        assert np.abs(mine[b, C:] - co[C:]).max() <= tol
        assert np.abs(mine[b, C:] - ct[b, C:]).max() <= tol
        # same operation sequence as the oracle: all but a handful of samples are bit-identical
        assert (mine[b, C:] == co[C:]).mean() > 0.99


def test_plane_sweep_full_size_properties(engine):
    """PSNet's real shape (32 channels, 128 planes, quarter-resolution KITTI): identity pose at
    any depth reproduces the target features; a NaN pose produces zeros, not a fault."""
    import ref_planesweep_torch as rp
    B, C, h, w, L = 1, 32, 92, 306, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    ref = torch.randn(B, C, h, w, device="cuda", generator=g)
    tgt = torch.randn(B, C, h, w, device="cuda", generator=g)
    K = synth.KITTI_K.copy(); K[:2] /= 4.0
    K4 = dev(K[None], torch.float32); Kinv4 = dev(np.linalg.inv(K)[None], torch.float32)
    eye = dev(np.concatenate([np.eye(3), np.zeros((3, 1))], 1)[None], torch.float32)
    cost = engine.plane_sweep(ref, tgt, eye, K4, Kinv4, L, 1.0)
    assert torch.equal(cost[:, :C], ref[:, :, None].expand(B, C, L, h, w))
    # identity warp: sample positions are the pixel centres up to float32 rounding (the
    # border pixels can round to just outside [-1,1] and are then zeroed, as in the reference:
    # compare inside)
    assert (cost[:, C:, :, 1:-1, 1:-1] - tgt[:, :, None, 1:-1, 1:-1]).abs().max() < 8 * 1.2e-7 * w * 2 * float(tgt.abs().max())
    # one plane against the reference's torch ops at full size
    sc = synth.make_pair(10, seed=1)
    pose = dev(np.concatenate([sc["R"], sc["t"][:, None]], 1)[None], torch.float32)
    cost = engine.plane_sweep(ref, tgt, pose, K4, Kinv4, L, 1.0)
    ct = rp.cost_volume(ref, tgt, pose, K4, Kinv4, L, 1.0)
    tol = 8 * 1.2e-7 * w * 2 * float(tgt.abs().max())
    diff = (cost - ct).abs()
    # planes at depth >= 4 (i < 32): the projection is well conditioned, strict bound.  Nearer planes
    # (depth -> |t_z|, Z -> 0) amplify the float32 rounding of X/Z without bound, in the reference
    # as much as here: there only the fraction of samples beyond the bound is limited.
    assert diff[:, :, :32].max() < tol
    assert (diff > tol).float().mean() < 1e-3
    bad = engine.plane_sweep(ref, tgt, pose * float("nan"), K4, Kinv4, 4, 1.0)
    assert torch.equal(bad[:, C:], torch.zeros_like(bad[:, C:]))


# ---------------------------------------------------------------------------------------------
# chunked submissions: host-buffer entry point and the optional solver/scorer overlap
# ---------------------------------------------------------------------------------------------
def test_chunked_host_path_and_overlap_equal_plain_submission(engine):
    """40 ragged pairs: (a) one plain device submission, (b) the host-buffer entry point (chunks of
    pairs start as their copies land), (c) the two-stream overlap mode — identical results."""
    rng = np.random.default_rng(3)
    ns = [int(n) for n in rng.integers(300, 3000, 40)]
    ns[7], ns[23] = 5, 2999
    pairs = [synth.make_pair(n, **{**synth.pair_variation(i), "seed": 900 + i}) for i, n in enumerate(ns)]
    x1h = np.concatenate([p["x1"] for p in pairs]); x2h = np.concatenate([p["x2"] for p in pairs])
    off = np.r_[0, np.cumsum(ns)]
    sets = np.stack([synth.make_sets(n, 512, 950 + i) for i, n in enumerate(ns)])
    plain = engine.compute_pose_batch(dev(x1h), dev(x2h), off, 1, THR, sets=dev(sets, torch.int32))
    Eh, Ph, sth = engine.compute_pose_batch_host(x1h, x2h, off, 1, THR, sets=sets)
    assert (Eh == plain.E.cpu().numpy()).all() and (Ph == plain.P.cpu().numpy()).all()
    assert (sth[:, :5] == plain.stats.cpu().numpy()[:, :5]).all()
    try:
        engine.set_overlap(True)
        assert engine.pipeline_chunks(40) == 2
        ov = engine.compute_pose_batch(dev(x1h), dev(x2h), off, 1, THR, sets=dev(sets, torch.int32), want_mask=True)
        Eh2, Ph2, sth2 = engine.compute_pose_batch_host(x1h, x2h, off, 1, THR, sets=sets)
    finally:
        engine.set_overlap(False)
    assert torch.equal(ov.E, plain.E) and torch.equal(ov.P, plain.P)
    assert torch.equal(ov.stats[:, :5], plain.stats[:, :5])
    assert (Eh2 == Eh).all() and (sth2[:, :5] == sth[:, :5]).all()
    # the masks returned by the overlapped run are the winners' exact masks
    for i in (0, 7, 23, 39):
        a, b = off[i], off[i + 1]
        c, m = oracle.score(x1h[a:b], x2h[a:b], ov.E[i].cpu().numpy().reshape(1, 9), THR, want_mask=True)
        assert int(c[0]) == int(ov.count[i]) and (ov.mask[a:b].cpu().numpy() == m[0]).all()


# ---------------------------------------------------------------------------------------------
# C ABI error behaviour: codes, never exit(), never print (essential_matrix.cu:17-24 calls exit)
# ---------------------------------------------------------------------------------------------
def test_cabi_rejects_bad_arguments_with_error_codes(engine):
    import ctypes as C
    L, ctx = engine.L, engine.ctx
    x = dev(np.random.default_rng(0).normal(size=(64, 2)))
    E = torch.empty(9, dtype=torch.float64, device="cuda")
    P = torch.empty(12, dtype=torch.float64, device="cuda")
    res = torch.zeros(8, dtype=torch.int32, device="cuda")
    ok = (ctx, None, x.data_ptr(), x.data_ptr(), 64, None, 1, 64, 64, 1e-3, 1, E.data_ptr(), P.data_ptr(), res.data_ptr(), None)

    def call(**kw):
        a = list(ok)
        for k, v in kw.items():
            a[int(k[1:])] = v
        return L.tv5_compute_pose(*a)

    assert call() == 0
    assert call(_0=None) == -1                      # no context
    assert call(_2=None) == -1                      # x1 null
    assert call(_4=0) == -1 and call(_4=-5) == -1   # N < 1
    assert call(_6=0) == -1                         # iters < 1
    assert call(_9=0.0) == -1 and call(_9=float("nan")) == -1 and call(_9=-1.0) == -1   # threshold
    assert call(_11=None) == -1 and call(_13=None) == -1                                 # outputs
    assert L.tv5_strerror(-1) == b"invalid argument" and L.tv5_strerror(-2) == b"CUDA runtime error"
    assert L.tv5_score(ctx, None, x.data_ptr(), x.data_ptr(), 64, None, 3, 1e-3, res.data_ptr(), None) == -1
    assert L.tv5_plane_sweep(ctx, None, None, None, None, None, None, 1, 1, 4, 4, 1, C.c_float(1.0), 0, None) == -1
    assert L.tv5_flow_to_points(ctx, None, x.data_ptr(), 1, 8, 8, x.data_ptr(), 0, 4, None, None, E.data_ptr(), E.data_ptr()) == -1  # margin eats the image
    assert L.tv5_optimise(ctx, None, x.data_ptr(), x.data_ptr(), -1, None, E.data_ptr(), 1e-4, 1.0, 3, None) == -1
    rec = torch.zeros(24, dtype=torch.float64, device="cuda")
    assert L.tv5_winner_record(ctx, None, E.data_ptr(), P.data_ptr(), res.data_ptr(), -1, rec.data_ptr()) == -1   # offset < 0
    assert L.tv5_winner_record(ctx, None, None, P.data_ptr(), res.data_ptr(), 0, rec.data_ptr()) == -1
    assert L.tv5_winner_pick(ctx, None, rec.data_ptr(), 0, E.data_ptr(), P.data_ptr(), res.data_ptr()) == -1      # no records
    assert L.tv5_debug_guard(ctx, 1, 0xFF) == -1    # only on a context that has not allocated anything yet
    bad, nb = C.c_int64(), C.c_int32()
    assert L.tv5_debug_poison(ctx, 0) == -1 and L.tv5_debug_stray_write(ctx, 1) == -1     # not in guard mode
    assert L.tv5_debug_check_guards(ctx, C.byref(bad), C.byref(nb)) == 0 and bad.value == 0 and nb.value == 0
    torch.cuda.synchronize()                        # the context is still healthy
    assert call() == 0


def test_fewer_than_five_points_and_degenerate_sets(engine):
    """N < 5 forces repeated indices in every minimal set (the reference's RNG does the same):
    the solver reports no solution instead of faulting, the pose call returns the zero result."""
    x1 = dev(np.array([[0.1, 0.2], [0.3, -0.1], [-0.2, 0.05]]))
    x2 = dev(np.array([[0.11, 0.2], [0.31, -0.12], [-0.19, 0.04]]))
    r = engine.compute_pose(x1, x2, 1, THR)
    assert r.count >= 0 and torch.isfinite(r.E).all() and torch.isfinite(r.P).all()
    sets = np.zeros((512, 5), np.int32)             # every set = the same point five times
    r = engine.compute_pose(x1, x2, 1, THR, sets=dev(sets, torch.int32))
    assert r.count == 0 and r.best_set == -1 and float(r.E.abs().sum()) == 0.0


@pytest.mark.parametrize("seed", range(4))
def test_scoring_random_property_counts_equal_oracle(engine, seed):
    """Random (not geometrically meaningful) E, points and thresholds, ragged sizes: exact counts,
    guard-band brackets and masks against the oracle."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1, 3000))
    M = int(rng.integers(1, 300))
    x1 = rng.normal(0, rng.uniform(0.05, 2.0), (n, 2))
    x2 = x1 + rng.normal(0, 10.0 ** rng.uniform(-5, -1), (n, 2))
    E = rng.normal(0, 1, (M, 9)) * 10.0 ** rng.uniform(-3, 3, (M, 1))
    thr = 10.0 ** rng.uniform(-5, -1)
    c_or, m_or = oracle.score(x1, x2, E, thr, want_mask=True)
    cnt, masks = engine.score(dev(x1), dev(x2), dev(E), thr, want_mask=True)
    assert (cnt.cpu().numpy() == c_or).all() and (unpack_masks(masks, n) == m_or).all()
    try:
        lo, hi = engine.score_bounds(dev(x1), dev(x2), dev(E), thr)
    except Exception:
        return                                      # band too wide for float32: the pipeline uses float64
    lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
    assert (lo <= c_or).all() and (c_or <= hi).all()


# ---------------------------------------------------------------------------------------------
# solver organisation: three kernels (default) vs the fused kernel
# ---------------------------------------------------------------------------------------------
def test_split_solver_is_bit_identical_to_fused_solver(engine, std_pair):
    sc, x1, x2 = std_pair
    sets = dev(synth.make_sets(10000, 4096, 77), torch.int32)
    try:
        engine.set_split_solver(False)
        f = engine.solve5(x1, x2, sets, True)
        f0 = engine.solve5(x1, x2, sets, False)
        rf = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    finally:
        engine.set_split_solver(True)
    s = engine.solve5(x1, x2, sets, True)
    s0 = engine.solve5(x1, x2, sets, False)
    rs = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    for a, b in ((f, s), (f0, s0)):
        for k in ("E", "P", "n_roots", "n_valid"):
            assert torch.equal(a[k], b[k]), k          # same arithmetic sequence per root: bit-identical
    assert torch.equal(rf.E, rs.E) and torch.equal(rf.P, rs.P) and torch.equal(rf.mask, rs.mask)
    assert torch.equal(rf.stats[:5], rs.stats[:5])      # count, set, (compacted) root, M, candidates


# ---------------------------------------------------------------------------------------------
# early exit (opt-in): staged scoring with exact hypothesis pruning
# ---------------------------------------------------------------------------------------------
def test_early_exit_gives_identical_results(engine, std_pair):
    """Winner, count, E, P, mask and the hypothesis total are those of the full scoring, on the
    standard pair, on ragged batches (including pairs too small to stage, a pair without inliers
    and a float64-path pair), and with the reference's minimal sets."""
    sc, x1, x2 = std_pair
    rng = np.random.default_rng(5)
    ns = [10000, 9999, 777, 5, 2048, 300, 4097, 1500, 17, 64, 130, 257]   # small ones: stages collapse
    pairs = [synth.make_pair(n, **{**synth.pair_variation(i), "seed": 400 + i}) for i, n in enumerate(ns)]
    pairs[5]["x2"] = rng.normal(0, 0.5, pairs[5]["x2"].shape)            # pure outliers
    pairs[7]["x1"] = pairs[7]["x1"] * 3000.0                              # |x| > 1024: float64 scorer
    X1 = dev(np.concatenate([p["x1"] for p in pairs])); X2 = dev(np.concatenate([p["x2"] for p in pairs]))
    off = np.r_[0, np.cumsum(ns)]
    sets = dev(np.stack([synth.make_sets(n, 2048, 60 + i) for i, n in enumerate(ns)]), torch.int32)

    def run():
        a = engine.compute_pose(x1, x2, 8, THR, want_mask=True)            # reference RNG table
        b = engine.compute_pose_batch(X1, X2, off, 4, THR, sets=sets, want_mask=True)
        return a, b

    full = run()
    try:
        engine.set_early_exit(True)
        fast = run()
    finally:
        engine.set_early_exit(False)
    for f, e in zip(full, fast):
        assert torch.equal(f.E, e.E) and torch.equal(f.P, e.P) and torch.equal(f.mask, e.mask)
        sf, se = f.stats.reshape(-1, 8), e.stats.reshape(-1, 8)
        assert torch.equal(sf[:, :4], se[:, :4])        # count, set, root, total hypotheses
        assert torch.equal(sf[:, 5], se[:, 5])          # same scorer path per pair
    assert int(fast[1].fast_path[7]) == 0 and int(fast[1].fast_path[0]) == 1


@pytest.mark.parametrize("outlier_frac,noise_px", [(0.0, 0.0), (0.5, 0.05), (0.85, 0.3), (1.0, 0.05)])
def test_early_exit_identical_across_inlier_regimes(engine, outlier_frac, noise_px):
    """Exactness of the pruning does not depend on how many inliers there are: noise-free (every
    all-inlier set ties at count N), half outliers, mostly outliers with heavy noise, no structure."""
    B, n = 8, 6000
    pairs = [synth.make_pair(n, seed=700 + i, outlier_frac=outlier_frac, noise_px=noise_px) for i in range(B)]
    X1 = dev(np.concatenate([p["x1"] for p in pairs])); X2 = dev(np.concatenate([p["x2"] for p in pairs]))
    off = np.arange(B + 1) * n
    sets = dev(np.stack([synth.make_sets(n, 4096, 800 + i) for i in range(B)]), torch.int32)
    full = engine.compute_pose_batch(X1, X2, off, 8, THR, sets=sets, want_mask=True)
    try:
        engine.set_early_exit(True)
        fast = engine.compute_pose_batch(X1, X2, off, 8, THR, sets=sets, want_mask=True)
    finally:
        engine.set_early_exit(False)
    assert torch.equal(full.E, fast.E) and torch.equal(full.P, fast.P) and torch.equal(full.mask, fast.mask)
    assert torch.equal(full.stats[:, :4], fast.stats[:, :4])


@pytest.mark.parametrize("shape", [(1, 1, 5, 7, 3), (2, 3, 9, 33, 5), (1, 2, 2, 2, 1), (1, 5, 17, 31, 33)])
def test_plane_sweep_odd_shapes(engine, shape):
    """Shapes that exercise the per-plane store shift (h*w odd, nlabel not a multiple of 32), one
    channel, the smallest image grid_sample accepts."""
    import ref_planesweep_torch as rp
    from test_oracle import sweep_tolerance
    B, C, h, w, L = shape
    rng = np.random.default_rng(sum(shape))
    ref = rng.normal(0, 1, (B, C, h, w)).astype(np.float32)
    tgt = rng.normal(0, 1, (B, C, h, w)).astype(np.float32)
    K = np.array([[0.9 * w, 0, 0.5 * w], [0, 0.9 * w, 0.5 * h], [0, 0, 1.0]])
    K4 = np.stack([K] * B).astype(np.float32); Kinv4 = np.stack([np.linalg.inv(K)] * B).astype(np.float32)
    R = synth.rodrigues((0.01, -0.02, 0.005)); t = np.array([0.1, -0.05, -0.3])
    pose = np.stack([np.concatenate([R, t[:, None]], 1)] * B).astype(np.float32)
    dv = [dev(a, torch.float32) for a in (ref, tgt, pose, K4, Kinv4)]
    mine = engine.plane_sweep(*dv, L, 1.0).cpu().numpy()
    ct = rp.cost_volume(*dv, L, 1.0).cpu().numpy()
    assert mine.shape == (B, 2 * C, L, h, w)
    assert (mine[:, :C] == ct[:, :C]).all()
    for b in range(B):
        co = oracle.plane_sweep_cost_volume(ref[b], tgt[b], pose[b], K4[b], Kinv4[b], L, 1.0)
        assert np.abs(mine[b, C:] - co[C:]).max() <= sweep_tolerance(h, w, tgt)
    # against torch: the same bound wherever the projection is well conditioned (Z >= 1: depth >= 1.3)
    far = [i for i in range(L) if 1.0 * L / (i + 1) >= 1.3]
    if far:
        assert np.abs(mine[:, C:, far] - ct[:, C:, far]).max() <= 4 * sweep_tolerance(h, w, tgt)


def test_flow_to_points_small_images_and_zero_margin(engine):
    rng = np.random.default_rng(2)
    for (H, W, margin) in ((3, 4, 0), (8, 8, 3), (21, 5, 2)):
        flow = rng.normal(0, 1, (2, 2, H, W)).astype(np.float32)
        Kinv = np.stack([np.linalg.inv(synth.KITTI_K)] * 2).astype(np.float32)
        x1, x2, off = engine.flow_to_points(dev(flow, torch.float32), dev(Kinv, torch.float32), margin)
        assert off[-1] == 2 * (H - 2 * margin) * (W - 2 * margin)
        for b in range(2):
            oa, oc = oracle.flow_to_points(flow[b], Kinv[b], margin) if margin else _flow_margin0(flow[b], Kinv[b])
            assert (x1[off[b]:off[b + 1]].cpu().numpy() == oa).all() and (x2[off[b]:off[b + 1]].cpu().numpy() == oc).all()


def _flow_margin0(flow, Kinv):
    """oracle with margin 0 (numpy's [0:-0] slice is empty, so go through the gather form)."""
    _, H, W = flow.shape
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return oracle.flow_to_points(flow, Kinv, 0, np.stack([xs.reshape(-1), ys.reshape(-1)], 1))


# ---------------------------------------------------------------------------------------------
# the geometric chain of SFMnet's eval forward in miniature (configs[4] without the networks):
# flow -> correspondences -> pose -> plane sweep -> depth.  Checks that P = [R|t] "from ref to
# target" is exactly what the plane-sweep warp expects (SURVEY 8(a) conventions).
# ---------------------------------------------------------------------------------------------
def test_flow_pose_sweep_chain_recovers_scene_depth(engine):
    H, W = 96, 160
    K = np.array([[140.0, 0, 80.0], [0, 140.0, 48.0], [0, 0, 1.0]])
    Kinv = np.linalg.inv(K)
    R = synth.rodrigues((0.01, -0.03, 0.005))
    t = np.array([0.9, 0.1, -0.42]); t /= np.linalg.norm(t)          # unit baseline: depth scale fixed
    # scene: a valley of two planes n.X = d (camera-1 coordinates), the solid being everything beyond
    # either plane, so from any camera in the free space the visible depth along a ray is the SMALLER
    # of the two plane depths (a single plane would leave the classical two-fold planar ambiguity)
    planes = [(np.array([0.40, 0.05, 1.0]), 12.0), (np.array([-0.45, -0.10, 1.0]), 12.0)]
    planes = [(n / np.linalg.norm(n), d / np.linalg.norm(n)) for n, d in planes]
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    pix = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1)
    ray = Kinv @ pix
    D1 = np.min([d / (n @ ray) for n, d in planes], axis=0)           # visible depth in camera 1
    X1 = ray * D1
    X2 = R @ X1 + t[:, None]
    p2 = K @ (X2 / X2[2:3])
    flow = np.stack([(p2[0] - xs.reshape(-1)).reshape(H, W), (p2[1] - ys.reshape(-1)).reshape(H, W)]).astype(np.float32)

    def texture(X):                                                   # smooth function of the 3-D point, 4 channels
        f = np.array([[0.9, 0.3, 0.2], [-0.4, 1.1, 0.1], [0.6, -0.7, 0.3], [1.3, 0.5, -0.2]])
        return np.sin(f @ X * 1.7 + np.arange(4)[:, None])

    planes2 = [(R @ n, d + (R @ n) @ t) for n, d in planes]          # the same planes in camera-2 coordinates
    D2 = np.min([d / (n @ ray) for n, d in planes2], axis=0)
    X_of_2 = R.T @ (ray * D2 - t[:, None])                            # camera-1 point seen by each pixel of image 2
    ref_fea = texture(X1).reshape(1, 4, H, W).astype(np.float32)
    tgt_fea = texture(X_of_2).reshape(1, 4, H, W).astype(np.float32)

    tf = dev(flow[None], torch.float32); tK = dev(K[None], torch.float32); tKi = dev(Kinv[None], torch.float32)
    P32, E32, r = engine.pose_from_flow(tf, tKi, 4, THR, margin=4)
    Pn = r.P[0].cpu().numpy()
    assert synth.rotation_error_deg(Pn[:, :3], R) < 0.05 and synth.translation_error_deg(Pn[:, 3], t) < 0.5
    L = 64
    cost = engine.plane_sweep(dev(ref_fea, torch.float32), dev(tgt_fea, torch.float32), P32, tK, tKi, L, 1.0)
    diff = (cost[0, :4] - cost[0, 4:]).abs().sum(0)                  # [L, H, W] photometric cost per plane
    valid = (cost[0, 4:].abs().sum(0) > 0)                            # planes whose warp lands inside the image
    diff = torch.where(valid, diff, torch.full_like(diff, 1e9))
    best = diff.argmin(0).cpu().numpy()                               # winning plane per pixel
    depth = 1.0 * L / (best + 1.0)
    gt = D1.reshape(H, W)
    inner = (slice(8, H - 8), slice(8, W - 8))
    # plane i sits at depth L/(i+1): the quantisation step around depth z is z^2/L
    step = gt[inner] ** 2 / L
    err = np.abs(depth[inner] - gt[inner])
    assert np.median(err / step) < 0.6 and (err < 1.5 * step).mean() > 0.9


def test_cuda_graph_replay_equals_plain_launches(engine, std_pair):
    """Single-pair submissions replay a captured CUDA graph: same results as plain launches, for
    changing input/output pointers, after the workspace has been re-allocated by a bigger batch, and
    for several shapes in the cache."""
    sc, x1, x2 = std_pair
    sets = dev(synth.make_sets(10000, 4096, 31), torch.int32)
    try:
        engine.set_graphs(False)
        plain = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
        plain_small = engine.compute_pose(x1[:777].contiguous(), x2[:777].contiguous(), 2, THR)
    finally:
        engine.set_graphs(True)
    for _ in range(3):                                                           # third occurrence: capture
        a = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    b = engine.compute_pose(x1.clone(), x2.clone(), 8, THR, sets=sets.clone(), want_mask=True)   # replay, new pointers
    small = engine.compute_pose(x1[:777].contiguous(), x2[:777].contiguous(), 2, THR)            # second shape
    # a large batch at 16 iterations re-allocates the workspace: cached graphs must be dropped
    big = [synth.make_pair(3000, seed=900 + i) for i in range(24)]
    engine.compute_pose_batch(dev(np.concatenate([p["x1"] for p in big])), dev(np.concatenate([p["x2"] for p in big])),
                              np.arange(25) * 3000, 16, THR)
    c = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)           # plain again, graphs dropped
    for r in (a, b, c):
        assert torch.equal(r.E, plain.E) and torch.equal(r.P, plain.P) and torch.equal(r.mask, plain.mask)
        assert torch.equal(r.stats[:5], plain.stats[:5])
    assert torch.equal(small.E, plain_small.E) and small.count == plain_small.count


def test_graphs_are_shared_by_a_bucket_of_point_counts(engine, std_pair):
    """SFMnet's keypoint path calls computeP with a different N on every pair.  Graphs are keyed on a
    bucket of N (the kernels read the true N from the descriptor; only grids and the tile length
    come from the bucket): calls with N anywhere inside a bucket, odd and even, at both its ends,
    replay one graph and must equal plain launches bit for bit — reference-RNG table included,
    whose scaling depends on N."""
    sc, x1, x2 = std_pair
    ns = [3585, 3600, 3999, 4095, 4096, 3777, 3586, 4001]       # one bucket: (3584, 4096]
    views = [(x1[:n].contiguous(), x2[:n].contiguous()) for n in ns]
    try:
        engine.set_graphs(False)
        plain = [engine.compute_pose(a, b, 5, THR, want_mask=True) for a, b in views]
        plain_h = [engine.compute_pose(a, b, 2, THR, sets=dev(synth.make_sets(a.shape[0], 1024, 5), torch.int32), want_mask=True)
                   for a, b in views]
    finally:
        engine.set_graphs(True)
    for rep in range(2):                                         # the third call of the bucket captures, the rest replay
        got = [engine.compute_pose(a, b, 5, THR, want_mask=True) for a, b in views]
        got_h = [engine.compute_pose(a, b, 2, THR, sets=dev(synth.make_sets(a.shape[0], 1024, 5), torch.int32), want_mask=True)
                 for a, b in views]
        for g, p in zip(got + got_h, plain + plain_h):
            assert torch.equal(g.E, p.E) and torch.equal(g.P, p.P) and torch.equal(g.mask, p.mask)
            assert torch.equal(g.stats[:5], p.stats[:5])


def test_default_minimal_sets_equal_the_reference_rng_table_for_any_n_and_iters(engine):
    """sets = NULL draws the reference's table without allocating or synchronising per call: the uniform
    draws are cached once per context (grown when a call asks for more iterations) and scaled by N on
    the stream.  Must equal tv5_ref_rng_sets (bit-identical to the reference's curand table,
    test_reference_rng_table) for every N, in any order of N and iteration counts."""
    sc = synth.make_pair(20000, seed=12)
    X1, X2 = dev(sc["x1"]), dev(sc["x2"])
    for n, iters in ((777, 1), (10000, 8), (5, 3), (20000, 40), (1, 2), (4097, 8), (10000, 2)):
        a, b = X1[:n].contiguous(), X2[:n].contiguous()
        r = engine.compute_pose(a, b, iters, THR, want_mask=True)
        r2 = engine.compute_pose(a, b, iters, THR, sets=engine.ref_rng_sets(n, iters), want_mask=True)
        assert torch.equal(r.E, r2.E) and torch.equal(r.P, r2.P) and torch.equal(r.mask, r2.mask)
        assert torch.equal(r.stats[:5], r2.stats[:5])
    # batch: every pair scales the same draws by its own N
    ns = [3000, 777, 4096]
    off = np.r_[0, np.cumsum(ns)]
    xb1 = torch.cat([X1[:n] for n in ns]).contiguous()
    xb2 = torch.cat([X2[:n] for n in ns]).contiguous()
    rb = engine.compute_pose_batch(xb1, xb2, off, 4, THR)
    for i, n in enumerate(ns):
        rs = engine.compute_pose(X1[:n].contiguous(), X2[:n].contiguous(), 4, THR, sets=engine.ref_rng_sets(n, 4))
        assert torch.equal(rb.E[i], rs.E) and int(rb.count[i]) == rs.count and int(rb.best_set[i]) == rs.best_set


def test_submissions_on_two_streams_of_one_context_do_not_race(engine, std_pair):
    """One context = one workspace.  Calls issued on different streams are ordered by an event inside
    the library (submission_enter / submission_leave), so alternating streams gives the serial results."""
    sc, x1, x2 = std_pair
    a1, a2 = x1[:6000].contiguous(), x2[:6000].contiguous()
    b1, b2 = x1[4000:].contiguous(), x2[4000:].contiguous()
    sa = dev(synth.make_sets(6000, 2048, 41), torch.int32)
    sb = dev(synth.make_sets(6000, 2048, 42), torch.int32)
    ra = engine.compute_pose(a1, a2, 4, THR, sets=sa, want_mask=True)
    rb = engine.compute_pose(b1, b2, 4, THR, sets=sb, want_mask=True)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(12):
        with torch.cuda.stream(s1):
            outs.append((engine.compute_pose(a1, a2, 4, THR, sets=sa, want_mask=True), ra))
        with torch.cuda.stream(s2):
            outs.append((engine.compute_pose(b1, b2, 4, THR, sets=sb, want_mask=True), rb))
    torch.cuda.synchronize()
    for got, want in outs:
        assert torch.equal(got.E, want.E) and torch.equal(got.P, want.P) and torch.equal(got.mask, want.mask)
        assert torch.equal(got.stats[:5], want.stats[:5])


def test_two_host_threads_sharing_one_context_are_serialised(engine, std_pair):
    """ctypes releases the GIL during a call, so two Python threads can be inside libtv5 at once.  Every
    entry point holds the context's mutex: the calls are serialised, results equal the serial ones."""
    import threading
    sc, x1, x2 = std_pair
    jobs = []
    for k in range(2):
        a = x1[k * 3000:k * 3000 + 5000].contiguous()
        b = x2[k * 3000:k * 3000 + 5000].contiguous()
        st = dev(synth.make_sets(5000, 1024, 90 + k), torch.int32)
        jobs.append((a, b, st, engine.compute_pose(a, b, 2, THR, sets=st, want_mask=True)))
    torch.cuda.synchronize()
    out = [[], []]
    errs = []

    def work(k):
        try:
            a, b, st, _ = jobs[k]
            for _ in range(40):
                out[k].append(engine.compute_pose(a, b, 2, THR, sets=st, want_mask=True))
        except Exception as ex:   # noqa: BLE001
            errs.append(ex)
    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    torch.cuda.synchronize()
    assert not errs, errs
    for k in range(2):
        want = jobs[k][3]
        assert len(out[k]) == 40
        for got in out[k]:
            assert torch.equal(got.E, want.E) and torch.equal(got.P, want.P) and torch.equal(got.mask, want.mask)
            assert torch.equal(got.stats[:5], want.stats[:5])


def test_early_exit_single_large_pair_with_graph_replay(engine):
    """One large pair is staged too (enough work) and, as a single-pair submission, goes through the
    CUDA-graph replay: four identical calls with early exit == the full scoring."""
    sc = synth.make_pair(120000, seed=77)
    x1, x2 = dev(sc["x1"]), dev(sc["x2"])
    sets = dev(synth.make_sets(120000, 4096, 78), torch.int32)
    full = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)
    try:
        engine.set_early_exit(True)
        runs = [engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True) for _ in range(4)]
    finally:
        engine.set_early_exit(False)
    for r in runs:
        assert torch.equal(r.E, full.E) and torch.equal(r.P, full.P) and torch.equal(r.mask, full.mask)
        assert torch.equal(r.stats[:4], full.stats[:4])
    again = engine.compute_pose(x1, x2, 8, THR, sets=sets, want_mask=True)   # back to full scoring: other graph key
    assert torch.equal(again.E, full.E) and torch.equal(again.stats[:5], full.stats[:5])
