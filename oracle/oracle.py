"""ctypes bindings for the CPU oracle (libtv5_oracle.so) and, when built, for the compiled
copies of the reference under oracle/_ref/ (see build_ref.sh).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def build(force=False):
    so = os.path.join(_HERE, "libtv5_oracle.so")
    src = os.path.join(_HERE, "tv5_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libtv5_oracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.tv5o_sampson_err.restype = C.c_double
        L.tv5o_sampson_err.argtypes = [_dp, C.c_double, C.c_double, C.c_double, C.c_double]
        L.tv5o_score.restype = None
        L.tv5o_score.argtypes = [_dp, _dp, C.c_int, _dp, C.c_int, C.c_double, _ip, _bp]
        L.tv5o_solve5.restype = C.c_int
        L.tv5o_solve5.argtypes = [_dp, _dp, _dp, _dp]
        L.tv5o_nullspace_basis.restype = None
        L.tv5o_nullspace_basis.argtypes = [_dp, _dp, _dp]
        L.tv5o_hidden_poly.restype = None
        L.tv5o_hidden_poly.argtypes = [_dp, _dp, _dp]
        L.tv5o_real_roots.restype = C.c_int
        L.tv5o_real_roots.argtypes = [_dp, C.c_int, _dp]
        L.tv5o_cheirality.restype = C.c_int
        L.tv5o_cheirality.argtypes = [_dp, _dp, _dp, C.c_int, _dp]
        L.tv5o_index_from_uniform.restype = C.c_int32
        L.tv5o_index_from_uniform.argtypes = [C.c_float, C.c_int]
        L.tv5o_ransac.restype = C.c_int
        L.tv5o_ransac.argtypes = [_dp, _dp, C.c_int, _ip, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_int, _dp, _dp, _ip, _ip, _bp]
        L.tv5o_solve_sets.restype = None
        L.tv5o_solve_sets.argtypes = [_dp, _dp, _ip, C.c_int, C.c_int, _dp, _dp, _ip, _ip]
        L.tv5o_decompose_uv.restype = None
        L.tv5o_decompose_uv.argtypes = [_dp, _dp, _dp]
        L.tv5o_decompose_angles.restype = None
        L.tv5o_decompose_angles.argtypes = [_dp, _dp]
        L.tv5o_optimise.restype = None
        L.tv5o_optimise.argtypes = [_dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def sampson_err(E, x1, y1, x2, y2):
    E = _f64(E).reshape(9)
    return lib().tv5o_sampson_err(_p(E, _dp), float(x1), float(y1), float(x2), float(y2))


def score(x1, x2, E_list, thr, n=None, want_mask=False):
    """counts[M] (and mask[M,n]) of the reference Sampson test on the first n points."""
    x1, x2 = _f64(x1), _f64(x2)
    E_list = _f64(E_list).reshape(-1, 9)
    n = x1.shape[0] if n is None else int(n)
    M = E_list.shape[0]
    counts = np.zeros(M, np.int32)
    mask = np.zeros((M, n), np.uint8) if want_mask else None
    lib().tv5o_score(_p(x1, _dp), _p(x2, _dp), n, _p(E_list, _dp), M, float(thr),
                     _p(counts, _ip), _p(mask, _bp))
    return (counts, mask) if want_mask else counts


def solve5(q, qp):
    """(E[n,3,3], w[n]) for one minimal set; q, qp are 5x2."""
    q, qp = _f64(q).reshape(5, 2), _f64(qp).reshape(5, 2)
    E = np.zeros((10, 9))
    w = np.zeros(10)
    n = lib().tv5o_solve5(_p(q, _dp), _p(qp, _dp), _p(E, _dp), _p(w, _dp))
    return E[:n].reshape(n, 3, 3).copy(), w[:n].copy()


def nullspace_basis(q, qp):
    q, qp = _f64(q).reshape(5, 2), _f64(qp).reshape(5, 2)
    B = np.zeros((4, 9))
    lib().tv5o_nullspace_basis(_p(q, _dp), _p(qp, _dp), _p(B, _dp))
    return B


def hidden_poly(q, qp):
    q, qp = _f64(q).reshape(5, 2), _f64(qp).reshape(5, 2)
    p = np.zeros(11)
    lib().tv5o_hidden_poly(_p(q, _dp), _p(qp, _dp), _p(p, _dp))
    return p


def real_roots(p):
    p = _f64(p)
    r = np.zeros(32)
    n = lib().tv5o_real_roots(_p(p, _dp), len(p) - 1, _p(r, _dp))
    return r[:n].copy()


def cheirality(q, qp, E):
    """(E_kept[n,3,3], P[n,3,4]) after the unanimous 5-point vote."""
    q, qp = _f64(q).reshape(5, 2), _f64(qp).reshape(5, 2)
    Eb = np.zeros((10, 9))
    E = _f64(E).reshape(-1, 9)
    Eb[: len(E)] = E
    P = np.zeros((10, 12))
    n = lib().tv5o_cheirality(_p(q, _dp), _p(qp, _dp), _p(Eb, _dp), len(E), _p(P, _dp))
    return Eb[:n].reshape(n, 3, 3).copy(), P[:n].reshape(n, 3, 4).copy()


def index_from_uniform(u, N):
    return lib().tv5o_index_from_uniform(float(np.float32(u)), int(N))


def solve_sets(x1, x2, sets, with_cheirality=True):
    x1, x2, sets = _f64(x1), _f64(x2), _i32(sets).reshape(-1, 5)
    H = sets.shape[0]
    E = np.zeros((H, 10, 9))
    P = np.zeros((H, 10, 12))
    nr = np.zeros(H, np.int32)
    nv = np.zeros(H, np.int32)
    lib().tv5o_solve_sets(_p(x1, _dp), _p(x2, _dp), _p(sets, _ip), H, int(with_cheirality),
                          _p(E, _dp), _p(P, _dp), _p(nr, _ip), _p(nv, _ip))
    return dict(E=E, P=P, n_roots=nr, n_valid=nv)


def ransac(x1, x2, sets, iters, thr, n_pre=None, n_full=None, with_cheirality=True,
           want_mask=False):
    """Reference-order RANSAC over sets[H,5] with H = n_threads*iters."""
    x1, x2, sets = _f64(x1), _f64(x2), _i32(sets).reshape(-1, 5)
    N = x1.shape[0]
    H = sets.shape[0]
    assert H % iters == 0
    n_pre = N if n_pre is None else int(n_pre)
    n_full = N if n_full is None else int(n_full)
    E = np.zeros(9)
    P = np.zeros(12)
    bs = np.zeros(1, np.int32)
    br = np.zeros(1, np.int32)
    mask = np.zeros(n_full, np.uint8) if want_mask else None
    cnt = lib().tv5o_ransac(_p(x1, _dp), _p(x2, _dp), N, _p(sets, _ip), H // iters, int(iters),
                            n_pre, n_full, float(thr), int(with_cheirality), _p(E, _dp),
                            _p(P, _dp), _p(bs, _ip), _p(br, _ip), _p(mask, _bp))
    out = dict(E=E.reshape(3, 3), P=P.reshape(3, 4), count=int(cnt), best_set=int(bs[0]),
               best_root=int(br[0]))
    if want_mask:
        out["mask"] = mask
    return out


def decompose_uv(E):
    Ew = _f64(E).reshape(9).copy()
    U = np.zeros(9); V = np.zeros(9)
    lib().tv5o_decompose_uv(_p(Ew, _dp), _p(U, _dp), _p(V, _dp))
    return U.reshape(3, 3), V.reshape(3, 3)


def decompose_angles(E):
    Ew = _f64(E).reshape(9).copy()
    par = np.zeros(5)
    lib().tv5o_decompose_angles(_p(Ew, _dp), _p(par, _dp))
    return par


def optimise(x1, x2, E, delta, alpha, max_reps):
    x1, x2 = _f64(x1), _f64(x2)
    Ew = _f64(E).reshape(9).copy()
    lib().tv5o_optimise(_p(Ew, _dp), _p(x1, _dp), _p(x2, _dp), x1.shape[0], float(delta), float(alpha), int(max_reps))
    return Ew.reshape(3, 3)


# ---------------------------------------------------------------------------------------------
# Compiled copies of the reference (oracle/_ref/, built by build_ref.sh).  Optional.
# ---------------------------------------------------------------------------------------------
_REF = os.path.join(_HERE, "_ref")
_ref_host = None


def ref_host_available():
    return os.path.exists(os.path.join(_REF, "libref_host.so"))


def ref_host():
    global _ref_host
    if _ref_host is None:
        L = C.CDLL(os.path.join(_REF, "libref_host.so"))
        L.ref_solve_sets.restype = C.c_int
        L.ref_solve_sets.argtypes = [_dp, _dp, C.c_int, _ip, C.c_int, _dp, _ip, _dp, _dp, _ip]
        _ref_host = L
    return _ref_host


def ref_solve_sets(x1, x2, sets):
    """The reference's own compute_E_matrices_optimized + compute_P_matrices, host-compiled."""
    x1, x2, sets = _f64(x1), _f64(x2), _i32(sets).reshape(-1, 5)
    H = sets.shape[0]
    E_all = np.zeros((H, 10, 9))
    E_valid = np.zeros((H, 10, 9))
    P_valid = np.zeros((H, 10, 12))
    nr = np.zeros(H, np.int32)
    nv = np.zeros(H, np.int32)
    rc = ref_host().ref_solve_sets(_p(x1, _dp), _p(x2, _dp), x1.shape[0], _p(sets, _ip), H,
                                   _p(E_all, _dp), _p(nr, _ip), _p(E_valid, _dp),
                                   _p(P_valid, _dp), _p(nv, _ip))
    assert rc == 0
    return dict(E_all=E_all, n_roots=nr, E=E_valid, P=P_valid, n_valid=nv)


# ---------------------------------------------------------------------------------------------
# optical flow -> normalised correspondences (front of pose_by_ransac), numpy float32
# ---------------------------------------------------------------------------------------------
def flow_to_points(flow, Kinv, margin=10, pts=None, cuda_division=True):
    """One image.  flow [2,H,W] float32, Kinv [3,3] float32 -> x1, x2 [n,2] float64.

    Restates, in float32 like the reference, flow2coord (models/SFMnet.py:298-318: pixel grid,
    grid + flow, homogeneous one), the selection of pose_by_ransac (dense crop :240-241 when
    pts is None; integer gather :251-254 for an integer (x, y) list; bilinear grid_sample with
    align_corners=True and zero padding :244-249 for a float list), bmm(K^-1, .) (:259-260),
    the [:, :2] slice (:262-263) and .double() (epipolar_utils.py:130).  The three-term
    products with K^-1 are accumulated k = 0, 1, 2 with fused multiply-adds like an SGEMM inner
    loop (emulated exactly in float64: a float32 product is exact there).
    cuda_division: torch divides a CUDA tensor by a Python scalar as x * (1/s) (the reference always
    runs this chain on the GPU); False gives torch's CPU behaviour x / s."""
    f32 = np.float32
    flow = np.asarray(flow, dtype=f32)
    K = np.asarray(Kinv, dtype=f32)
    _, H, W = flow.shape
    gx = np.broadcast_to(np.arange(W, dtype=f32)[None, :], (H, W))
    gy = np.broadcast_to(np.arange(H, dtype=f32)[:, None], (H, W))
    c1 = np.stack([gx, gy, np.ones((H, W), f32)])                  # [3,H,W]
    c2 = np.stack([gx + flow[0], gy + flow[1], np.ones((H, W), f32)])
    if pts is None:
        a = c1[:, margin:H - margin, margin:W - margin].reshape(3, -1)
        b = c2[:, margin:H - margin, margin:W - margin].reshape(3, -1)
    elif np.issubdtype(np.asarray(pts).dtype, np.integer):
        p = np.asarray(pts)
        a, b = c1[:, p[:, 1], p[:, 0]], c2[:, p[:, 1], p[:, 0]]
    else:
        p = np.asarray(pts, dtype=f32)
        if cuda_division:
            gxn = (f32(2.0) * p[:, 0]) * (f32(1.0) / f32(max(W - 1, 1))) - f32(1.0)
            gyn = (f32(2.0) * p[:, 1]) * (f32(1.0) / f32(max(H - 1, 1))) - f32(1.0)
        else:
            gxn = f32(2.0) * p[:, 0] / f32(max(W - 1, 1)) - f32(1.0)
            gyn = f32(2.0) * p[:, 1] / f32(max(H - 1, 1)) - f32(1.0)
        ix = ((gxn + f32(1.0)) / f32(2.0)) * f32(W - 1)
        iy = ((gyn + f32(1.0)) / f32(2.0)) * f32(H - 1)
        x0, y0 = np.floor(ix), np.floor(iy)
        wx1, wx0, wy1, wy0 = ix - x0, (x0 + f32(1.0)) - ix, iy - y0, (y0 + f32(1.0)) - iy
        a = np.zeros((3, p.shape[0]), f32)
        b = np.zeros((3, p.shape[0]), f32)
        for dx, dy, w in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
            xs, ys = x0.astype(np.int64) + dx, y0.astype(np.int64) + dy
            ok = (xs >= 0) & (xs < W) & (ys >= 0) & (ys < H)
            xc, yc = np.clip(xs, 0, W - 1), np.clip(ys, 0, H - 1)
            for ch in range(3):
                for dst, src in ((a, c1), (b, c2)):
                    term = _fma32(src[ch, yc, xc], w, dst[ch])
                    dst[ch] = np.where(ok, term, dst[ch])

    def apply(c):
        out = np.empty((c.shape[1], 2), np.float64)
        for r in range(2):
            acc = (K[r, 0] * c[0]).astype(f32)
            acc = _fma32(K[r, 1], c[1], acc)
            acc = _fma32(K[r, 2], c[2], acc)
            out[:, r] = acc.astype(np.float64)
        return out

    return apply(a), apply(b)


def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64 and the sum of
    it with a float32 needs < 2*53 bits only in far-apart-exponent cases that the final rounding
    to float32 absorbs (double rounding can differ in 1 ulp; the tests allow it)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# plane-sweep cost volume (consumer of P), numpy float32
# ---------------------------------------------------------------------------------------------
def plane_sweep_cost_volume(ref_fea, tgt_fea, pose, K, Kinv, nlabel, mindepth, by_depth=False,
                            cuda_division=True):
    """One batch element.  ref_fea, tgt_fea [C,h,w] float32; pose [3,4], K, Kinv [3,3] float32
    (quarter-resolution intrinsics) -> cost [2C, nlabel, h, w] float32.

    Restates PSNet.forward's label loop (models/PSNet.py:141-157): plane i has depth
    mindepth*nlabel/(i+1) (or (i+1)*mindepth), the target features are inverse-warped onto it
    (models/inverse_warp.py:121-153: cam = (Kinv pix) * depth :31-45; proj = K [R|t]; X/Z
    normalised to [-1,1], out-of-range -> 2 :48-78; bilinear grid_sample, zero padding,
    align_corners=True) and stacked under the reference features.  float32 throughout, 3-term
    products accumulated k = 0,1,2 with fma like an SGEMM inner loop."""
    f32 = np.float32
    ref_fea = np.asarray(ref_fea, f32); tgt_fea = np.asarray(tgt_fea, f32)
    pose = np.asarray(pose, f32); K = np.asarray(K, f32); Kinv = np.asarray(Kinv, f32)
    C, h, w = ref_fea.shape
    gx = np.broadcast_to(np.arange(w, dtype=f32)[None, :], (h, w))
    gy = np.broadcast_to(np.arange(h, dtype=f32)[:, None], (h, w))
    one = np.ones((h, w), f32)

    def mat3(M, v):   # rows of M times the 3-vector field v, sgemm order
        return [_fma32(M[r, 2], v[2], _fma32(M[r, 1], v[1], (M[r, 0] * v[0]).astype(f32))) for r in range(3)]

    ray = mat3(Kinv, [gx, gy, one])
    proj = np.empty((3, 4), f32)
    for r in range(3):
        for c in range(4):
            proj[r, c] = _fma32(K[r, 2], pose[2, c], _fma32(K[r, 1], pose[1, c], f32(K[r, 0] * pose[0, c])))
    cost = np.zeros((2 * C, nlabel, h, w), f32)
    d2d = f32(f32(f32(1.0) * f32(mindepth)) * f32(nlabel))
    for i in range(nlabel):
        if by_depth:
            depth = f32(f32(f32(1.0) * f32(i + 1)) * f32(mindepth))
        elif cuda_division:
            depth = f32(d2d * f32(f32(1.0) / f32(i + 1 + 1e-16)))
        else:
            depth = f32(d2d / f32(i + 1 + 1e-16))
        cam = [(r * depth).astype(f32) for r in ray]
        pc = [(v + proj[r, 3]).astype(f32) for r, v in enumerate(mat3(proj[:, :3], cam))]
        Z = np.maximum(pc[2], f32(1e-3))
        if cuda_division:
            xn = ((f32(2.0) * (pc[0] / Z)).astype(f32) * f32(f32(1.0) / f32(w - 1))).astype(f32) - f32(1.0)
            yn = ((f32(2.0) * (pc[1] / Z)).astype(f32) * f32(f32(1.0) / f32(h - 1))).astype(f32) - f32(1.0)
        else:
            xn = (f32(2.0) * (pc[0] / Z)).astype(f32) / f32(w - 1) - f32(1.0)
            yn = (f32(2.0) * (pc[1] / Z)).astype(f32) / f32(h - 1) - f32(1.0)
        xn = np.where((xn > 1) | (xn < -1), f32(2.0), xn).astype(f32)
        yn = np.where((yn > 1) | (yn < -1), f32(2.0), yn).astype(f32)
        ix = (((xn + f32(1.0)) / f32(2.0)) * f32(w - 1)).astype(f32)
        iy = (((yn + f32(1.0)) / f32(2.0)) * f32(h - 1)).astype(f32)
        x0, y0 = np.floor(ix), np.floor(iy)
        wx1, wx0, wy1, wy0 = ix - x0, (x0 + f32(1.0)) - ix, iy - y0, (y0 + f32(1.0)) - iy
        acc = np.zeros((C, h, w), f32)
        for dx, dy, wt in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
            with np.errstate(invalid="ignore"):
                xs = np.nan_to_num(x0, nan=-10.0).astype(np.int64) + dx
                ys = np.nan_to_num(y0, nan=-10.0).astype(np.int64) + dy
            ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
            xc, yc = np.clip(xs, 0, w - 1), np.clip(ys, 0, h - 1)
            term = _fma32(tgt_fea[:, yc, xc], wt[None].astype(f32), acc)
            acc = np.where(ok[None], term, acc)
        cost[:C, i] = ref_fea
        cost[C:, i] = acc
    return cost
