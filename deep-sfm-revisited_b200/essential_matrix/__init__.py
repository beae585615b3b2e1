"""Drop-in replacement for the reference's `essential_matrix` extension module.

The reference builds this name from RANSAC_FiveP (setup.py:5-18) and exposes five functions
(essential_matrix_wrapper.cpp:102-108).  `models/SFMnet.py:11` and `epipolar_utils.py:4` import
it by this name, so putting this package's parent directory on sys.path swaps the pose stage
without touching the reference's Python.

  computeP(x1, x2, num_test_points, num_ransac_test_points, num_ransac_iterations, thr)
      -> (E f64 CUDA [3,3], P f64 CUDA [3,4], n_inliers int-like)
  initialise(same six arguments) -> E f64 CUDA [3,3]
  optimise(x1, x2, E_init, delta, alpha, MaxReps) -> E f64 [3,3]   (GPU refinement, tv5_optimise)
  decompose(E) -> f64 [5] Givens angles;  decomposeUV(E) -> (U, V) f64 [3,3]

`n_inliers` is a LazyCount: it behaves like the Python int the reference returns, but reads the
device counter (one stream sync) only when its value is actually used.  SFMnet ignores it
(models/SFMnet.py:267), so the pose stage stays asynchronous.
"""
import torch as _torch
import tv5 as _tv5


class LazyCount:
    """int-like view of the device-side best inlier count."""

    def __init__(self, result):
        self._r = result
        self._v = None

    def _get(self):
        if self._v is None:
            self._v = self._r.count
        return self._v

    def __int__(self):
        return self._get()

    __index__ = __int__

    def __repr__(self):
        return str(self._get())

    def __eq__(self, o):
        return self._get() == int(o)

    def __lt__(self, o):
        return self._get() < int(o)

    def __le__(self, o):
        return self._get() <= int(o)

    def __gt__(self, o):
        return self._get() > int(o)

    def __ge__(self, o):
        return self._get() >= int(o)

    def __hash__(self):
        return hash(self._get())

    def __add__(self, o):
        return self._get() + o

    __radd__ = __add__

    def __sub__(self, o):
        return self._get() - o

    def __rsub__(self, o):
        return o - self._get()

    def __mul__(self, o):
        return self._get() * o

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._get() / o

    def __rtruediv__(self, o):
        return o / self._get()

    def __float__(self):
        return float(self._get())

    def __bool__(self):
        return self._get() != 0

    def __format__(self, spec):
        return format(self._get(), spec)


def _check_input_init(x, name):
    # CHECK_INPUT_INIT, essential_matrix_wrapper.cpp:39-42 (same order, same wording)
    if not isinstance(x, _torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if x.dtype != _torch.float64:
        raise RuntimeError(f"{name} must be a double tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def _ransac(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
            inlier_threshold, with_cheirality):
    _check_input_init(input1, "input1")
    _check_input_init(input2, "input2")
    eng = _tv5.get_engine(input1.device)
    iters, thr = int(num_ransac_iterations), float(inlier_threshold)
    if input1.dim() == 2 and input1.shape[1] == 2:
        return eng.compute_pose(input1, input2, iters, thr, n_pre=int(num_test_points),
                                n_full=int(num_ransac_test_points), with_cheirality=with_cheirality)
    # Any other shape: the reference takes num_points = input1.size(0) for SAMPLING
    # (essential_matrix.cu:118,199) and reads both tensors as flat (x, y) arrays, scoring the first
    # num_test_points / num_ransac_test_points of them.  epipolar_utils.compute_E_matrix calls it
    # this way with [1, n, 2] tensors (epipolar_utils.py:70-73), i.e. every minimal set is five
    # times point 0.  Same rule here: minimal sets from the reference RNG table of size(0) points,
    # scoring domain clipped to the points actually present (the reference would read past the end).
    if input1.numel() % 2 or input1.numel() < 2 or input2.numel() != input1.numel():
        raise RuntimeError("input1 and input2 must hold the same number of (x, y) rows")
    rows = max(int(input1.shape[0]), 1) if input1.dim() else 1
    flat1, flat2 = input1.view(-1, 2), input2.view(-1, 2)
    sets = eng.ref_rng_sets(rows, iters)
    return eng.compute_pose(flat1, flat2, iters, thr, n_pre=int(num_test_points),
                            n_full=int(num_ransac_test_points), sets=sets, with_cheirality=with_cheirality)


def computeP(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
             inlier_threshold):
    """essential_matrix.computeP — ProjectionMatrixRansacWrapper, wrapper.cpp:59-71."""
    r = _ransac(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
                inlier_threshold, True)
    return r.E, r.P, LazyCount(r)


def initialise(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
               inlier_threshold):
    """essential_matrix.initialise — EssentialMatrixInitialiseWrapper, wrapper.cpp:45-57.
    (The reference also prints the inlier count on every call, essential_matrix.cu:170; this
    implementation never prints.)"""
    return _ransac(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
                   inlier_threshold, False).E


def _require_double_contiguous(x, name):
    # CHECK_INPUT_OPT, essential_matrix_wrapper.cpp:43
    if x.dtype != _torch.float64:
        raise RuntimeError(f"{name} must be a double tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def optimise(input1, input2, E_init, delta, alpha, MaxReps):
    """essential_matrix.optimise — EssentialMatrixOptimiseWrapper, wrapper.cpp:73-87 ->
    polish_E_robust_parametric (polish_E.cu:1470-1577), which the reference runs on one CPU core
    over CPU tensors.  Here the refinement runs on the GPU (tv5_optimise): CPU tensors (the
    reference's calling convention, epipolar_utils.py:76) are copied to the current CUDA device
    and the result comes back on E_init's device; CUDA tensors are used in place.
    As in the reference the number of points is input1.size(0) (essential_matrix.cu:86)."""
    _require_double_contiguous(input1, "input1")
    _require_double_contiguous(input2, "input2")
    _require_double_contiguous(E_init, "E_init")
    n = int(input1.shape[0])
    if input1.is_cuda:
        eng = _tv5.get_engine(input1.device)
        x1 = input1.reshape(-1, 2)[:n]
        x2 = input2.to(input1.device).reshape(-1, 2)[:n]
        E = eng.optimise(x1, x2, E_init, float(delta), float(alpha), int(MaxReps))
        return E.to(E_init.device)
    eng = _tv5.get_engine()
    E = eng.optimise_host(input1.reshape(-1, 2)[:n].numpy(), input2.reshape(-1, 2)[:n].numpy(),
                          E_init.cpu().numpy(), float(delta), float(alpha), int(MaxReps))
    return _torch.from_numpy(E).to(E_init.device)


def decompose(Emat):
    """essential_matrix.decompose — EssentialMatrixDecompose (essential_matrix.cu:29-44 ->
    Edecomp, polish_E.cu:246-338): the five Givens angles (x, y, z, u, v), float64 [5] on Emat's
    device.  A CPU tensor is decomposed on the calling thread by libtv5 (as the reference does); a
    CUDA tensor — a host segfault in the reference — by the device kernel."""
    if Emat.dtype != _torch.float64:
        raise RuntimeError("expected scalar type Double but found " + str(Emat.dtype))
    if Emat.is_cuda:
        return _tv5.get_engine(Emat.device).decompose_batch(Emat.reshape(1, 3, 3), want_uv=False)["angles"][0]
    return _torch.from_numpy(_tv5.decompose_host(Emat.contiguous().numpy()))


def decomposeUV(Emat):
    """essential_matrix.decomposeUV — EssentialMatrixDecomposeUV (essential_matrix.cu:49-70):
    (U, V) float64 [3,3] with E ~ U diag(1,1,0) V^T."""
    if Emat.dtype != _torch.float64:
        raise RuntimeError("expected scalar type Double but found " + str(Emat.dtype))
    if Emat.is_cuda:
        r = _tv5.get_engine(Emat.device).decompose_batch(Emat.reshape(1, 3, 3), want_angles=False)
        return r["U"][0], r["V"][0]
    U, V = _tv5.decompose_uv_host(Emat.contiguous().numpy())
    return _torch.from_numpy(U), _torch.from_numpy(V)
