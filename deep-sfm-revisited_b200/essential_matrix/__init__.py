"""Drop-in replacement for the reference's `essential_matrix` extension module.

The reference builds this name from RANSAC_FiveP (setup.py:5-18) and exposes five functions
(essential_matrix_wrapper.cpp:102-108).  `models/SFMnet.py:11` and `epipolar_utils.py:4` import
it by this name, so putting this package's parent directory on sys.path swaps the pose stage
without touching the reference's Python.

  computeP(x1, x2, num_test_points, num_ransac_test_points, num_ransac_iterations, thr)
      -> (E f64 CUDA [3,3], P f64 CUDA [3,4], n_inliers int-like)
  initialise(same six arguments) -> E f64 CUDA [3,3]
  optimise / decompose / decomposeUV: host-side functions of the reference (polish_E.cu); not on
      the accelerated path and not provided yet — they raise
      NotImplementedError rather than silently computing something else.

`n_inliers` is a LazyCount: it behaves like the Python int the reference returns, but reads the
device counter (one stream sync) only when its value is actually used.  SFMnet ignores it
(models/SFMnet.py:267), so the pose stage stays asynchronous.
"""
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


def computeP(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
             inlier_threshold):
    """essential_matrix.computeP — ProjectionMatrixRansacWrapper, wrapper.cpp:59-71."""
    r = _tv5.get_engine(getattr(input1, "device", None) if getattr(input1, "is_cuda", False) else None) \
        .compute_pose(input1, input2, int(num_ransac_iterations), float(inlier_threshold),
                      n_pre=int(num_test_points), n_full=int(num_ransac_test_points),
                      with_cheirality=True)
    return r.E, r.P, LazyCount(r)


def initialise(input1, input2, num_test_points, num_ransac_test_points, num_ransac_iterations,
               inlier_threshold):
    """essential_matrix.initialise — EssentialMatrixInitialiseWrapper, wrapper.cpp:45-57.
    (The reference also prints the inlier count on every call, essential_matrix.cu:170; this
    implementation never prints.)"""
    r = _tv5.get_engine(getattr(input1, "device", None) if getattr(input1, "is_cuda", False) else None) \
        .compute_pose(input1, input2, int(num_ransac_iterations), float(inlier_threshold),
                      n_pre=int(num_test_points), n_full=int(num_ransac_test_points),
                      with_cheirality=False)
    return r.E


def optimise(input1, input2, E_init, delta, alpha, MaxReps):
    raise NotImplementedError("essential_matrix.optimise (host IRLS refinement, polish_E.cu:1470-1577) "
                              "is outside the accelerated path; see DESIGN.md 'out of scope'")


def decompose(Emat):
    raise NotImplementedError("essential_matrix.decompose (polish_E.cu:147-338) is outside the "
                              "accelerated path; see DESIGN.md 'out of scope'")


def decomposeUV(Emat):
    raise NotImplementedError("essential_matrix.decomposeUV (polish_E.cu:147-244) is outside the "
                              "accelerated path; see DESIGN.md 'out of scope'")
