"""Loading of libtv5.so through ctypes, with the prototypes of include/tv5.h."""
import ctypes as C
import os
import re
import subprocess

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPO = os.path.dirname(_PKG)
_SO = os.environ.get("TV5_LIB") or os.path.join(_PKG, "libtv5.so")  # TV5_LIB: dev builds only
_HEADER = os.path.join(_REPO, "include", "tv5.h")

TV5_N_STAGES = 6
STAGE_NAMES = ("prep", "solve", "plan", "score_bounds", "candidates_exact", "finalize")


class Tv5Error(RuntimeError):
    pass


class Tv5Result(C.Structure):
    _fields_ = [("count", C.c_int32), ("best_set", C.c_int32), ("best_root", C.c_int32),
                ("n_hypotheses", C.c_int32), ("n_candidates", C.c_int32),
                ("fast_path", C.c_int32), ("reserved", C.c_int32 * 2)]


def lib_path():
    return _SO


def build_library(force=False, verbose=False):
    """Compile csrc/ into libtv5.so for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_PKG, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir)
            if f.endswith((".cu", ".cuh", ".h"))] + [_HEADER]
    if not force and os.path.exists(_SO) and all(
            os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    cmd = ["make", "-C", src_dir] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise Tv5Error("building libtv5.so failed")
    return _SO


def exported_symbols():
    """Names declared in include/tv5.h (functions beginning with tv5_)."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tv5_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load_library():
    """dlopen libtv5.so and attach prototypes.  Raises Tv5Error if it is missing: there is no
    fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise Tv5Error(f"{_SO} not found: build it with `make -C {os.path.join(_PKG, 'csrc')}` "
                       "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(_SO)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    L.tv5_create.restype = C.c_int
    L.tv5_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.tv5_destroy.restype = C.c_int
    L.tv5_destroy.argtypes = [vp]
    L.tv5_strerror.restype = C.c_char_p
    L.tv5_strerror.argtypes = [C.c_int]
    L.tv5_last_cuda_error.restype = C.c_int
    L.tv5_last_cuda_error.argtypes = [vp]
    L.tv5_version.restype = C.c_int
    L.tv5_version.argtypes = []
    L.tv5_device_sm_count.restype = C.c_int
    L.tv5_device_sm_count.argtypes = [vp]
    L.tv5_compute_pose.restype = C.c_int
    L.tv5_compute_pose.argtypes = [vp, vp, dp, dp, C.c_int, ip, C.c_int, C.c_int, C.c_int,
                                   C.c_double, C.c_int, dp, dp, vp, vp]
    L.tv5_compute_pose_batch.restype = C.c_int
    L.tv5_compute_pose_batch.argtypes = [vp, vp, C.c_int, dp, dp, C.POINTER(C.c_int64), ip,
                                         C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp, dp,
                                         vp, vp]
    L.tv5_compute_pose_batch_host.restype = C.c_int
    L.tv5_compute_pose_batch_host.argtypes = [vp, vp, C.c_int, dp, dp, C.POINTER(C.c_int64), ip,
                                              C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, dp,
                                              dp, vp]
    L.tv5_solve5.restype = C.c_int
    L.tv5_solve5.argtypes = [vp, vp, dp, dp, C.c_int, ip, C.c_int, C.c_int, dp, dp, ip, ip]
    L.tv5_score.restype = C.c_int
    L.tv5_score.argtypes = [vp, vp, dp, dp, C.c_int, dp, C.c_int, C.c_double, ip, vp]
    L.tv5_score_bounds.restype = C.c_int
    L.tv5_score_bounds.argtypes = [vp, vp, dp, dp, C.c_int, dp, C.c_int, C.c_double, ip, ip]
    L.tv5_ref_rng_sets.restype = C.c_int
    L.tv5_ref_rng_sets.argtypes = [vp, vp, C.c_int, C.c_int, ip]
    L.tv5_decompose.restype = C.c_int
    L.tv5_decompose.argtypes = [dp, dp]
    L.tv5_decompose_uv.restype = C.c_int
    L.tv5_decompose_uv.argtypes = [dp, dp, dp]
    L.tv5_decompose_batch.restype = C.c_int
    L.tv5_decompose_batch.argtypes = [vp, vp, dp, C.c_int, dp, dp, dp]
    L.tv5_optimise.restype = C.c_int
    L.tv5_optimise.argtypes = [vp, vp, dp, dp, C.c_int, vp, dp, C.c_double, C.c_double, C.c_int, ip]
    L.tv5_optimise_batch.restype = C.c_int
    L.tv5_optimise_batch.argtypes = [vp, vp, C.c_int, dp, dp, C.POINTER(C.c_int64), vp, dp,
                                     C.c_double, C.c_double, C.c_int, ip]
    L.tv5_optimise_host.restype = C.c_int
    L.tv5_optimise_host.argtypes = [vp, vp, dp, dp, C.c_int, dp, C.c_double, C.c_double, C.c_int, dp]
    L.tv5_flow_to_points.restype = C.c_int
    L.tv5_flow_to_points.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp,
                                     C.POINTER(C.c_int64), dp, dp]
    L.tv5_pose_from_flow.restype = C.c_int
    L.tv5_pose_from_flow.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp,
                                     C.POINTER(C.c_int64), ip, C.c_int, C.c_double, C.c_int, vp, vp,
                                     vp, dp, dp]
    L.tv5_plane_sweep.restype = C.c_int
    L.tv5_plane_sweep.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_int, vp]
    L.tv5_winner_record.restype = C.c_int
    L.tv5_winner_record.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.tv5_winner_pick.restype = C.c_int
    L.tv5_winner_pick.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp]
    L.tv5_debug_guard.restype = C.c_int
    L.tv5_debug_guard.argtypes = [vp, C.c_int, C.c_int]
    L.tv5_debug_poison.restype = C.c_int
    L.tv5_debug_poison.argtypes = [vp, C.c_int]
    L.tv5_debug_stray_write.restype = C.c_int
    L.tv5_debug_stray_write.argtypes = [vp, C.c_int]
    L.tv5_debug_check_guards.restype = C.c_int
    L.tv5_debug_check_guards.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.tv5_measure_fp32_peak.restype = C.c_int
    L.tv5_measure_fp32_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.tv5_set_force_exact.restype = C.c_int
    L.tv5_set_force_exact.argtypes = [vp, C.c_int]
    L.tv5_set_graphs.restype = C.c_int
    L.tv5_set_graphs.argtypes = [vp, C.c_int]
    L.tv5_set_early_exit.restype = C.c_int
    L.tv5_set_early_exit.argtypes = [vp, C.c_int]
    L.tv5_set_split_solver.restype = C.c_int
    L.tv5_set_split_solver.argtypes = [vp, C.c_int]
    L.tv5_set_overlap.restype = C.c_int
    L.tv5_set_overlap.argtypes = [vp, C.c_int]
    L.tv5_profile_enable.restype = C.c_int
    L.tv5_profile_enable.argtypes = [vp, C.c_int]
    L.tv5_profile_read.restype = C.c_int
    L.tv5_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]
    _lib = L
    return L
