"""tv5 — host-side Python API of libtv5, the B200 two-view relative-pose engine.

PyTorch is used for device memory and streams only; all computation happens in the hand-written
CUDA kernels of csrc/ behind the C ABI of include/tv5.h.  There is no CPU fallback: importing
works anywhere (the library is loaded lazily), but every compute entry point raises if
libtv5.so or a CUDA device is missing.
"""
from .lib import (Tv5Error, lib_path, load_library, build_library, exported_symbols)  # noqa: F401
from .engine import (Engine, get_engine, PoseResult, compute_pose, compute_pose_batch,  # noqa: F401
                     solve5, score, score_bounds, ref_rng_sets, decompose_host, decompose_uv_host)
from . import synth  # noqa: F401
