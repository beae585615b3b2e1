"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libref_twin_cuda.so — the reference's own
solver / cheirality / ComputeError<double> / RNG sources compiled by nvcc with thin dump entry
points (oracle/ref_twin/ref_twin_cuda.cu).  GPU box only; None when the library is absent."""
import ctypes as C
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libref_twin_cuda.so")
_T = None


def load():
    global _T
    if _T is None and os.path.exists(PATH):
        T = C.CDLL(PATH)
        vp = C.c_void_p
        T.ref_rng_sets.argtypes = [C.c_int, C.c_int, vp]
        T.ref_score.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_double, vp, vp]
        T.ref_solve_sets.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
        _T = T
    return _T


def score(T, x1, x2, n_test, E_list, thr):
    """int32 [M] inlier counts of the reference's ComputeError<double> (one GPU thread per E)."""
    M = E_list.shape[0]
    cnt = torch.zeros(M, dtype=torch.int32, device=x1.device)
    torch.cuda.synchronize()
    rc = T.ref_score(x1.data_ptr(), x2.data_ptr(), int(n_test), E_list.data_ptr(), M, float(thr), cnt.data_ptr(), None)
    assert rc == 0, f"cuda error {rc}"
    return cnt


def solve_sets(T, x1, x2, sets):
    """The reference's compute_E_matrices_optimized + compute_P_matrices per minimal set."""
    H = sets.shape[0]
    dev = x1.device
    E_all = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    E_val = torch.zeros(H, 10, 9, dtype=torch.float64, device=dev)
    P_val = torch.zeros(H, 10, 12, dtype=torch.float64, device=dev)
    nr = torch.zeros(H, dtype=torch.int32, device=dev)
    nv = torch.zeros(H, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    rc = T.ref_solve_sets(x1.data_ptr(), x2.data_ptr(), x1.shape[0], sets.data_ptr(), H, E_all.data_ptr(),
                          nr.data_ptr(), E_val.data_ptr(), P_val.data_ptr(), nv.data_ptr())
    assert rc == 0, f"cuda error {rc}"
    return dict(E=E_val, P=P_val, n_roots=nr, n_valid=nv)
