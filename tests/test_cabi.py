"""CPU tests of the boundary: libtv5.so loads, exports every symbol include/tv5.h declares, and
the host API refuses to run without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest
import torch

import tv5
from tv5 import lib as tv5lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    names = tv5lib.exported_symbols()
    assert len(names) >= 15 and "tv5_compute_pose" in names
    L = C.CDLL(tv5lib.build_library())
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tv5.h but not exported"
    assert L.tv5_version() == 110
    L.tv5_strerror.restype = C.c_char_p
    assert L.tv5_strerror(0) == b"ok" and L.tv5_strerror(-1) == b"invalid argument"


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "tv5.h")).read()
    for cite in ("essential_matrix.cu:190-280", "essential_matrix.cu:110-184", "kernel_functions.cu:231-264",
                 "cheirality.cu:4-214", "essential_matrix_5pt.cu:1224-1249"):
        assert cite in text


def test_result_struct_layout_matches_header():
    assert C.sizeof(tv5lib.Tv5Result) == 32
    text = open(os.path.join(ROOT, "include", "tv5.h")).read()
    body = re.search(r"typedef struct tv5_result \{(.*?)\} tv5_result;", text, re.S).group(1)
    fields = re.findall(r"int32_t\s+(\w+)", body)
    assert fields == [f[0] for f in tv5lib.Tv5Result._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(tv5.Tv5Error):
        tv5.get_engine()
    L = tv5lib.load_library()
    h = C.c_void_p()
    assert L.tv5_create(0, C.byref(h)) == -4  # TV5_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deep-sfm-revisited_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("tv5_oracle", "").lower() or f == "synth.py", f


def test_shim_signatures_and_errors():
    import essential_matrix as em
    for name in ("initialise", "optimise", "computeP", "decompose", "decomposeUV"):
        assert callable(getattr(em, name))
    if not torch.cuda.is_available():
        x = torch.zeros(10, 2, dtype=torch.float64)
        with pytest.raises((RuntimeError, tv5.Tv5Error)):
            em.computeP(x, x, 10, 10, 1, 1e-4)
    with pytest.raises(RuntimeError, match="must be a double tensor"):
        em.optimise(torch.zeros(4, 2), torch.zeros(4, 2, dtype=torch.float64), torch.eye(3, dtype=torch.float64), 1e-4, 1.0, 3)
    with pytest.raises(RuntimeError, match="E_init must be contiguous"):
        em.optimise(torch.zeros(4, 2, dtype=torch.float64), torch.zeros(4, 2, dtype=torch.float64),
                    torch.ones(3, 3, dtype=torch.float64).t(), 1e-4, 1.0, 3)


def test_host_decompose_matches_reference_golden():
    """tv5_decompose / tv5_decompose_uv run on the host like the reference's Edecomp
    (polish_E.cu:147-338); bit-identical to the reference extension's outputs."""
    import numpy as np
    import essential_matrix as em
    g = np.load(os.path.join(ROOT, "tests", "golden", "polish_ref.npz"))
    for i in range(g["E"].shape[0]):
        E = torch.from_numpy(g["E"][i].copy())
        U, V = em.decomposeUV(E)
        a = em.decompose(E)
        assert (U.numpy() == g["U"][i]).all() and (V.numpy() == g["V"][i]).all()   # bit-exact
        assert (a.numpy() == g["angles"][i]).all()
        assert a.dtype == torch.float64 and a.shape == (5,) and U.shape == (3, 3)
    with pytest.raises(RuntimeError):
        em.decompose(torch.eye(3))
