"""CPU-side checks of the boundary: libdppb200.so loads, exports every symbol include/dpp_b200.h
declares, the ctypes mirrors match the header, and -- with no GPU -- the product path fails loudly
instead of computing anything on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dpp_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from perphil_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), name
        assert name in _lib._PROTOTYPES, f"{name} missing from the ctypes prototypes"
    assert set(_lib._PROTOTYPES) == set(names)


def test_struct_layouts_match_header():
    from perphil_b200 import _lib

    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)

    def fields(struct):
        body = re.search(r"typedef struct \{([^{}]*)\} %s;" % struct, src).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ctype, names = decl.split(None, 1)
            for nm in names.split(","):
                out.append((re.sub(r"\[.*", "", nm.strip()), ctype))
        return out

    for struct, cls in (("dpp_options", _lib.DppOptions), ("dpp_result", _lib.DppResult), ("dpp_info", _lib.DppInfo)):
        hdr = fields(struct)
        assert [n for n, _ in hdr] == [n for n, _ in cls._fields_], struct
        for (n, ct), (_, pyt) in zip(hdr, cls._fields_):
            size = {"int32_t": 4, "int64_t": 8, "double": 8}[ct]
            base = pyt._type_ if hasattr(pyt, "_length_") else pyt
            assert ctypes.sizeof(base) == size, (struct, n)


def test_default_options_are_the_reference_tolerances():
    from perphil_b200 import _lib

    lib = _lib.load()
    o = _lib.DppOptions()
    lib.dpp_default_options(ctypes.byref(o))
    assert (o.rtol, o.atol, o.max_it, o.gmres_restart, o.dtol) == (1e-8, 1e-12, 50000, 30, 1e4)  # parameters.py:1,14-16


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import perphil_b200 as pb
    from perphil_b200.backend import DppError

    mesh = pb.UnitSquareMesh(2, 2)
    _, V = pb.create_function_spaces(mesh)
    W = V * V
    prm = pb.DPPParameters()
    bcs = [pb.DirichletBC(W.sub(0), pb.Constant(0.0), "on_boundary"), pb.DirichletBC(W.sub(1), pb.Constant(0.0), "on_boundary")]
    with pytest.raises(DppError, match="no CPU fallback"):
        pb.solve_dpp(W, prm, bcs, solver_parameters=pb.B200_CG_JACOBI_PARAMS)
    with pytest.raises(RuntimeError, match="reference"):
        pb.solve_dpp(W, prm, bcs, solver_parameters={"ksp_type": "gmres"})  # routed to perphil, absent here


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "perphil_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "dpp_oracle" not in txt, f


def test_x_segment_plan_matches_the_measured_optima():
    """Host-only launch planning of the plane-streaming kernels (csrc/dpp_internal.cuh: choose_x_segments): the
    segment counts that tools/sched_sweep.py measured fastest on a B200 (296 resident CTAs, 128 interior tiles of
    a 257^2 plane): 33-plane slab -> 2 (48.5 us; 1: 64, 3: 57, 4: 51.5), 128 layers -> 4 (159.5 us; 2: 164.9,
    9: 165.7), 256 layers -> 9 (303 us; 2 segments of 128 planes: 327, 7: 335, 11: 302, 13: 307)."""
    from perphil_b200 import _lib as L

    lib = L.load()
    plan = lambda planes: lib.dpp_plan_x_segments(128, planes, 296, 4096)
    assert plan(31) == 2
    assert plan(127) == 4
    assert plan(255) == 9
    assert lib.dpp_plan_x_segments(128, 1, 296, 4096) == 1
    assert lib.dpp_plan_x_segments(0, 31, 296, 4096) < 0   # DPP_ERR_INVALID
    # never more CTAs than allowed, never a run shorter than 4 planes (unless a single segment)
    for tiles, planes in ((153, 257), (16, 500), (2000, 64), (8, 9)):
        n = lib.dpp_plan_x_segments(tiles, planes, 296, 4096)
        assert n >= 1 and (n == 1 or (tiles * n <= 4096 and -(-planes // n) >= 4))
