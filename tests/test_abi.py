"""CPU-side checks of the C-ABI library: it loads, exports every symbol the
header declares, and refuses to compute without a CUDA device."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as ge


@pytest.fixture(scope="module")
def pb():
    ge.build_cuda()
    import pdeode_b200
    return pdeode_b200


def test_library_exports_every_declared_symbol(pb):
    lib = pb._lib()
    names = pb.exported_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pdeode_b200.h but not exported"


def test_header_cites_reference_lines(pb):
    with open(pb.HEADER_PATH) as f:
        txt = f.read()
    # every entry point of the path cites the reference lines it replaces
    for fn, cite in [("pdeode_step", "P:700-742"), ("pdeode_mass", "P:772-791"),
                     ("pdeode_peak", "P:978-1004"), ("pdeode_set_state", "P:133-138"),
                     ("pdeode_get_state", "P:685-687")]:
        i = txt.index("int " + fn + "(")
        assert cite in txt[max(0, i - 900):i], (fn, cite)
    assert "torch" not in txt.lower()


def test_params_struct_layout(pb):
    # 9 doubles + 6 ints, no padding surprises: must match pdeode_params
    assert C.sizeof(pb.Params) == 9 * 8 + 6 * 4
    import oracle_lib as ol
    assert C.sizeof(ol.Params) == C.sizeof(pb.Params)
    assert [f[0] for f in ol.Params._fields_] == [f[0] for f in pb.Params._fields_]


def test_strerror_and_version(pb):
    lib = pb._lib()
    assert lib.pdeode_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.pdeode_strerror(pb.ERR_NODEVICE)
    assert b"10000" in lib.pdeode_strerror(pb.ERR_PICARD_CAP)
    assert b"sm_100a" in lib.pdeode_version()
    assert pb.Solver.bytes_per_cell(pb.KERNEL_CG_SPMV) == 48
    assert pb.Solver.bytes_per_cell(pb.KERNEL_CG_UPDATE) == 48


def test_bad_arguments_are_rejected_before_touching_cuda(pb):
    lib = pb._lib()
    h = C.c_void_p()
    p = pb.default_params(8)
    assert lib.pdeode_create(C.byref(h), 2, 8, C.byref(p), 0) == pb.ERR_ARG     # row < 3
    assert lib.pdeode_create(C.byref(h), 8, 1, C.byref(p), 0) == pb.ERR_ARG     # col < 2
    assert lib.pdeode_create(None, 8, 8, C.byref(p), 0) == pb.ERR_ARG
    assert lib.pdeode_step(None, None, None, None) == pb.ERR_ARG


def test_no_device_means_error_not_fallback(pb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pb.PdeodeError) as ei:
        pb.Solver(33, 33, pb.default_params(33))
    assert ei.value.code == pb.ERR_NODEVICE


def test_product_code_never_touches_the_oracle():
    """The product path must not link, import or execute anything under oracle/."""
    root = ge.ROOT
    bad = []
    for base, _, files in os.walk(os.path.join(root, "pde-odesolver_b200")):
        for f in files:
            if f.endswith((".cu", ".h", ".cpp", ".py", ".f90", ".sh", "Makefile")):
                with open(os.path.join(base, f), errors="ignore") as fh:
                    if re.search(r"oracle", fh.read(), re.I):
                        bad.append(os.path.join(base, f))
    assert not bad, bad
