"""CPU checks of the drop-in boundary: libsphb200.so loads, exports every symbol that
include/sphb200.h declares, the ctypes mirrors match the C structs, and the product path fails
LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from sphexample_b200 import _abi, lib as sphlib

import util


def test_library_exports_every_declared_symbol(sph_lib):
    names = sphlib.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(sph_lib, n), f"{n} declared in include/sphb200.h but not exported"
    assert sph_lib.sphb200_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """compile a C program against the header and compare sizeof/offsetof with the ctypes mirror"""
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "sphb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(sphb200_params), offsetof(sphb200_params, rho0),
         offsetof(sphb200_params, k), offsetof(sphb200_params, motions), sizeof(sphb200_motion),
         sizeof(sphb200_report), offsetof(sphb200_report, total_time), offsetof(sphb200_params, cubic_eps));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(sphlib.ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    P, R = _abi.Params, _abi.Report
    want = [C.sizeof(P), P.rho0.offset, P.k.offset, P.motions.offset, C.sizeof(_abi.Motion), C.sizeof(R),
            R.total_time.offset, P.cubic_eps.offset]
    assert got == want


def test_header_cites_the_reference_for_every_entry_point():
    text = open(sphlib.HEADER).read()
    assert text.count("src/") >= 20   # file:line citations of the functions each entry replaces
    assert 'extern "C"' in text


def test_create_fails_loudly_without_gpu(sph_lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    case = util.case_c1()
    p = util.params_of(case)
    h = C.c_void_p()
    rc = sph_lib.sphb200_create(C.byref(p), 0, C.byref(h))
    assert rc == _abi.ECUDA and not h.value
    msg = sph_lib.sphb200_last_error(None).decode()
    assert "no CPU fallback" in msg
    from sphexample_b200.simulation import Simulation, SphError
    with pytest.raises(SphError):
        Simulation(p)


def test_create_rejects_bad_params(sph_lib):
    case = util.case_c1()
    for field, value in (("abi_version", 99), ("dim", 4), ("real_bytes", 2), ("viscosity", 9)):
        p = util.params_of(case)
        setattr(p, field, value)
        h = C.c_void_p()
        assert sph_lib.sphb200_create(C.byref(p), 0, C.byref(h)) == _abi.EINVAL
    assert sph_lib.sphb200_step(None, 1, 0, None) == _abi.EINVAL


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    pkg = os.path.join(sphlib.ROOT, "sphexample_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text and "sph_oracle" not in text, f
