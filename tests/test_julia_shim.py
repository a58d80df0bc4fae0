"""The Julia binding (julia/SPHExampleB200.jl) cannot be executed here (no Julia in the image), so
its struct layouts and ccall signatures are checked TEXTUALLY against the C header's ctypes mirror:
same field names, order and widths, every ccall names an exported symbol with the right arity."""
import ctypes as C
import os
import re

from sphexample_b200 import _abi
from sphexample_b200 import lib as sphlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "julia", "SPHExampleB200.jl"), encoding="utf-8").read()
JL = {"Int32": C.c_int32, "Int64": C.c_int64, "Float64": C.c_double}


def _struct_fields(name):
    body = re.search(r"struct " + name + r"\b(.*?)\nend", SRC, re.S).group(1)
    body = re.sub(r"#.*", "", body)
    out = []
    for stmt in re.split(r"[;\n]", body):
        m = re.match(r"\s*(\w+)::([\w{},. ]+?)\s*$", stmt)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return out


def test_params_struct_matches_the_header():
    jl = _struct_fields("SphParams")
    cf = _abi.Params._fields_
    assert [n for n, _ in jl] == [n for n, _ in cf]
    for (n, jt), (_, ct) in zip(jl, cf):
        if n == "motions":
            assert jt == "NTuple{SPHB200_MAX_MOTIONS,SphMotion}" and ct._length_ == _abi.MAX_MOTIONS
        else:
            assert JL[jt] is ct, n
    assert re.search(r"const SPHB200_MAX_MOTIONS = (\d+)", SRC).group(1) == str(_abi.MAX_MOTIONS)
    assert re.search(r"const SPHB200_ABI_VERSION = Int32\((\d+)\)", SRC).group(1) == str(_abi.ABI_VERSION)


def test_motion_and_report_structs_match_the_header():
    jm = _struct_fields("SphMotion")
    assert [n for n, _ in jm] == [n for n, _ in _abi.Motion._fields_]
    assert dict(jm)["direction"] == "NTuple{3,Float64}" and dict(jm)["group_marker"] == "Int64"
    jr = _struct_fields("SphReport")
    assert [n for n, _ in jr][:len(_abi.Report._fields_)] == [n for n, _ in _abi.Report._fields_]
    for (n, jt), (_, ct) in zip(jr, _abi.Report._fields_):
        assert JL[jt] is ct, n


def test_every_ccall_names_an_exported_symbol_with_the_right_arity():
    declared = set(sphlib.declared_symbols())
    header = open(sphlib.HEADER).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    calls = re.findall(r"ccall\(\(:(\w+), libsphb200\),\s*(\w+),\s*\((.*?)\),", SRC, re.S)
    assert {c[0] for c in calls} >= {"sphb200_create", "sphb200_destroy", "sphb200_upload", "sphb200_download",
                                     "sphb200_simulation_loop", "sphb200_get_report", "sphb200_last_error"}
    for sym, ret, args in calls:
        assert sym in declared, sym
        proto = re.search(r"\b" + sym + r"\s*\((.*?)\)\s*;", header, re.S).group(1)
        n_c = 0 if proto.strip() in ("", "void") else proto.count(",") + 1
        n_jl = len([a for a in re.split(r",(?![^{}]*\})", args) if a.strip()])
        assert n_c == n_jl, (sym, n_c, n_jl)
