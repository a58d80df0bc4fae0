"""ctypes loader of libsphb200.so (the C-ABI declared in include/sphb200.h).

The library is the product; there is no CPU fallback.  Loading fails loudly when the shared
object has not been built (python -c 'import __graft_entry__ as g; g.build()').
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import sys

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsphb200.so")
if os.environ.get("SPHB200_LIB"):          # a variant build (tuning sweeps, scripts/build_variants.sh); never built implicitly
    LIB_PATH = os.path.abspath(os.environ["SPHB200_LIB"])
HEADER = os.path.join(ROOT, "include", "sphb200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc cross-compiles for sm_100a (works without a GPU)."""
    deps = sources() + [HEADER]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in deps)
    if os.environ.get("SPHB200_LIB"):
        stale = force = False
    if force or stale:
        os.makedirs(LIB_DIR, exist_ok=True)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB_PATH, os.path.join(CSRC, "sphb200.cu"), "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            sys.stderr.write(res.stderr)
    return LIB_PATH


def declared_symbols():
    """Every function name include/sphb200.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphb200_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() (nvcc, sm_100a). "
                           "sphexample_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    P = C.POINTER
    sim = vp
    sig = {
        "sphb200_abi_version": (C.c_int, []),
        "sphb200_create": (C.c_int, [P(_abi.Params), C.c_int, P(vp)]),
        "sphb200_destroy": (C.c_int, [sim]),
        "sphb200_last_error": (C.c_char_p, [sim]),
        "sphb200_upload": (C.c_int, [sim, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "sphb200_download": (C.c_int, [sim, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "sphb200_num_particles": (i64, [sim]),
        "sphb200_set_time": (C.c_int, [sim, dbl, i64]),
        "sphb200_simulation_loop": (C.c_int, [sim, dbl, P(_abi.Report)]),
        "sphb200_step": (C.c_int, [sim, i64, C.c_int, P(_abi.Report)]),
        "sphb200_get_report": (C.c_int, [sim, P(_abi.Report)]),
        "sphb200_launch_count": (i64, [sim]),
        "sphb200_set_stream": (C.c_int, [sim, vp]),
        "sphb200_set_option": (C.c_int, [sim, C.c_char_p, dbl]),
        "sphb200_get_stat": (C.c_int, [sim, C.c_char_p, P(dbl)]),
        "sphb200_stage_times": (C.c_int, [sim, P(dbl), C.c_int]),
        "sphb200_update_neighbors": (C.c_int, [sim, P(i64)]),
        "sphb200_get_cell_list": (C.c_int, [sim, P(i64), vp, vp]),
        "sphb200_pressure": (C.c_int, [sim, C.c_int]),
        "sphb200_neighbor_loop": (C.c_int, [sim, C.c_int, vp, vp]),
        "sphb200_delta_t": (C.c_int, [sim, P(dbl)]),
        "sphb200_progress_motion": (C.c_int, [sim, dbl]),
        "sphb200_apply_mdbc": (C.c_int, [sim]),
        "sphb200_half_time_step": (C.c_int, [sim, dbl]),
        "sphb200_full_time_step": (C.c_int, [sim, dbl]),
        "sphb200_download_half": (C.c_int, [sim, vp, vp, vp, vp]),
        "sphb200_download_aux": (C.c_int, [sim, vp, vp, vp, vp]),
        "sphb200_comm_unique_id": (C.c_int, [vp]),
        "sphb200_comm_init": (C.c_int, [sim, vp, C.c_int, C.c_int, C.c_int]),
        "sphb200_set_slab": (C.c_int, [sim, i64, i64]),
        "sphb200_set_ghost_nodes": (C.c_int, [sim, i64, vp, P(i64)]),
        "sphb200_column_histogram": (C.c_int, [sim, C.c_int, P(i64), P(i64), vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
