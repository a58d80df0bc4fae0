"""Host-side mirror of the reference's step API on top of the C-ABI (libsphb200.so).

`Simulation` owns one device handle; its methods carry the reference's function names
(src/SPHCellList.jl:3, src/TimeStepping.jl:3, src/SimulationEquations.jl:3) so that parity
tests read like the reference's own loop.  `RunSimulation` mirrors the driver
(src/SPHCellList.jl:808-930) with the output/log/ParaView stages left to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional

import numpy as np

from . import _abi
from .config import make_params, next_output_time
from .lib import lib
from .preprocess import SimParticles


# order of SPHB200_STAGE_* in include/sphb200.h, with the reference's timer labels
STAGE_NAMES = ("01 Update TimeStep (+S0/S1 reductions, control)", "02 Calculate IndexCounter (UpdateNeighbors!)",
               "Motion (first) + state-n snapshots", "04 Apply MDBC before Half TimeStep", "neighbour-list build + reorder",
               "05-07 First NeighborLoop + half step (fused)", "Motion (second)",
               "08-11 Second NeighborLoop + full step (fused)", "12 Update MetaData", "slab: all-reduce (incl. wait for the slowest rank)",
               "slab: pass-1 halo exchange", "slab: pass-1 halo exchange not hidden", "slab: pass-2 halo exchange",
               "slab: pass-2 halo exchange not hidden", "", "")


class SphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sphb200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Simulation:
    def __init__(self, params: _abi.Params, device: int = 0):
        self._L = lib()
        self.params = params
        self.D = int(params.dim)
        self.dtype = np.float64 if params.real_bytes == 8 else np.float32
        h = C.c_void_p()
        rc = self._L.sphb200_create(C.byref(params), int(device), C.byref(h))
        if rc != 0:
            raise SphError(rc, (self._L.sphb200_last_error(None) or b"").decode())
        self._h = h

    # ---- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.sphb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise SphError(rc, (self._L.sphb200_last_error(self._h) or b"").decode())

    # ---- state transfer -------------------------------------------------------------------
    def upload(self, particles: SimParticles):
        """Hand the SimParticles table to the device (the library copies)."""
        f = lambda a: np.ascontiguousarray(a, dtype=self.dtype)
        self._keep = [f(particles.Position), f(particles.Velocity), f(particles.Acceleration), f(particles.Density),
                      np.ascontiguousarray(particles.Type, np.uint8), np.ascontiguousarray(particles.GroupMarker, np.uint64),
                      np.ascontiguousarray(particles.ID, np.int64), f(particles.GhostPoints), f(particles.GhostNormals)]
        n = self._keep[0].shape[0]
        assert self._keep[0].shape[1] == self.D
        self._ck(self._L.sphb200_upload(self._h, n, *[_ptr(a) for a in self._keep]))
        self.N = n
        return self

    def upload_arrays(self, position, velocity, density, types, acceleration=None, group=None, ids=None):
        """Upload from caller-owned (already typed, contiguous) host arrays without conversions."""
        n = position.shape[0]
        for name, a, dt, shape in (("position", position, self.dtype, (n, self.D)), ("velocity", velocity, self.dtype, (n, self.D)),
                                   ("acceleration", acceleration, self.dtype, (n, self.D)), ("density", density, self.dtype, (n,)),
                                   ("types", types, np.uint8, (n,)), ("group", group, np.uint64, (n,)), ("ids", ids, np.int64, (n,))):
            if a is None:
                continue
            if a.dtype != dt or a.shape != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"upload_arrays: {name} must be a C-contiguous {np.dtype(dt).name} array of shape {shape}, "
                                 f"got {a.dtype} {a.shape} (the library reads raw memory)")
        self._ck(self._L.sphb200_upload(self._h, n, _ptr(position), _ptr(velocity), _ptr(acceleration), _ptr(density),
                                        _ptr(types), _ptr(group), _ptr(ids), None, None))
        self.N = n

    def download(self, order: str = "device", fields=("Position", "Velocity", "Acceleration", "Density", "Pressure",
                                                        "ID", "Type", "GroupMarker", "Cells")):
        n, d, t = int(self._L.sphb200_num_particles(self._h)), self.D, self.dtype
        out = {}
        mk = {"Position": ((n, d), t), "Velocity": ((n, d), t), "Acceleration": ((n, d), t), "Density": ((n,), t),
              "Pressure": ((n,), t), "ID": ((n,), np.int64), "Type": ((n,), np.uint8), "GroupMarker": ((n,), np.uint64),
              "Cells": ((n, d), np.int64)}
        for k in fields:
            out[k] = np.empty(*mk[k])
        g = lambda k: _ptr(out.get(k))
        self._ck(self._L.sphb200_download(self._h, 1 if order == "id" else 0, g("Position"), g("Velocity"),
                                          g("Acceleration"), g("Density"), g("Pressure"), g("ID"), g("Type"),
                                          g("GroupMarker"), g("Cells")))
        return out

    def download_into(self, position=None, velocity=None, density=None, pressure=None, order: str = "device"):
        self._ck(self._L.sphb200_download(self._h, 1 if order == "id" else 0, _ptr(position), _ptr(velocity), None,
                                          _ptr(density), _ptr(pressure), None, None, None, None))

    def download_half(self):
        n, d, t = self.N, self.D, self.dtype
        x, v, r, p = np.empty((n, d), t), np.empty((n, d), t), np.empty(n, t), np.empty(n, t)
        self._ck(self._L.sphb200_download_half(self._h, _ptr(x), _ptr(v), _ptr(r), _ptr(p)))
        return {"Position": x, "Velocity": v, "Density": r, "Pressure": p}

    def download_aux(self):
        n, d, t = self.N, self.D, self.dtype
        out = {}
        if self.params.shifting:
            out["gradC"], out["divr"] = np.empty((n, d), t), np.empty(n, t)
        if self.params.kernel_output:
            out["Kernel"], out["KernelGradient"] = np.empty(n, t), np.empty((n, d), t)
        self._ck(self._L.sphb200_download_aux(self._h, _ptr(out.get("gradC")), _ptr(out.get("divr")),
                                              _ptr(out.get("Kernel")), _ptr(out.get("KernelGradient"))))
        return out

    # ---- the loop -------------------------------------------------------------------------
    def SimulationLoop(self, next_output: float) -> dict:
        """SimulationLoop(...), src/SPHCellList.jl:727-805"""
        r = _abi.Report()
        self._ck(self._L.sphb200_simulation_loop(self._h, float(next_output), C.byref(r)))
        return r.as_dict()

    def step(self, n: int = 1, reset_delta_x: bool = False) -> dict:
        r = _abi.Report()
        self._ck(self._L.sphb200_step(self._h, int(n), int(bool(reset_delta_x)), C.byref(r)))
        return r.as_dict()

    def report(self) -> dict:
        r = _abi.Report()
        self._ck(self._L.sphb200_get_report(self._h, C.byref(r)))
        return r.as_dict()

    def set_time(self, t: float, iteration: int = 0):
        self._ck(self._L.sphb200_set_time(self._h, float(t), int(iteration)))

    def set_stream(self, cuda_stream: int):
        self._ck(self._L.sphb200_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_option(self, name: str, value: float):
        self._ck(self._L.sphb200_set_option(self._h, name.encode(), float(value)))

    def stat(self, name: str) -> float:
        v = C.c_double()
        self._ck(self._L.sphb200_get_stat(self._h, name.encode(), C.byref(v)))
        return float(v.value)

    def stage_times(self) -> dict:
        """Device time (ms) of the stages of ONE extra step, keyed by the names of include/sphb200.h
        (SPHB200_STAGE_*): the reference's "01" .. "12" timer report, src/SPHCellList.jl:748-800."""
        ms = (C.c_double * len(STAGE_NAMES))()
        self._ck(self._L.sphb200_stage_times(self._h, ms, len(STAGE_NAMES)))
        return {k: float(v) for k, v in zip(STAGE_NAMES, ms) if k}

    @property
    def num_particles(self) -> int:
        """particles this handle reports (slab mode: the owned ones)"""
        return int(self._L.sphb200_num_particles(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.sphb200_launch_count(self._h))

    # ---- the reference's exported step functions ---------------------------------------------
    def UpdateNeighbors(self) -> int:
        ic = C.c_int64()
        self._ck(self._L.sphb200_update_neighbors(self._h, C.byref(ic)))
        return int(ic.value)

    def cell_list(self):
        nc = C.c_int64()
        self._ck(self._L.sphb200_get_cell_list(self._h, C.byref(nc), None, None))
        cells = np.empty((nc.value, self.D), np.int64)
        start = np.empty(nc.value + 1, np.int64)
        self._ck(self._L.sphb200_get_cell_list(self._h, C.byref(nc), _ptr(cells), _ptr(start)))
        return cells, start

    def Pressure(self, half: int = 0):
        self._ck(self._L.sphb200_pressure(self._h, int(half)))

    def NeighborLoop(self, pass_: int = 0):
        """ResetStep! + NeighborLoop! + ReductionStep!; returns (dρdt, acceleration) in device order."""
        d = np.empty(self.N, self.dtype)
        a = np.empty((self.N, self.D), self.dtype)
        self._ck(self._L.sphb200_neighbor_loop(self._h, int(pass_), _ptr(d), _ptr(a)))
        return d, a

    def DeltaT(self) -> float:
        dt = C.c_double()
        self._ck(self._L.sphb200_delta_t(self._h, C.byref(dt)))
        return float(dt.value)

    def ProgressMotion(self, dt2: float):
        self._ck(self._L.sphb200_progress_motion(self._h, float(dt2)))

    def ApplyMDBCBeforeHalf(self):
        self._ck(self._L.sphb200_apply_mdbc(self._h))

    def HalfTimeStep(self, dt2: float):
        self._ck(self._L.sphb200_half_time_step(self._h, float(dt2)))

    def FullTimeStep(self, dt: float):
        self._ck(self._L.sphb200_full_time_step(self._h, float(dt)))

    # ---- slabs ----------------------------------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, world: int, axis: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._L.sphb200_comm_init(self._h, buf, int(rank), int(world), int(axis)))

    def set_slab(self, lo: int, hi: int):
        self._ck(self._L.sphb200_set_slab(self._h, int(lo), int(hi)))

    def set_ghost_nodes(self, ghost_points: np.ndarray, particle_ids: np.ndarray):
        """Slab-mode SimpleMDBC: the GLOBAL ghost-node table (same on every rank) — the nonzero rows of
        SimParticles.GhostPoints and the IDs of their particles, ascending (src/SPHCellList.jl:228-231)."""
        gp = np.ascontiguousarray(ghost_points, dtype=self.dtype).reshape(-1, self.D)
        ids = np.ascontiguousarray(particle_ids, dtype=np.int64)
        assert gp.shape[0] == ids.shape[0]
        self._ck(self._L.sphb200_set_ghost_nodes(self._h, int(ids.shape[0]), _ptr(gp), ids.ctypes.data_as(C.POINTER(C.c_int64))))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    rc = lib().sphb200_comm_unique_id(buf)
    if rc != 0:
        raise SphError(rc, "ncclGetUniqueId failed (is libnccl loadable?)")
    return bytes(buf)


def simulation_from_case(case, device: int = 0, geometry=()) -> Simulation:
    p = make_params(case.meta, case.consts, case.kernel, case.viscosity, case.diffusion, geometry)
    sim = Simulation(p, device)
    sim.upload(case.particles)
    return sim


def RunSimulation(*, SimGeometry=(), SimMetaData, SimConstants, SimKernel, SimParticles, SimViscosity,
                  SimDensityDiffusion, SimLogger=None, device: int = 0,
                  save_particles: Optional[Callable[[int, dict, dict], None]] = None, max_outputs: Optional[int] = None):
    """RunSimulation(; ...), src/SPHCellList.jl:808-930: the outer `while true` over output
    intervals with the hot loop on the GPU.  Output writers / logger stay the caller's
    (save_particles(output_index, state, report) is called where the reference saves, :891-894)."""
    params = make_params(SimMetaData, SimConstants, SimKernel, SimViscosity, SimDensityDiffusion, SimGeometry)
    sim = Simulation(params, device)
    sim.upload(SimParticles)
    meta = SimMetaData
    if meta.TotalTime != 0.0 or meta.Iteration != 0:      # a restart: the device clock drives ProgressMotion and the loop condition
        sim.set_time(meta.TotalTime, meta.Iteration)
    outputs = 0
    close_files = None
    if save_particles is None and meta.SaveLocation:
        # the reference's own output: VTKHDF particle files under SaveLocation (SetupVTKOutput, :845-849),
        # the initial state first, with OutputIterationCounter = 1
        from . import output as _out
        os.makedirs(meta.SaveLocation, exist_ok=True)
        grid = bool(getattr(meta, "ExportGridCells", False))
        fns = _out.SetupVTKOutput(meta.SaveLocation, meta.SimulationName or "Simulation",
                                  export_single=bool(getattr(meta, "ExportSingleVTKHDF", True)),
                                  variable_names=getattr(meta, "OutputVariables", None), particles=SimParticles,
                                  export_grid_cells=grid, H=float(SimKernel.H))
        save_vtk, close_files = fns[0], fns[1]

        def save_particles(counter, state, rep):
            t = rep["total_time"] if rep else meta.TotalTime
            save_vtk(counter, t, state)
            if grid:                                               # output.save_grid(...), :850,891
                if rep is None:
                    sim.UpdateNeighbors()                          # the reference builds the cell list before the first output (:838-843)
                fns[2](counter, t, sim.cell_list()[0])
        meta.OutputIterationCounter = 1
        save_particles(1, sim.download(), None)
    try:
        while True:
            rep = sim.SimulationLoop(next_output_time(meta))                          # :883
            steps = rep["iteration"] - meta.Iteration
            meta.Iteration, meta.TotalTime, meta.CurrentTimeStep = rep["iteration"], rep["total_time"], rep["current_dt"]
            meta.StepsTakenForLastOutput = steps
            meta.TimeSteps.append(rep["current_dt"])                                   # :884
            meta.OutputIterationCounter += 1                                          # :888
            if save_particles is not None:
                save_particles(meta.OutputIterationCounter, sim.download(), rep)      # :891-894
            if SimLogger is not None:
                SimLogger(meta, rep)
            outputs += 1
            if meta.TotalTime > meta.SimulationTime or (max_outputs is not None and outputs >= max_outputs):   # :909
                break
        state = sim.download(order="id")
    finally:
        if close_files is not None:
            close_files()                                                             # "13B Close Data Streams", :911
        sim.close()
    return state
