"""ctypes mirror of include/sphb200.h (struct layouts and constants only)."""
import ctypes as C

ABI_VERSION = 1
MAX_MOTIONS = 16

OK, EINVAL, ECUDA, ESTATE, ECAPACITY, ENCCL, ENUMERIC = 0, -1, -2, -3, -4, -5, -6

# ParticleType enum values, src/SimulationGeometry.jl:10-14
FLUID, FIXED, MOVING = 1, 2, 3

KERNEL_WENDLANDC2, KERNEL_CUBICSPLINE = 0, 1
VISC_ZERO, VISC_ARTIFICIAL, VISC_LAMINAR, VISC_LAMINAR_SPS = 0, 1, 2, 3
DDT_ZERO, DDT_ZERO_GRAVITY_LINEAR, DDT_LINEAR, DDT_COMPLEX = 0, 1, 2, 3


class Motion(C.Structure):
    _fields_ = [
        ("group_marker", C.c_int64),
        ("velocity", C.c_double),
        ("start_time", C.c_double),
        ("duration", C.c_double),
        ("direction", C.c_double * 3),
    ]


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("dim", C.c_int32),
        ("real_bytes", C.c_int32),
        ("kernel", C.c_int32),
        ("viscosity", C.c_int32),
        ("diffusion", C.c_int32),
        ("shifting", C.c_int32),
        ("kernel_output", C.c_int32),
        ("mdbc", C.c_int32),
        ("n_motions", C.c_int32),
        # SimulationConstants
        ("rho0", C.c_double),
        ("dx", C.c_double),
        ("m0", C.c_double),
        ("alpha", C.c_double),
        ("g", C.c_double),
        ("c0", C.c_double),
        ("gamma", C.c_double),
        ("gamma_inv", C.c_double),
        ("delta_phi", C.c_double),
        ("cfl", C.c_double),
        ("cb", C.c_double),
        ("cb_inv", C.c_double),
        ("nu0", C.c_double),
        ("blin_constant", C.c_double),
        ("smagorinsky_constant", C.c_double),
        # SPHKernelInstance
        ("k", C.c_double),
        ("h", C.c_double),
        ("h_inv", C.c_double),
        ("H", C.c_double),
        ("H_inv", C.c_double),
        ("H2", C.c_double),
        ("alphaD", C.c_double),
        ("eta2", C.c_double),
        ("cubic_eps", C.c_double),
        ("motions", Motion * MAX_MOTIONS),
    ]


class Report(C.Structure):
    _fields_ = [
        ("iteration", C.c_int64),
        ("index_counter", C.c_int64),
        ("n_rebuilds", C.c_int64),
        ("n_particles", C.c_int64),
        ("n_halo", C.c_int64),
        ("total_time", C.c_double),
        ("current_dt", C.c_double),
        ("delta_x", C.c_double),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}
