"""ctypes mirror of include/swiftgpu.h (the C ABI of libswiftgpu).

Only plain ctypes here: the product path is the CUDA shared library
``swift_b200/libswiftgpu.so``; there is no CPU fallback. Loading fails loudly
if the library has not been built (``python -c 'import __graft_entry__ as g;
g.build()'`` or ``make -C swift_b200/csrc``).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

ABI_VERSION = 2  # SWIFTGPU_ABI_VERSION of include/swiftgpu.h
SCHEME_MINIMAL, SCHEME_GADGET2, SCHEME_SPHENIX = 0, 1, 2
SCHEMES = {"minimal": 0, "gadget2": 1, "sphenix": 2}

PHASE_SORT, PHASE_DENSITY, PHASE_GHOST, PHASE_GRADIENT = 1, 2, 4, 8
PHASE_EXTRA_GHOST, PHASE_FORCE, PHASE_END_FORCE, PHASE_ALL = 16, 32, 64, 0x7F

LAYOUT_FIELDS = [
    "size", "id", "x", "v", "a_hydro", "mass", "h", "u", "u_dt", "entropy",
    "entropy_dt", "rho", "wcount", "wcount_dh", "rho_dh", "rot_v", "div_v",
    "f", "pressure", "P_over_rho2", "soundspeed", "v_sig", "h_dt", "balsara",
    "div_v_dt", "div_v_previous_step", "visc_alpha", "laplace_u", "diff_alpha",
    "alpha_visc_max_ngb", "time_bin", "depth_h", "min_ngb_time_bin",
]


class PartLayout(C.Structure):
    _fields_ = [(n, C.c_int32) for n in LAYOUT_FIELDS]

    def as_dict(self):
        return {n: getattr(self, n) for n in LAYOUT_FIELDS}

    @classmethod
    def from_dict(cls, d):
        out = cls()
        for n in LAYOUT_FIELDS:
            setattr(out, n, int(d[n]))
        return out


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("scheme", C.c_int32),
        ("device", C.c_int32), ("periodic", C.c_int32),
        ("dim", C.c_double * 3),
        ("eta_neighbours", C.c_float), ("h_tolerance", C.c_float),
        ("h_max", C.c_float), ("h_min", C.c_float),
        ("max_smoothing_iterations", C.c_int32),
        ("use_mass_weighted_num_ngb", C.c_int32),
        ("CFL_condition", C.c_float),
        ("viscosity_alpha", C.c_float), ("viscosity_alpha_max", C.c_float),
        ("viscosity_alpha_min", C.c_float), ("viscosity_length", C.c_float),
        ("diffusion_alpha", C.c_float), ("diffusion_beta", C.c_float),
        ("diffusion_alpha_max", C.c_float), ("diffusion_alpha_min", C.c_float),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("layout", PartLayout),
    ]


class Step(C.Structure):
    _fields_ = [
        ("ti_current", C.c_int64), ("max_active_bin", C.c_int32),
        ("with_cosmology", C.c_int32), ("time_base", C.c_double),
        ("a", C.c_float), ("H", C.c_float),
    ]


class XpartLayout(C.Structure):
    _fields_ = [("size", C.c_int32), ("x_diff", C.c_int32), ("x_diff_sort", C.c_int32), ("v_full", C.c_int32),
                ("u_full", C.c_int32)]


class DriftArgs(C.Structure):
    _fields_ = [("dt_drift", C.c_double), ("dt_kick_hydro", C.c_double), ("dt_therm", C.c_double),
                ("minimal_internal_energy", C.c_float), ("init_particles", C.c_int32)]


class Cell(C.Structure):
    _fields_ = [
        ("loc", C.c_double * 3), ("width", C.c_double * 3),
        ("dmin", C.c_float), ("h_min_allowed", C.c_float),
        ("h_max_allowed", C.c_float), ("h_max", C.c_float),
        ("h_max_active", C.c_float), ("h_max_old", C.c_float),
        ("dx_max_part", C.c_float), ("dx_max_part_old", C.c_float),
        ("dx_max_sort", C.c_float), ("dx_max_sort_old", C.c_float),
        ("depth", C.c_int32), ("split", C.c_int32), ("parent", C.c_int32),
        ("progeny", C.c_int32 * 8), ("nodeID", C.c_int32), ("top", C.c_int32),
        ("count", C.c_int32), ("first_part", C.c_int64),
        ("ti_end_min", C.c_int64),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_sort", C.c_double), ("ms_density", C.c_double),
        ("ms_ghost", C.c_double), ("ms_gradient", C.c_double),
        ("ms_extra_ghost", C.c_double), ("ms_force", C.c_double),
        ("ms_end_force", C.c_double),
        ("n_density", C.c_int64), ("n_gradient", C.c_int64),
        ("n_force", C.c_int64), ("n_launches", C.c_int64),
        ("t_density", C.c_int64), ("t_gradient", C.c_int64), ("t_force", C.c_int64),
        ("ghost_iterations", C.c_int32), ("ghost_unconverged", C.c_int32),
        ("force_list_rebuilds", C.c_int32), ("gradient_list_rebuilds", C.c_int32),
        ("n_host_syncs", C.c_int64),
    ]


# numpy dtype equivalent of struct swiftgpu_cell
def cell_dtype():
    import numpy as np
    return np.dtype([
        ("loc", "<f8", 3), ("width", "<f8", 3), ("dmin", "<f4"),
        ("h_min_allowed", "<f4"), ("h_max_allowed", "<f4"), ("h_max", "<f4"),
        ("h_max_active", "<f4"), ("h_max_old", "<f4"), ("dx_max_part", "<f4"),
        ("dx_max_part_old", "<f4"), ("dx_max_sort", "<f4"),
        ("dx_max_sort_old", "<f4"), ("depth", "<i4"), ("split", "<i4"),
        ("parent", "<i4"), ("progeny", "<i4", 8), ("nodeID", "<i4"),
        ("top", "<i4"), ("count", "<i4"), ("first_part", "<i8"),
        ("ti_end_min", "<i8")], align=False)


# (name, restype, argtypes) of every symbol include/swiftgpu.h declares.
VP, I32, I64 = C.c_void_p, C.c_int32, C.c_int64
EXPORTS = [
    ("swiftgpu_default_layout", C.c_int, [C.c_int, C.POINTER(PartLayout)]),
    ("swiftgpu_default_config", C.c_int, [C.c_int, C.POINTER(Config)]),
    ("swiftgpu_init", C.c_int, [C.POINTER(VP), C.POINTER(Config)]),
    ("swiftgpu_destroy", None, [VP]),
    ("swiftgpu_last_error", C.c_char_p, [VP]),
    ("swiftgpu_upload_cells", C.c_int, [VP, VP, I32, VP, I32]),
    ("swiftgpu_upload_parts", C.c_int, [VP, VP, I64]),
    ("swiftgpu_upload_parts_device", C.c_int, [VP, VP, I64]),
    ("swiftgpu_upload_parts_local", C.c_int, [VP, VP, I64, I64]),
    ("swiftgpu_download_parts_local", C.c_int, [VP, VP, I64]),
    ("swiftgpu_set_step", C.c_int, [VP, C.POINTER(Step)]),
    ("swiftgpu_set_stream", C.c_int, [VP, VP]),
    ("swiftgpu_run_sort", C.c_int, [VP]),
    ("swiftgpu_run_density", C.c_int, [VP]),
    ("swiftgpu_run_ghost", C.c_int, [VP]),
    ("swiftgpu_run_gradient", C.c_int, [VP]),
    ("swiftgpu_run_extra_ghost", C.c_int, [VP]),
    ("swiftgpu_run_force", C.c_int, [VP]),
    ("swiftgpu_run_end_force", C.c_int, [VP]),
    ("swiftgpu_run_step", C.c_int, [VP, C.c_uint32]),
    ("swiftgpu_download_parts", C.c_int, [VP, VP, I64]),
    ("swiftgpu_download_parts_device", C.c_int, [VP, VP, I64]),
    ("swiftgpu_download_cells", C.c_int, [VP, VP, I32]),
    ("swiftgpu_download_counts", C.c_int, [VP, VP, VP, VP, I64]),
    ("swiftgpu_download_timestep", C.c_int, [VP, VP, I64]),
    ("swiftgpu_upload_xparts", C.c_int, [VP, C.POINTER(XpartLayout), VP, I64]),
    ("swiftgpu_download_xparts", C.c_int, [VP, VP, I64]),
    ("swiftgpu_run_drift", C.c_int, [VP, C.POINTER(DriftArgs)]),
    ("swiftgpu_run_kick", C.c_int, [VP, C.c_int, C.c_float]),
    ("swiftgpu_run_limiter", C.c_int, [VP, I32]),
    ("swiftgpu_get_stats", C.c_int, [VP, C.POINTER(Stats)]),
    ("swiftgpu_download_sort", C.c_int, [VP, I32, I32, VP, VP, VP]),
    ("swiftgpu_worklist_stats", C.c_int, [C.POINTER(Config), C.POINTER(Step), VP, I32, VP, I32, C.c_int, VP]),
    ("swiftgpu_worklist_digest", C.c_int, [C.POINTER(Config), C.POINTER(Step), VP, I32, VP, I32, C.c_int, VP]),
    ("swiftgpu_nccl_unique_id", C.c_int, [VP]),
    ("swiftgpu_halo_setup", C.c_int, [VP, VP]),
    ("swiftgpu_halo_exchange", C.c_int, [VP, C.c_int]),
    ("swiftgpu_halo_plan", C.c_int, [C.POINTER(Config), VP, I32, VP, I32, I32, VP, VP, VP, VP, VP, VP]),
]

_lib = None
_host = None


def lib_path():
    # SWIFTGPU_LIB: another BUILD of the same CUDA library (A/B of compile-time parameters)
    return os.environ.get("SWIFTGPU_LIB") or os.path.join(HERE, "libswiftgpu.so")


def load():
    """Load libswiftgpu.so (CUDA). Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: the CUDA extension is the only compute "
                "path (no CPU fallback). Build it with __graft_entry__.build().")
        lib = C.CDLL(path, mode=os.RTLD_NOW)
        for name, res, args in EXPORTS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def load_host():
    """Load libswiftgpu_host.so: host-only helpers (tree builder, AoS packer)
    standing in for the SWIFT engine that normally owns the cells/particles."""
    global _host
    if _host is None:
        path = os.path.join(HERE, "libswiftgpu_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing; run __graft_entry__.build()")
        lib = C.CDLL(path, mode=os.RTLD_NOW)
        lib.swifthost_build_tree.restype = VP
        lib.swifthost_build_tree.argtypes = [VP, VP, VP, I64, VP, VP, C.c_int,
                                             C.c_int, I64, VP, VP, VP]
        lib.swifthost_tree_ncells.restype = I32
        lib.swifthost_tree_ncells.argtypes = [VP]
        lib.swifthost_tree_ntop.restype = I32
        lib.swifthost_tree_ntop.argtypes = [VP]
        lib.swifthost_tree_copy.restype = None
        lib.swifthost_tree_copy.argtypes = [VP, VP, VP]
        lib.swifthost_tree_free.restype = None
        lib.swifthost_tree_free.argtypes = [VP]
        lib.swifthost_pack_parts.restype = None
        lib.swifthost_pack_parts.argtypes = [C.POINTER(PartLayout), C.c_int, I64] + [VP] * 14
        _host = lib
    return _host
