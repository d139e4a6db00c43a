"""ctypes wrapper of oracle/_ref/libswiftref_<scheme>.so - the UNMODIFIED
reference compiled by oracle/Makefile and driven by oracle/ref_driver.c.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from swift_b200 import abi  # noqa: E402  (struct definitions of the boundary only)

VP = C.c_void_p
_libs = {}


def available(variant):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", f"libswiftref_{variant}.so"))


def load(variant):
    """variant: minimal | gadget2 | sphenix | sphenix_chk"""
    if variant not in _libs:
        path = os.path.join(ROOT, "oracle", "_ref", f"libswiftref_{variant}.so")
        lib = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL)
        lib.swiftref_create.restype = VP
        lib.swiftref_create.argtypes = [C.POINTER(abi.Config), C.POINTER(abi.Step), VP, C.c_int, VP, C.c_int, VP, C.c_longlong]
        lib.swiftref_destroy.argtypes = [VP]
        lib.swiftref_run.argtypes = [VP, C.c_uint, C.c_int, VP]
        lib.swiftref_set_parts.argtypes = [VP, VP]
        lib.swiftref_get_parts.argtypes = [VP, VP]
        lib.swiftref_get_cells.argtypes = [VP, VP]
        lib.swiftref_get_counts.argtypes = [VP, VP, VP, VP]
        lib.swiftref_get_sort.argtypes = [VP, C.c_int, C.c_int, VP, VP]
        lib.swiftref_get_timesteps.argtypes = [VP, VP]
        lib.swiftref_layout.argtypes = [C.POINTER(abi.PartLayout)]
        lib.swiftref_xpart_layout.argtypes = [C.POINTER(abi.XpartLayout)]
        lib.swiftref_set_xparts.argtypes = [VP, VP]
        lib.swiftref_get_xparts.argtypes = [VP, VP]
        lib.swiftref_run_drift.argtypes = [VP, C.c_longlong, C.c_float, C.c_int]
        lib.swiftref_run_kick.argtypes = [VP, C.c_int, C.c_float]
        lib.swiftref_run_limiter.argtypes = [VP, C.c_int]
        lib.swiftref_space_split.argtypes = [C.POINTER(abi.Config), C.POINTER(abi.Step), VP, C.c_longlong, VP, VP, VP, C.c_int]
        _libs[variant] = lib
    return _libs[variant]


def layout(variant):
    L = abi.PartLayout()
    load(variant).swiftref_layout(C.byref(L))
    return L


def space_split(variant, cfg, step, parts_u8, loc, width, max_cells=1 << 16):
    """The reference's space_split_recursive on one top-level cell. Returns (cells, parts re-ordered
    as the reference's cell_split left them)."""
    lib = load(variant)
    parts = np.ascontiguousarray(parts_u8, dtype=np.uint8).copy()
    n = parts.size // cfg.layout.size
    loc = np.ascontiguousarray(loc, dtype=np.float64)
    width = np.ascontiguousarray(width, dtype=np.float64)
    cells = np.zeros(max_cells, dtype=abi.cell_dtype())
    nc = lib.swiftref_space_split(C.byref(cfg), C.byref(step), parts.ctypes.data, n, loc.ctypes.data,
                                  width.ctypes.data, cells.ctypes.data, max_cells)
    if nc < 0:
        raise RuntimeError(f"swiftref_space_split failed ({nc})")
    return cells[:nc].copy(), parts


class Reference:
    """One reference 'engine' over a fixed tree and particle array."""

    def __init__(self, variant, cfg, step, cells, top, parts_u8):
        self.lib = load(variant)
        self.variant = variant
        self.nparts = parts_u8.size // cfg.layout.size
        self.ncells = cells.shape[0]
        self.cfg = cfg
        cells = np.ascontiguousarray(cells)
        top = np.ascontiguousarray(top, dtype=np.int32)
        self.h = self.lib.swiftref_create(C.byref(cfg), C.byref(step), cells.ctypes.data, cells.shape[0],
                                          top.ctypes.data, top.shape[0], parts_u8.ctypes.data, self.nparts)
        if not self.h:
            raise RuntimeError("swiftref_create failed (scheme/layout mismatch?)")
        self._cells = cells

    def run(self, mask=abi.PHASE_ALL, threads=1):
        sec = np.zeros(7)
        self.lib.swiftref_run(self.h, mask, threads, sec.ctypes.data)
        return dict(zip(("sort", "density", "ghost", "gradient", "extra_ghost", "force", "end_force"), sec))

    def set_parts(self, parts_u8):
        self.lib.swiftref_set_parts(self.h, parts_u8.ctypes.data)

    def parts(self):
        out = np.zeros(self.nparts * self.cfg.layout.size, dtype=np.uint8)
        self.lib.swiftref_get_parts(self.h, out.ctypes.data)
        return out

    def cells(self):
        out = self._cells.copy()
        self.lib.swiftref_get_cells(self.h, out.ctypes.data)
        return out

    def counts(self):
        nd = np.zeros(self.nparts, np.int32); ng = np.zeros_like(nd); nf = np.zeros_like(nd)
        if self.lib.swiftref_get_counts(self.h, nd.ctypes.data, ng.ctypes.data, nf.ctypes.data) != 0:
            return None
        return nd, ng, nf

    def timesteps(self):
        """hydro_compute_timestep of every particle (the reference's own function)."""
        dt = np.zeros(self.nparts, np.float32)
        self.lib.swiftref_get_timesteps(self.h, dt.ctypes.data)
        return dt

    def xpart_layout(self):
        X = abi.XpartLayout()
        self.lib.swiftref_xpart_layout(C.byref(X))
        return X

    def set_xparts(self, xparts_u8):
        self.lib.swiftref_set_xparts(self.h, xparts_u8.ctypes.data)

    def xparts(self):
        out = np.zeros(self.nparts * self.xpart_layout().size, dtype=np.uint8)
        self.lib.swiftref_get_xparts(self.h, out.ctypes.data)
        return out

    def drift(self, ti_old, minimal_internal_energy=0.0, init_particles=1):
        """The reference's cell_drift_part on every local top-level cell, from ti_old to ti_current."""
        self.lib.swiftref_run_drift(self.h, int(ti_old), float(minimal_internal_energy), int(init_particles))

    def kick(self, which, minimal_internal_energy=0.0):
        """The reference's runner_do_kick1 (which=1) / runner_do_kick2 (which=2) on every local top-level cell."""
        self.lib.swiftref_run_kick(self.h, int(which), float(minimal_internal_energy))

    def wakeup_offset(self):
        return int(self.lib.swiftref_wakeup_offset())

    def limiter(self, threads=1):
        """The reference's runner_dosub_{self,pair}1_limiter over the density tasks."""
        self.lib.swiftref_run_limiter(self.h, int(threads))

    def sort(self, cell, sid):
        n = int(self._cells["count"][cell])
        d = np.zeros(n, np.float32); i = np.zeros(n, np.int32)
        r = self.lib.swiftref_get_sort(self.h, cell, sid, d.ctypes.data, i.ctypes.data)
        return (d, i) if r >= 0 else None

    def close(self):
        if self.h:
            self.lib.swiftref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
