"""ctypes wrapper of oracle/libswiftport_<scheme>.so - the plain-C restatement
of the reference path (oracle/swift_port.c). TEST INFRASTRUCTURE."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from swift_b200 import abi  # noqa: E402

VP = C.c_void_p
_libs = {}


def load(scheme_name):
    if scheme_name not in _libs:
        path = os.path.join(ROOT, "oracle", f"libswiftport_{scheme_name}.so")
        lib = C.CDLL(path, mode=os.RTLD_NOW | os.RTLD_LOCAL)
        lib.port_create.restype = VP
        lib.port_create.argtypes = [C.POINTER(abi.Config), C.POINTER(abi.Step), VP, C.c_int, VP, C.c_int, VP, C.c_longlong]
        lib.port_destroy.argtypes = [VP]
        lib.port_run.argtypes = [VP, C.c_uint]
        lib.port_get_parts.argtypes = [VP, C.c_uint, VP]
        lib.port_get_cells.argtypes = [VP, VP]
        lib.port_get_counts.argtypes = [VP, VP, VP, VP]
        lib.port_ghost_iterations.argtypes = [VP]
        lib.port_ghost_redo.argtypes = [VP, VP]
        lib.port_get_gross.argtypes = [VP] * 10
        lib.port_kernel_deval.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        lib.port_sub_pairs.argtypes = [C.c_int, VP, VP]
        _libs[scheme_name] = lib
    return _libs[scheme_name]


class Port:
    def __init__(self, scheme_name, cfg, step, cells, top, parts_u8):
        self.lib = load(scheme_name)
        self.cfg = cfg
        self.nparts = parts_u8.size // cfg.layout.size
        self._cells = np.ascontiguousarray(cells)
        top = np.ascontiguousarray(top, dtype=np.int32)
        self._parts_in = parts_u8.copy()
        self.h = self.lib.port_create(C.byref(cfg), C.byref(step), self._cells.ctypes.data, self._cells.shape[0],
                                      top.ctypes.data, top.shape[0], parts_u8.ctypes.data, self.nparts)
        if not self.h:
            raise RuntimeError("port_create failed")
        self.mask = 0

    def run(self, mask=abi.PHASE_ALL):
        self.mask |= mask
        return self.lib.port_run(self.h, mask)

    def parts(self):
        out = self._parts_in.copy()
        self.lib.port_get_parts(self.h, self.mask, out.ctypes.data)
        return out

    def cells(self):
        out = self._cells.copy()
        self.lib.port_get_cells(self.h, out.ctypes.data)
        return out

    def counts(self):
        nd = np.zeros(self.nparts, np.int32); ng = np.zeros_like(nd); nf = np.zeros_like(nd)
        self.lib.port_get_counts(self.h, nd.ctypes.data, ng.ctypes.data, nf.ctypes.data)
        return nd, ng, nf

    def gross(self):
        """dict of the un-cancelled sums behind a_hydro, u_dt, h_dt, div_v, rho_dh, laplace_u and the
        squared kernel-noise sums of the first, second and last (swift_port.c: edge_weight)."""
        names = ("a_hydro", "u_dt", "h_dt", "div_v", "rho_dh", "laplace_u", "a_hydro_sq", "u_dt_sq", "laplace_u_sq")
        arrs = [np.zeros(self.nparts, np.float32) for _ in names]
        self.lib.port_get_gross(self.h, *[a.ctypes.data for a in arrs])
        return dict(zip(names, (a.astype(np.float64) for a in arrs)))

    def ghost_iterations(self):
        return self.lib.port_ghost_iterations(self.h)

    def ghost_redo(self):
        """particles that entered re-run k of the ghost, k = 0..31"""
        out = (C.c_longlong * 32)()
        self.lib.port_ghost_redo(self.h, out)
        return [int(v) for v in out]

    def close(self):
        if self.h:
            self.lib.port_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
