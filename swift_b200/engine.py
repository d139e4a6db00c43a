"""Thin host-side mirror of the C ABI: `SwiftGPU` owns one libswiftgpu handle.

Method names follow the reference's task types (sort, density, ghost,
gradient, extra_ghost, force, end_force; src/runner_main.c:214-375). Every
call goes through the C ABI of libswiftgpu.so; failures raise RuntimeError
with swiftgpu_last_error(), like the reference's error() macro.
"""
import ctypes as C

import numpy as np

from . import abi


def nccl_unique_id():
    """128-byte NCCL id (call on rank 0, broadcast to the other ranks)."""
    lib = abi.load()
    buf = (C.c_char * 128)()
    if lib.swiftgpu_nccl_unique_id(C.addressof(buf)) != 0:
        msg = lib.swiftgpu_last_error(None)
        raise RuntimeError(f"swiftgpu_nccl_unique_id failed: {msg.decode() if msg else ''}")
    return bytes(buf)


class SwiftGPU:
    def __init__(self, cfg):
        self.lib = abi.load()
        self.cfg = cfg
        self.h = abi.VP()
        rc = self.lib.swiftgpu_init(C.byref(self.h), C.byref(cfg))
        if rc != 0:
            msg = self.lib.swiftgpu_last_error(None)
            raise RuntimeError(f"swiftgpu_init failed ({rc}): {msg.decode() if msg else ''}")
        self.nparts = 0
        self.ncells = 0

    def _ck(self, rc, what):
        if rc != 0:
            msg = self.lib.swiftgpu_last_error(self.h)
            raise RuntimeError(f"{what} failed: {msg.decode() if msg else rc}")

    def upload_cells(self, cells, top):
        self._cells = np.ascontiguousarray(cells)
        top = np.ascontiguousarray(top, dtype=np.int32)
        self.ncells = self._cells.shape[0]
        self._ck(self.lib.swiftgpu_upload_cells(self.h, self._cells.ctypes.data, self.ncells,
                                                top.ctypes.data, top.shape[0]), "upload_cells")

    def upload_parts(self, parts_u8):
        assert parts_u8.dtype == np.uint8 and parts_u8.flags.c_contiguous
        self.nparts = parts_u8.size // self.cfg.layout.size
        self._ck(self.lib.swiftgpu_upload_parts(self.h, parts_u8.ctypes.data, self.nparts), "upload_parts")

    def upload_parts_ptr(self, host_ptr, nparts):
        self.nparts = nparts
        self._ck(self.lib.swiftgpu_upload_parts(self.h, host_ptr, nparts), "upload_parts")

    def upload_parts_local(self, host_ptr, nlocal, ntotal):
        """Only the rank's own particles ([0, nlocal) of the host array); the proxies arrive by the halo exchange."""
        self.nparts = ntotal
        self.nlocal = nlocal
        self._ck(self.lib.swiftgpu_upload_parts_local(self.h, host_ptr, nlocal, ntotal), "upload_parts_local")

    def download_parts_local(self, host_ptr):
        self._ck(self.lib.swiftgpu_download_parts_local(self.h, host_ptr, self.nlocal), "download_parts_local")

    def upload_parts_device(self, dev_ptr, nparts):
        self.nparts = nparts
        self._ck(self.lib.swiftgpu_upload_parts_device(self.h, dev_ptr, nparts), "upload_parts_device")

    def set_step(self, step):
        self._ck(self.lib.swiftgpu_set_step(self.h, C.byref(step)), "set_step")

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._ck(self.lib.swiftgpu_set_stream(self.h, cuda_stream), "set_stream")

    def halo_setup(self, unique_id):
        """unique_id: the 128 bytes of swiftgpu_nccl_unique_id() obtained on rank 0."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.swiftgpu_halo_setup(self.h, C.addressof(buf)), "halo_setup")

    def halo_exchange(self, phase):
        self._ck(self.lib.swiftgpu_halo_exchange(self.h, phase), "halo_exchange")

    def run_sort(self):
        self._ck(self.lib.swiftgpu_run_sort(self.h), "run_sort")

    def run_density(self):
        self._ck(self.lib.swiftgpu_run_density(self.h), "run_density")

    def run_ghost(self):
        self._ck(self.lib.swiftgpu_run_ghost(self.h), "run_ghost")

    def run_gradient(self):
        self._ck(self.lib.swiftgpu_run_gradient(self.h), "run_gradient")

    def run_extra_ghost(self):
        self._ck(self.lib.swiftgpu_run_extra_ghost(self.h), "run_extra_ghost")

    def run_force(self):
        self._ck(self.lib.swiftgpu_run_force(self.h), "run_force")

    def run_end_force(self):
        self._ck(self.lib.swiftgpu_run_end_force(self.h), "run_end_force")

    def run_step(self, mask=abi.PHASE_ALL):
        self._ck(self.lib.swiftgpu_run_step(self.h, mask), "run_step")

    def download_parts(self, out=None):
        if out is None:
            out = np.zeros(self.nparts * self.cfg.layout.size, dtype=np.uint8)
        self._ck(self.lib.swiftgpu_download_parts(self.h, out.ctypes.data, self.nparts), "download_parts")
        return out

    def download_parts_ptr(self, host_ptr):
        self._ck(self.lib.swiftgpu_download_parts(self.h, host_ptr, self.nparts), "download_parts")

    def download_parts_device(self, dev_ptr):
        self._ck(self.lib.swiftgpu_download_parts_device(self.h, dev_ptr, self.nparts), "download_parts_device")

    def download_cells(self):
        out = self._cells.copy()
        self._ck(self.lib.swiftgpu_download_cells(self.h, out.ctypes.data, self.ncells), "download_cells")
        return out

    def download_timestep(self):
        """hydro_compute_timestep of the active particles (-1 for inactive ones)."""
        dt = np.empty(self.nparts, dtype=np.float32)
        self._ck(self.lib.swiftgpu_download_timestep(self.h, dt.ctypes.data, self.nparts), "download_timestep")
        return dt

    # ---- drift on the device ----
    def upload_xparts(self, xlayout, xparts_u8):
        """struct xpart[] of the caller (one per uploaded part), layout = offsetof() of x_diff, x_diff_sort, v_full."""
        xparts_u8 = np.ascontiguousarray(xparts_u8, dtype=np.uint8)
        self._xlayout = xlayout
        self._nx = xparts_u8.size // xlayout.size
        self._ck(self.lib.swiftgpu_upload_xparts(self.h, C.byref(xlayout), xparts_u8.ctypes.data, self._nx),
                 "upload_xparts")

    def download_xparts(self):
        out = np.zeros(self._nx * self._xlayout.size, dtype=np.uint8)
        self._ck(self.lib.swiftgpu_download_xparts(self.h, out.ctypes.data, self._nx), "download_xparts")
        return out

    def run_drift(self, dt_drift, dt_kick_hydro=None, dt_therm=None, minimal_internal_energy=0.0, init_particles=1):
        """cell_drift_part (force = 1) over every local cell; the three factors as cell_drift.c:236-252 derives them."""
        a = abi.DriftArgs(float(dt_drift), float(dt_drift if dt_kick_hydro is None else dt_kick_hydro),
                          float(dt_drift if dt_therm is None else dt_therm), float(minimal_internal_energy),
                          int(init_particles))
        self._ck(self.lib.swiftgpu_run_drift(self.h, C.byref(a)), "run_drift")

    def run_kick(self, which, minimal_internal_energy=0.0):
        """runner_do_kick1 (which=1) / runner_do_kick2 (which=2, + hydro_reset_predicted_values) on the device."""
        self._ck(self.lib.swiftgpu_run_kick(self.h, int(which), float(minimal_internal_energy)), "run_kick")

    def run_limiter(self, wakeup_offset):
        """runner_dosub_{self,pair}1_limiter: wake-up flags (limiter_data.wakeup at `wakeup_offset` of struct part)."""
        self._ck(self.lib.swiftgpu_run_limiter(self.h, int(wakeup_offset)), "run_limiter")

    def download_counts(self):
        nd = np.zeros(self.nparts, np.int32)
        ng = np.zeros_like(nd)
        nf = np.zeros_like(nd)
        self._ck(self.lib.swiftgpu_download_counts(self.h, nd.ctypes.data, ng.ctypes.data, nf.ctypes.data,
                                                   self.nparts), "download_counts")
        return nd, ng, nf

    def download_sort(self, cell, sid):
        """(indices, key_min, key_max) of the sorted array of (cell, sid)."""
        n = int(self._cells["count"][cell])
        idx = np.zeros(n, np.int32)
        kmin, kmax = C.c_float(), C.c_float()
        self._ck(self.lib.swiftgpu_download_sort(self.h, cell, sid, idx.ctypes.data, C.addressof(kmin),
                                                 C.addressof(kmax)), "download_sort")
        return idx, kmin.value, kmax.value

    def stats(self):
        s = abi.Stats()
        self._ck(self.lib.swiftgpu_get_stats(self.h, C.byref(s)), "get_stats")
        return s

    def close(self):
        if self.h:
            self.lib.swiftgpu_destroy(self.h)
            self.h = abi.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
