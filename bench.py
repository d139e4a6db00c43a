#!/usr/bin/env python
"""bench.py - SPH pair interactions/s of the hot path (sort -> density -> ghost
-> [gradient -> extra ghost] -> force -> end_force) on a synthetic gas box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload sphenix512|sedov128|clustered256|sphenix1024a5|uniform32|...]

One "step" = one pass of the whole path over the box, i.e. what SWIFT's task
graph runs between drift and kick2 for an all-active step 0
(engine_init_particles, engine.c:2430). Workload: BASELINE.json configs[3],
the 512^3 synthetic SPHENIX gas box the north_star quotes its target on - on
one GPU at N=1, and the SAME box split over 2/4/8 GPUs by the reference's
partition_uniform_grid (strong scaling, NCCL halo exchange) at N>1.

Printed JSON line (driver contract):
  value   directed pair interactions / s with the AoS particle array already
          resident in HBM (AoS->SoA transpose, all phases and SoA->AoS included)
  e2e     same metric through the C ABI with pinned HOST buffers
          (swiftgpu_upload_parts -> run_step -> swiftgpu_download_parts)
  roofline      dominant neighbour-loop kernel: algorithmic FP32 flops of the
                interactions it evaluated / its CUDA-event time, against the
                FP32-pipe peak (non-tensor pairwise FP32 work)
  roofline_hbm  same kernel against the measured HBM copy bandwidth
  cpu_baseline  the unmodified reference (oracle/_ref, all host threads) on a
                bounded sample of the same workload

The interactions counted are the `useful' ones: for every active particle the
number of runner_iact_nonsym_<loop> calls of its FINAL density pass, gradient
pass and force pass (identical on both arms because neighbour sets are
bit-exact); re-runs of the ghost are extra work on both arms and are reported
separately (`interactions_incl_ghost_reruns`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

# algorithmic FP32 flops per directed interaction (FMA = 2)
FLOPS = {"density": 64, "gradient": 45, "force_minimal": 110, "force_gadget2": 110, "force_sphenix": 135}
# algorithmic HBM bytes per ACTIVE particle of each loop (SURVEY 8d "field traffic")
BYTES = {"density": 46 + 32, "gradient": 60 + 12, "force": 78 + 25}

WORKLOADS = {
    # name: (scheme, L, generator, active fraction)
    # "brick": ONE periodic unit box of L^3 jittered-lattice particles (counter-based jitter, +-0.2
    # spacing, smooth shear velocity field); with --gpus N the SAME box is split over the ranks
    # (strong scaling), every rank generating only its brick and the halo cells it holds proxies of.
    "sphenix512": ("sphenix", 512, "brick", 1.0),   # BASELINE config 3: the headline (134 M particles)
    "sphenix256": ("sphenix", 256, "brick", 1.0),
    "sphenix128": ("sphenix", 128, "brick", 1.0),
    "sphenix64": ("sphenix", 64, "brick", 1.0),
    # multi-time-step: ~5 % of the particles active (clustered in space), the rest are neighbours only
    "sphenix128a5": ("sphenix", 128, "brick", 0.05),
    "sphenix1024a5": ("sphenix", 1024, "brick", 0.05),  # BASELINE config 4 (8 GPUs)
    "uniform32": ("minimal", 32, "uniform", 1.0),   # BASELINE config 0
    "sedov64": ("gadget2", 64, "sedov", 1.0),
    "sedov128": ("gadget2", 128, "sedov", 1.0),     # BASELINE config 1
    "clustered128": ("sphenix", 128, "clustered", 1.0),
    "clustered256": ("sphenix", 256, "clustered", 1.0),  # BASELINE config 2
}
DEFAULT_WORKLOAD = "sphenix512"

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}  # partition.c:112-120 bricks


def top_grid_of(L):
    from swift_b200 import host
    tg = host.default_top_grid(L, max_top=64 if L >= 1024 else 32)
    if os.environ.get("SWIFTGPU_TOPGRID"):  # experiments: other leaf sizes
        tg = (int(os.environ["SWIFTGPU_TOPGRID"]),) * 3
    return tg


def make_workload(name, world=1, rank=0, all_active=False):
    """Builds the case of `rank`: config, step scalars, the rank's sub-tree (own top-level cells + the
    foreign ones that touch them), and the AoS particle array with the rank's own particles FIRST
    ([0, n_local)) and the proxies behind them. world > 1: strong scaling, the same box split by
    partition_uniform_grid over GRIDS[world]."""
    import util
    from swift_b200 import abi, host
    scheme, L, gen, active = WORKLOADS[name]
    sid = abi.SCHEMES[scheme]
    grid = GRIDS[world]
    if world > 1 and gen != "brick":
        raise SystemExit(f"workload {name} has no multi-GPU generator")
    cdim = top_grid_of(L)
    if gen == "uniform":
        ic = host.uniform_box(L, sid)
    elif gen == "sedov":
        ic = host.sedov_box(L, sid)
    elif gen == "clustered":
        ic = host.clustered_box(L, sid)
    else:
        ic = host.brick_box(L, sid, grid=grid, rank=rank, top=cdim[0], active_fraction=1.0 if all_active else active)
    mab = 56 if (active >= 1.0 or all_active) else 1
    c = util.make_case(scheme, ic, cdim, rank_grid=grid, rank=rank, max_active_bin=mab, pack=(world == 1))
    c.sub_tree = c.tree
    c.n_local = c.n
    if world > 1:
        sub, _, sel, is_local = host.extract_rank(c.tree, None, c.layout, rank, local_first=True)
        c.sub_tree = sub
        c.parts = host.pack_parts(c.layout, c.scheme, sub, ic)
        c.n_local = int(is_local.sum())
        c.n = int(sel.shape[0])
        assert is_local[:c.n_local].all()
    c.ids = np.asarray(ic["id"])[c.sub_tree.perm]
    c.time_bin_ic = np.asarray(ic["time_bin"])[c.sub_tree.perm]
    c.active_fraction = active
    c.synthetic_inactive = False
    if active < 1.0 and not all_active and (L >= 512 or os.environ.get("SWIFTGPU_SYNTH")):
        # Too large for the preparatory all-active step (its worklists and frame arrays would not fit
        # next to the benchmark's own buffers): the inactive neighbours carry the force-union members of
        # the unperturbed medium instead (rho = rho_bar, P = (gamma-1) u rho, c_s, f = 0, balsara = 0.5).
        # The kernels do the same work on them; parity of active subsets is tested at small size.
        rho0 = float(ic["_rho0"])
        u = host.field(c.parts, c.layout, "u").astype(np.float64)
        P = (host.HYDRO_GAMMA - 1.0) * u * rho0
        host.field(c.parts, c.layout, "rho")[:] = rho0
        host.field(c.parts, c.layout, "pressure")[:] = P.astype(np.float32)
        host.field(c.parts, c.layout, "soundspeed")[:] = np.sqrt(host.HYDRO_GAMMA * P / rho0).astype(np.float32)
        host.field(c.parts, c.layout, "f")[:] = 0.0
        host.field(c.parts, c.layout, "balsara")[:] = 0.5
        c.synthetic_inactive = True
    del ic
    return c


def prepare_inactive_state(c, g, world):
    """Multi-time-step workloads: inactive particles are neighbours, and what a neighbour contributes
    in the force loop are the force-union members of `its last step`. Take them from an all-active
    step of the same box run here on the GPU, then apply the time bins. (c: the case built with the
    real time bins; g: a handle on c's tree with the halo set up.)"""
    from swift_b200 import abi, host
    tb = host.field(c.parts, c.layout, "time_bin")
    saved = tb.copy()
    tb[:] = 1
    step_all = __import__("util").make_step(56)
    g.set_step(step_all)
    # the tree c was built with carries ti_end_min / h_max_active of the real bins: an all-active
    # pass needs every cell active, so run it on cells with ti_end_min = ti_current
    cells = c.sub_tree.cells.copy()
    cells["ti_end_min"] = step_all.ti_current
    cells["h_max_active"] = cells["h_max"]
    g.upload_cells(cells, c.sub_tree.top)
    if world > 1:
        g.halo_setup(c.nccl_id)
    g.upload_parts(c.parts)
    g.run_step(abi.PHASE_ALL)
    out = g.download_parts().copy()
    c.parts[:] = out
    host.field(c.parts, c.layout, "time_bin")[:] = saved
    g.upload_cells(c.sub_tree.cells, c.sub_tree.top)
    if world > 1:
        g.halo_setup(c.nccl_id)
    g.set_step(c.step)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []   # (host time, csv line)
        self.stop_flag = False
        self.proc = None
        self.window = (0.0, float("inf"))

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [s for (t, s) in self.samples if self.window[0] <= t <= self.window[1] + 0.15]
        if not inside:  # timed region shorter than the sampling period: nearest samples under load
            inside = [s for (t, s) in self.samples if t >= self.window[0] - 1.0]
        for s in inside:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def ref_variant(scheme):
    """The reference build the CPU arm times: the optimised one (-O3 -ffast-math -funroll-loops
    -march=x86-64-v3, the flags of the reference's default build, m4/ax_cc_maxopt.m4:174-200) when it
    travelled with the snapshot, else the bit-stable parity build."""
    from oracle import ref
    if ref.available(scheme + "_fast"):
        return scheme + "_fast", "optimised build (-O3 -ffast-math -funroll-loops -march=x86-64-v3)"
    if ref.available(scheme):
        return scheme, "parity build (-O3 -ffp-contract=off, no -ffast-math)"
    return None, "C restatement (oracle/swift_port.c)"


CPU_SAMPLE = {"sphenix512": "sphenix128", "sphenix256": "sphenix128", "sphenix1024a5": "sphenix128a5",
              "clustered256": "clustered128"}


def cpu_reference_run(workload, steps, warmup, sample_name=None):
    """Times the UNMODIFIED reference (oracle/_ref) - or the C port when the reference library did not
    travel - on a bounded sample of the workload with all host threads. The sample is the same
    generator at the same particle density (per-particle work is independent of the box size).
    Returns (info dict, seconds per step, useful interactions, particles)."""
    import util
    from oracle import port, ref
    from swift_b200 import abi, host
    scheme = WORKLOADS[workload][0]
    if sample_name is None:
        sample_name = CPU_SAMPLE.get(workload, workload)
    c = make_workload(sample_name)
    cores = os.cpu_count() or 1
    if WORKLOADS[sample_name][3] < 1.0:
        # inactive neighbours need the force-union members of their last step: an all-active pass first
        c_all = make_workload(sample_name, all_active=True)
        pa = port.Port(scheme, c_all.cfg, c_all.step, c_all.tree.cells, c_all.tree.top, c_all.parts)
        pa.run(abi.PHASE_ALL)
        parts = pa.parts()
        pa.close()
        host.field(parts, c.layout, "time_bin")[:] = c.time_bin_ic
        c.parts = parts
    # count the useful interactions once with the port (checker role, untimed)
    p = port.Port(scheme, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
    p.run(abi.PHASE_ALL)
    nd, ng, nf = p.counts()
    useful = int(nd.sum()) + int(ng.sum()) + int(nf.sum())
    p.close()
    variant, how = ref_variant(scheme)
    kind = "reference" if variant else "port"
    times = []
    if kind == "reference":
        o = ref.Reference(variant, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
        for it in range(warmup + steps):
            o.set_parts(c.parts)
            t0 = time.perf_counter()
            o.run(abi.PHASE_ALL, threads=cores)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
        o.close()
    else:
        for it in range(warmup + steps):
            q = port.Port(scheme, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
            t0 = time.perf_counter()
            q.run(abi.PHASE_ALL)
            dt = time.perf_counter() - t0
            q.close()
            if it >= warmup:
                times.append(dt)
    sec = float(np.mean(times))
    same = sample_name == workload
    info = {"value": useful / sec, "unit": "interactions/s", "cores": cores, "kind": kind,
            "sample": f"{sample_name}: {c.n} particles ({scheme}), " +
                      ("the benchmarked configuration itself" if same else
                       f"same generator and particle density as {workload} (per-particle work is size-independent; "
                       f"NOT the same configuration: the ratio to the GPU line is an extrapolation)") +
                      f"; {useful} useful directed interactions per step, {sec * 1e3:.1f} ms/step wall-clock, "
                      f"mean of {len(times)} step(s), {cores} pthreads, {how}",
            "same_config": same}
    return info, sec, useful, c.n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    info, sec, useful, n = cpu_reference_run(args.workload, max(1, min(args.steps, 3)), max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "SPH pair interactions/s (density+gradient+force)",
            "value": info["value"], "unit": "interactions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scheme": WORKLOADS[args.workload][0],
                       "note": "reference CPU path on a bounded sample of the workload (cpu_baseline.sample)"},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def newest_traffic(workload, kernel, loop):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the newest
    committed ncu --set full capture of this command (profiles/*_traffic.json, written by
    scripts/profile_summary.py from the .ncu-rep); None when no capture of this workload exists."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json"))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        v = d.get(workload, {}).get(f"{kernel}:{loop}")
        if v is not None:
            best = (float(v), os.path.basename(f))
    return best


def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: pin the process (and with it the pages of its pinned host buffers, which are
    placed by first touch) to the CPUs of the NUMA node the rank's GPU hangs off, so that the eight
    ranks' host<->device copies do not all cross one socket's memory controller. Returns the node or
    None when the topology cannot be read (nothing is changed then)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def halo_parity_check(world, rank, local_rank, nccl_id):
    """Multi-GPU correctness seen on the driver box: a 64^3 SPHENIX box split over the ranks (halo
    exchanges inside run_step, local particles only across the host boundary) against the SAME box run
    on one GPU by this rank alone; compared per particle id on the rank's local particles."""
    import util
    from swift_b200 import abi, host
    from swift_b200.engine import SwiftGPU
    name = "sphenix64"
    cm = make_workload(name, world, rank)
    cm.cfg.device = local_rank
    g = SwiftGPU(cm.cfg)
    g.upload_cells(cm.sub_tree.cells, cm.sub_tree.top)
    g.set_step(cm.step)
    g.halo_setup(nccl_id)
    pin = np.ascontiguousarray(cm.parts)
    g.upload_parts_local(pin.ctypes.data, cm.n_local, cm.n)
    g.run_step(abi.PHASE_ALL)
    out = np.zeros(cm.n_local * cm.layout.size, np.uint8)
    g.download_parts_local(out.ctypes.data)
    ndm, ngm, nfm = g.download_counts()
    g.close()
    c1 = make_workload(name, 1, 0)
    c1.cfg.device = local_rank
    g1 = SwiftGPU(c1.cfg)
    g1.upload_cells(c1.tree.cells, c1.tree.top)
    g1.set_step(c1.step)
    g1.upload_parts(c1.parts)
    g1.run_step(abi.PHASE_ALL)
    ref = g1.download_parts()
    nd1, ng1, nf1 = g1.download_counts()
    g1.close()
    # match by particle id
    order = np.argsort(c1.ids)
    pos = order[np.searchsorted(c1.ids[order], cm.ids[:cm.n_local])]
    size = cm.layout.size
    ref_loc = np.ascontiguousarray(ref.reshape(-1, size)[pos]).reshape(-1)
    rep = util.parity_report(out, ref_loc, cm.layout, "sphenix")
    worst = max(v for k, v in rep.items() if k in ("h", "rho", "pressure", "a_hydro", "u_dt", "h_dt", "v_sig"))
    counts_ok = bool(np.array_equal(ndm[:cm.n_local], nd1[pos]) and np.array_equal(ngm[:cm.n_local], ng1[pos]) and
                     np.array_equal(nfm[:cm.n_local], nf1[pos]))
    return {"box": name, "n_local": int(cm.n_local), "max_rel_err": float(worst), "flips": int(rep["flips"]),
            "counts_identical": counts_ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="swiftgpu")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-halo-parity", action="store_true")
    ap.add_argument("--no-resident", action="store_true", help="skip the device-resident drift + step leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (and its pinned host copies): memory-bound configurations")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = DEFAULT_WORKLOAD
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import util  # noqa: F401
    from swift_b200 import abi, host
    from swift_b200.engine import SwiftGPU

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libswiftgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    args.warmup = max(args.warmup, 3)
    scheme, L, gen, active = WORKLOADS[args.workload]
    if world not in GRIDS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")

    def new_nccl_id():
        from swift_b200.engine import nccl_unique_id
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return idt.cpu().numpy().tobytes()

    halo_parity = None
    if world > 1 and not args.no_halo_parity:
        hp = halo_parity_check(world, rank, local_rank, new_nccl_id())
        t = torch.tensor([hp["max_rel_err"], float(hp["flips"]), 0.0 if hp["counts_identical"] else 1.0],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        halo_parity = {"box": hp["box"], "ranks": world, "max_rel_err": float(t[0]), "flips_max_per_rank": int(t[1]),
                       "counts_identical": bool(t[2] == 0.0),
                       "what": "local particles of every rank (NCCL halos, local-only host copies) vs the same box on one GPU"}

    c = make_workload(args.workload, world, rank)
    c.cfg.device = local_rank
    n = c.n
    n_local = c.n_local
    psize = c.layout.size

    g = SwiftGPU(c.cfg)
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    g.upload_cells(c.sub_tree.cells, c.sub_tree.top)
    g.set_step(c.step)
    if world > 1:
        c.nccl_id = new_nccl_id()
        g.halo_setup(c.nccl_id)
    if active < 1.0 and not c.synthetic_inactive:
        prepare_inactive_state(c, g, world)

    if args.no_e2e:
        host_in = torch.from_numpy(c.parts)
        host_out = None
        dev_in = host_in.to("cuda", non_blocking=False)
        dev_out = dev_in  # the download of the device-resident leg goes back into the same buffer
    else:
        host_in = torch.from_numpy(c.parts).pin_memory()
        host_out = torch.empty(n_local * psize, dtype=torch.uint8).pin_memory()
        dev_in = host_in.to("cuda", non_blocking=False)
        dev_out = torch.empty_like(dev_in)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        g.upload_parts_device(dev_in.data_ptr(), n)
        g.run_step(abi.PHASE_ALL)
        if dev_out is not dev_in:
            g.download_parts_device(dev_out.data_ptr())

    def step_e2e():
        # the rank's OWN particles cross the host boundary; proxies arrive over NVLink
        g.upload_parts_local(host_in.data_ptr(), n_local, n)
        g.run_step(abi.PHASE_ALL)
        g.download_parts_local(host_out.data_ptr())

    with torch.cuda.stream(stream):
        # ---- device-resident timing ----
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(args.warmup):
            step_device()
        l0 = g.stats().n_launches
        barrier()
        t_begin = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phase_ms = {k: 0.0 for k in ("sort", "density", "ghost", "gradient", "extra_ghost", "force", "end_force")}
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
            st = g.stats()
            for k in phase_ms:
                phase_ms[k] += getattr(st, "ms_" + k)
        e1.record(stream)
        barrier()
        sampler.window = (t_begin, time.time())
        ms_total = e0.elapsed_time(e1)
        st = g.stats()
        launches = st.n_launches - l0
        nd, ng, nf = g.download_counts()
        nd, ng, nf = nd[:n_local], ng[:n_local], nf[:n_local]
        useful = int(nd.sum()) + int(ng.sum()) + int(nf.sum())
        executed = int(st.n_density + st.n_gradient + st.n_force)

        # ---- end-to-end timing (pinned host buffers through the C ABI) ----
        ms_e2e = float("nan")
        if not args.no_e2e:
            step_e2e()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for _ in range(args.steps):
                step_e2e()
            f1.record(stream)
            barrier()
            ms_e2e = f0.elapsed_time(f1)
        # ---- device-resident leg: drift + step without the particles crossing the host boundary ----
        resident = None
        if world == 1 and not args.no_resident:
            X = abi.XpartLayout(48, 0, 12, 24, 36)  # synthetic struct xpart: x_diff, x_diff_sort, v_full, u_full + padding
            xp = np.zeros((n, 48), np.uint8)
            xp[:, 24:36] = np.ascontiguousarray(host.field(c.parts, c.layout, "v").reshape(n, 3),
                                                dtype=np.float32).view(np.uint8).reshape(n, 12)
            xp[:, 36:40] = np.ascontiguousarray(host.field(c.parts, c.layout, "entropy" if scheme == "gadget2" else "u"),
                                                dtype=np.float32).view(np.uint8).reshape(n, 4)
            g.upload_parts_device(dev_in.data_ptr(), n)
            g.upload_xparts(X, xp.ravel())
            del xp
            g.run_step(abi.PHASE_ALL)  # a_hydro, h_dt, u_dt to drift with
            # the global time-step an engine would take: a fraction of the smallest CFL time-step of the
            # box (hydro_compute_timestep of the end_force epilogue), at most ~1e-3 particle spacings.
            # The kicks derive their half step from the particles' time bin: (ti_step / 2) * time_base
            # with ti_step = 4 for the active bin 1 (timeline.h:59) - so time_base carries dt
            cfl = g.download_timestep()
            cfl = cfl[cfl > 0]
            dt = min(0.25 * float(cfl.min()) if cfl.size else 1.0, 1e-3 / L / 0.05)
            step_res = abi.Step()
            for f, _ in abi.Step._fields_:
                setattr(step_res, f, getattr(c.step, f))
            step_res.time_base = dt / 4.0
            g.set_step(step_res)

            def step_resident():
                # fixed-time-step leapfrog (runner_do_kick2, runner_do_kick1, cell_drift_part, the hydro step)
                g.run_kick(2)
                g.run_kick(1)
                g.run_drift(dt, init_particles=1)
                g.run_step(abi.PHASE_ALL)
            step_resident()
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(args.steps):
                step_resident()
            r1.record(stream)
            barrier()
            ms_res = r0.elapsed_time(r1) / args.steps
            rd, rg, rf = g.download_counts()
            useful_res = int(rd.sum()) + int(rg.sum()) + int(rf.sum())
            resident = {"value": useful_res / (ms_res * 1e-3), "unit": "interactions/s", "ms_per_step": ms_res,
                        "interactions_per_step": useful_res, "dt_drift": dt,
                        "ghost_iterations": int(g.stats().ghost_iterations),
                        "what": "fixed-time-step leapfrog on the device-resident state: swiftgpu_run_kick(2), "
                                "swiftgpu_run_kick(1), swiftgpu_run_drift, swiftgpu_run_step per step; no particle crosses "
                                "the host boundary (SURVEY 8f rows 2 and 4)"}
        sampler.stop()

    # max over ranks
    t = torch.tensor([ms_total, ms_e2e], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(useful), float(executed), float(n_local), float(nf.sum()), float(nd.sum())],
                       device="cuda", dtype=torch.float64)
    ph = torch.tensor([phase_ms[k] for k in sorted(phase_ms)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    useful_all, executed_all, n_all = float(tot[0]), float(tot[1]), float(tot[2])
    nf_all, nd_all = float(tot[3]), float(tot[4])
    phase_ms = {k: float(v) for k, v in zip(sorted(phase_ms), ph)}
    ms_step = ms_total / args.steps
    value = useful_all / (ms_step * 1e-3)
    e2e_value = useful_all / (ms_e2e / args.steps * 1e-3) if ms_e2e == ms_e2e else None

    # ---- roofline of the dominant kernel (CUDA-event phase times of the library, max over ranks) ----
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    props = torch.cuda.get_device_properties(local_rank)
    sms = props.multi_processor_count
    fp32_peak = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12 * world  # TFLOP/s at max SM clock, all GPUs
    ms_force = phase_ms["force"] / args.steps
    ms_density = phase_ms["density"] / args.steps
    fl_force = FLOPS["force_" + scheme]
    cand = {"density": st.t_density, "gradient": st.t_gradient, "force": st.t_force}
    # (CUDA-event ms of the phase = one launch of the loop kernel, algorithmic flops, algorithmic bytes)
    kernels = {
        "force": (ms_force, nf_all * fl_force, n_all * active * BYTES["force"]),
        "density": (ms_density, nd_all * FLOPS["density"], n_all * active * BYTES["density"]),
    }
    dom = max(kernels, key=lambda k: kernels[k][0])
    kms, kflops, kbytes = kernels[dom]
    achieved_tf = kflops / (kms * 1e-3) / 1e12
    achieved_gbs = kbytes / (kms * 1e-3) / 1e9
    kname = {"pipe": "k_pipe", "tile": "k_tile", "cta": "k_cta", "warp": "k_loop"}.get(
        os.environ.get("SWIFTGPU_LOOPS", "pipe"), "k_pipe")
    tr = newest_traffic(args.workload, kname, dom) if world == 1 else None
    traffic = tr[0] if tr else None
    roofline = {"bound": "fp32", "kernel": (kname + "<FORCE,%s>" % scheme) if dom == "force" else (kname + "<DENSITY>"),
                "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
                "traffic": traffic,
                "traffic_unit": ("bytes/launch (ncu dram read+write, profiles/%s)" % tr[1]) if tr else None,
                "algorithmic_bytes": kbytes,
                "peak_source": f"{world} x {sms} SMs x 128 FP32 lanes x 2 x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz, {peak_src})",
                "ms_per_launch": kms, "flops_per_interaction": fl_force if dom == "force" else FLOPS["density"],
                "candidates_per_hit": (cand[dom] / max(1.0, float(nf.sum() if dom == "force" else nd.sum())))}
    roofline_hbm = {"bound": "hbm", "kernel": roofline["kernel"], "achieved": achieved_gbs, "peak": hbm_peak * world,
                    "unit": "GB/s", "frac": achieved_gbs / (hbm_peak * world), "traffic": traffic,
                    "peak_source": f"{world} x MEASURED_PEAKS.json hbm_gbs ({peak_src})"}

    line = {
        "metric": "SPH pair interactions/s (density+gradient+force)", "value": value, "unit": "interactions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload if world == 1 else
                   "%s split over %s bricks (strong scaling: the same %d^3 box, partition_uniform_grid)" % (
                       args.workload, "x".join(map(str, GRIDS[world])), L),
                   "scheme": scheme, "particles": int(n_all), "particles_per_gpu_incl_halo": int(n),
                   "active_fraction": active,
                   "inactive_state": ("synthetic (unperturbed medium)" if c.synthetic_inactive else "from an all-active step") if active < 1.0 else None,
                   "l2": "inputs larger than L2 (%.0f MB AoS + SoA state per GPU and step)" % (n * psize / 1e6),
                   "top_grid": list(top_grid_of(L)),
                   "ghost_iterations": int(st.ghost_iterations)},
        "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": n_local * psize,
                "d2h_bytes_per_step": n_local * psize, "ms_per_step": (ms_e2e / args.steps) if e2e_value else None,
                "note": "per GPU: the rank's own particles through the C ABI (pinned host AoS in and out); proxies by NCCL"
                        + ("; rank 0 bound to NUMA node %d of its GPU" % numa if numa is not None else "")},
        "gpu_launches": int(launches),
        "interactions_per_step": useful_all, "interactions_incl_ghost_reruns": executed_all,
        "phase_ms": {k: v / args.steps for k, v in phase_ms.items()},
        "host_syncs_per_step": float(st.n_host_syncs) / max(1, args.steps + args.warmup + 1 + args.steps),
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "clocks": sampler.summary(),
    }
    if halo_parity is not None:
        line["halo_parity"] = halo_parity
    if resident is not None:
        line["resident"] = resident
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            info, _, _, _ = cpu_reference_run(args.workload, 1, 0)
            line["cpu_baseline"] = info
        except Exception as e:  # the baseline is a reported extra, never the product path
            line["cpu_baseline"] = {"value": None, "unit": "interactions/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
