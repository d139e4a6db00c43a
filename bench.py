#!/usr/bin/env python
"""bench.py - SPH pair interactions/s of the hot path (sort -> density -> ghost
-> [gradient -> extra ghost] -> force -> end_force) on a synthetic gas box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload sedov128|sphenix256|uniform32|...]

One "step" = one pass of the whole path over the box, i.e. what SWIFT's task
graph runs between drift and kick2 for an all-active step 0
(engine_init_particles, engine.c:2430). Workload at N=1: BASELINE.json
configs[1], the Sedov blast on a perturbed 128^3 lattice, Gadget2 SPH.

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
    # name: (scheme, L, generator, top grid)
    "uniform32": ("minimal", 32, "uniform"),
    "sedov64": ("gadget2", 64, "sedov"),
    "sedov128": ("gadget2", 128, "sedov"),
    "sphenix128": ("sphenix", 128, "jitter"),
    "sphenix256": ("sphenix", 256, "jitter"),
    "sphenix512": ("sphenix", 512, "jitter"),  # 134 M particles: ~110 GB of HBM with the bench's device copies
    # multi-time-step: ~5 % of the particles active (clustered in space), the rest are neighbours only
    "sphenix128a5": ("sphenix", 128, "active5"),
    "clustered128": ("sphenix", 128, "clustered"),
    "clustered256": ("sphenix", 256, "clustered"),
}


GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}  # partition.c:112-120 bricks


def make_workload(name, world=1, rank=0):
    """Builds the case of `rank`. world > 1: WEAK scaling - the box is made of
    one L^3 brick per GPU (brick grid GRIDS[world]); the rank keeps its brick
    and a proxy copy of the foreign top-level cells that touch it."""
    import util
    from swift_b200 import abi, host
    scheme, L, gen = WORKLOADS[name]
    sid = abi.SCHEMES[scheme]
    bricks = GRIDS[world]
    if world > 1 and gen in ("uniform", "clustered"):
        raise SystemExit(f"workload {name} has no multi-GPU generator")
    if gen == "uniform":
        ic = host.uniform_box(L, sid)
    elif gen == "sedov":
        ic = host.sedov_box(L, sid, bricks=bricks)
    elif gen == "clustered":
        ic = host.clustered_box(L, sid)
    elif gen == "active5":
        if world > 1:
            raise SystemExit(f"workload {name} has no multi-GPU generator")
        ic = host.jittered_box(L, sid, jitter=0.2, seed=42, active_fraction=0.05)
    else:
        ic = host.jittered_box(L, sid, jitter=0.2, seed=42, bricks=bricks)
    tg = host.default_top_grid(L)
    if os.environ.get("SWIFTGPU_TOPGRID"):  # experiments: other leaf sizes
        tg = (int(os.environ["SWIFTGPU_TOPGRID"]),) * 3
    cdim = tuple(t * b for t, b in zip(tg, bricks))
    dim = tuple(float(b) for b in bricks)
    if gen == "active5":
        # inactive particles carry the force-union members of "their last step": take them from an
        # all-active step of the same box (run here on the GPU), then apply the time bins
        from swift_b200.engine import SwiftGPU
        c_all = util.make_case(scheme, dict(ic, time_bin=np.ones_like(ic["time_bin"])), cdim)
        g0 = SwiftGPU(c_all.cfg)
        g0.upload_cells(c_all.tree.cells, c_all.tree.top)
        g0.upload_parts(c_all.parts)
        g0.set_step(c_all.step)
        g0.run_step(abi.PHASE_ALL)
        parts_all = g0.download_parts().copy()
        g0.close()
        c = util.make_case(scheme, ic, cdim, max_active_bin=1)
        c.parts = parts_all
        host.field(c.parts, c.layout, "time_bin")[:] = ic["time_bin"][c.tree.perm]
    else:
        c = util.make_case(scheme, ic, cdim, rank_grid=bricks, rank=rank, dim=dim, pack=(world == 1))
    c.sub_tree = c.tree
    if world > 1:
        sub, _, sel, is_local = host.extract_rank(c.tree, None, c.layout, rank)
        c.sub_tree = sub
        c.parts = host.pack_parts(c.layout, c.scheme, sub, ic)
        c.n_local = int(is_local.sum())
        c.n = int(sel.shape[0])
    else:
        c.n_local = c.n
    return c


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


def cpu_reference_run(workload, steps, warmup, sample_name=None):
    """Times the UNMODIFIED reference (oracle/_ref) - or the C port when the
    reference library did not travel - on a bounded sample of the workload with
    all host threads. Returns (interactions/s, dict)."""
    import util
    from oracle import port, ref
    from swift_b200 import abi
    scheme = WORKLOADS[workload][0]
    if sample_name is None:
        sample_name = {"sedov128": "sedov64", "sphenix128": "sphenix64", "sphenix256": "sphenix64",
                       "clustered128": "clustered64", "clustered256": "clustered64"}.get(workload, workload)
    if sample_name not in WORKLOADS:
        WORKLOADS[sample_name] = (scheme, 64, WORKLOADS[workload][2])
    c = make_workload(sample_name)
    cores = os.cpu_count() or 1
    # count the useful interactions once with the port (checker role, untimed)
    p = port.Port(scheme, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
    p.run(abi.PHASE_ALL)
    nd, ng, nf = p.counts()
    useful = int(nd.sum()) + int(ng.sum()) + int(nf.sum())
    p.close()
    kind = "reference" if ref.available(scheme) else "port"
    times = []
    if kind == "reference":
        o = ref.Reference(scheme, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
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
    info = {"value": useful / sec, "unit": "interactions/s", "cores": cores, "kind": kind,
            "sample": f"{sample_name}: {c.n} particles ({scheme}), same generator and density as {workload}; "
                      f"{useful} useful directed interactions per step, {sec * 1e3:.1f} ms/step wall-clock, "
                      f"mean of {len(times)} step(s), {cores} pthreads"}
    return info, sec, useful, c.n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    info, sec, useful, n = cpu_reference_run(args.workload, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": "SPH pair interactions/s (density+gradient+force)",
            "value": info["value"], "unit": "interactions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scheme": WORKLOADS[args.workload][0],
                       "note": "reference CPU path on a bounded sample of the workload"},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "interactions/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="swiftgpu")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "sedov128"
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    args.warmup = max(args.warmup, 3)
    scheme = WORKLOADS[args.workload][0]
    if world not in GRIDS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    c = make_workload(args.workload, world, rank)
    c.cfg.device = local_rank
    n = c.n
    psize = c.layout.size

    g = SwiftGPU(c.cfg)
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    g.upload_cells(c.sub_tree.cells, c.sub_tree.top)
    g.set_step(c.step)
    if world > 1:
        from swift_b200.engine import nccl_unique_id
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        g.halo_setup(idt.cpu().numpy().tobytes())

    host_in = torch.from_numpy(c.parts).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
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
        g.download_parts_device(dev_out.data_ptr())

    def step_e2e():
        g.upload_parts_ptr(host_in.data_ptr(), n)
        g.run_step(abi.PHASE_ALL)
        g.download_parts_ptr(host_out.data_ptr())

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
        useful = int(nd.sum()) + int(ng.sum()) + int(nf.sum())
        executed = int(st.n_density + st.n_gradient + st.n_force)

        # ---- end-to-end timing (pinned host buffers through the C ABI) ----
        step_e2e()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        f1.record(stream)
        barrier()
        sampler.stop()
        ms_e2e = f0.elapsed_time(f1)

    # max over ranks
    t = torch.tensor([ms_total, ms_e2e], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(useful), float(executed)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    useful_all, executed_all = float(tot[0]), float(tot[1])
    ms_step = ms_total / args.steps
    value = useful_all / (ms_step * 1e-3)
    e2e_value = useful_all / (ms_e2e / args.steps * 1e-3)

    # ---- roofline of the dominant kernel (CUDA-event phase times of the library) ----
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()
    props = torch.cuda.get_device_properties(local_rank)
    sms = props.multi_processor_count
    fp32_peak = sms * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # TFLOP/s at max SM clock
    ms_force = phase_ms["force"] / args.steps
    ms_density = phase_ms["density"] / args.steps
    fl_force = FLOPS["force_" + scheme]
    cand = {"density": st.t_density, "gradient": st.t_gradient, "force": st.t_force}
    # (CUDA-event ms of the phase = one launch of the loop kernel, algorithmic flops, algorithmic bytes)
    kernels = {
        "force": (ms_force, float(nf.sum()) * fl_force, c.n_local * BYTES["force"]),
        "density": (ms_density, float(nd.sum()) * FLOPS["density"], c.n_local * BYTES["density"]),
    }
    dom = max(kernels, key=lambda k: kernels[k][0])
    kms, kflops, kbytes = kernels[dom]
    achieved_tf = kflops / (kms * 1e-3) / 1e12
    achieved_gbs = kbytes / (kms * 1e-3) / 1e9
    kname = {"tile": "k_tile", "cta": "k_cta", "warp": "k_loop"}.get(os.environ.get("SWIFTGPU_LOOPS", "tile"), "k_tile")
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel, from the committed
    # ncu --set full capture of this command (profiles/r01c_summary.md); null for workloads not captured
    TRAFFIC = {("sedov128", "k_tile", "force"): 368.951552e6 + 60.983552e6,
               ("sedov128", "k_tile", "density"): 301.467904e6 + 62.868992e6}
    traffic = TRAFFIC.get((args.workload, kname, dom)) if world == 1 else None
    roofline = {"bound": "fp32", "kernel": (kname + "<FORCE,%s>" % scheme) if dom == "force" else (kname + "<DENSITY>"),
                "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
                "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01c_summary.md)",
                "algorithmic_bytes": kbytes,
                "peak_source": f"{sms} SMs x 128 FP32 lanes x 2 x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz, {peak_src})",
                "ms_per_launch": kms, "flops_per_interaction": fl_force if dom == "force" else FLOPS["density"],
                "candidates_per_hit": (cand[dom] / max(1.0, float(nf.sum() if dom == "force" else nd.sum())))}
    roofline_hbm = {"bound": "hbm", "kernel": roofline["kernel"], "achieved": achieved_gbs, "peak": hbm_peak,
                    "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "traffic": traffic,
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_src})"}

    line = {
        "metric": "SPH pair interactions/s (density+gradient+force)", "value": value, "unit": "interactions/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ("" if world == 1 else " x%d bricks %s (weak scaling: one brick per GPU)" % (world, "x".join(map(str, GRIDS[world])))),
                   "scheme": scheme, "particles": int(c.n_local) * world, "particles_per_gpu_incl_halo": int(n),
                   "active_fraction": 0.05 if WORKLOADS[args.workload][2] == "active5" else 1.0,
                   "l2": "inputs larger than L2 (%.0f MB AoS + SoA state per step)" % (n * psize / 1e6),
                   "top_grid": list(host.default_top_grid(WORKLOADS[args.workload][1])),
                   "ghost_iterations": int(st.ghost_iterations)},
        "e2e": {"value": e2e_value, "unit": "interactions/s", "h2d_bytes_per_step": n * psize,
                "d2h_bytes_per_step": n * psize, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "interactions_per_step": useful_all, "interactions_incl_ghost_reruns": executed_all,
        "phase_ms": {k: v / args.steps for k, v in phase_ms.items()},
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "clocks": sampler.summary(),
    }
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
