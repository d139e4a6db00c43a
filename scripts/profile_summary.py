"""Turns the ncu outputs of a gpurun call into the tracked text summaries under
profiles/:  python scripts/profile_summary.py <tag> <launches.csv> <full.ncu-rep> [workload]
With a workload name it also writes profiles/<tag>_traffic.json (DRAM bytes per launch of the
neighbour-loop kernels: what bench.py reports as roofline.traffic)."""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out = []
rows = list(csv.reader(open(launches)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
out.append(f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n")
out.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / tot:.3f} |")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
HH = rr[0]
want = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out.append(f"\n# {tag}: ncu --set full, per captured launch\n")
for r in rr[2:]:
    out.append(f"## {r[HH.index('Kernel Name')]}  (units: {', '.join(rr[1][HH.index(w)] for w in want[:1] if w in HH)})")
    for w in want:
        if w in HH:
            out.append(f"- {w} = {r[HH.index(w)]} {rr[1][HH.index(w)]}")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
if len(sys.argv) > 4:
    traffic = {}
    for r in rr[2:]:
        nm = r[HH.index("Kernel Name")]
        mm = re.match(r"void k_pipe<(\d+), (\d+), (\d+), (\d+), (\d+)(?:, \d+)?>", nm)
        if not mm: continue
        loop = {"0": "density", "1": "gradient", "2": "force"}[mm.group(1)]
        if loop == "density" and mm.group(5) == "256": loop = "subset"
        def val(name):
            v, u = float(r[HH.index(name)].replace(",", "")), rr[1][HH.index(name)]
            return 0.0 if v != v else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        # the launch that did the work (device-gated launches of the other task size move nothing)
        key = "k_pipe:" + loop
        traffic[key] = max(traffic.get(key, 0.0), val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    json.dump({sys.argv[4]: traffic, "source": f"ncu --set full --clock-control none, largest captured launch of each kernel ({os.path.basename(rep)})"},
              open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_summary.md")
