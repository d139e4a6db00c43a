"""Summarise an .ncu-rep: per kernel duration, issue utilisation and the top
warp-stall reasons (raw page), plus the hottest source lines (source page)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("##", r[idx["Kernel Name"]][:60])
    for k in keys:
        if k in idx:
            print("   %-70s %s" % (k, r[idx[k]]))
    s = sorted([(float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stall], reverse=True)[:7]
    print("   stalls/issue:", [(round(a, 2), b) for a, b in s])
if top:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    # split per kernel blocks
    blocks = src.split("\n\n")
    for blk in blocks:
        rr = list(csv.reader(io.StringIO(blk)))
        if len(rr) < 3: continue
        h = None
        for i, r in enumerate(rr):
            if "Source" in r and any("Samples" in c for c in r):
                h = i; break
        if h is None: continue
        hd = rr[h]
        si = hd.index("Source")
        cand = [i for i, c in enumerate(hd) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
        if not cand: continue
        ci = cand[0]
        ei = [i for i, c in enumerate(hd) if c == "Instructions Executed"]
        tot = 0; lines = []
        for r in rr[h + 1:]:
            if len(r) <= ci: continue
            try: v = float(r[ci])
            except: continue
            tot += v
            ie = r[ei[0]] if ei else ""
            lines.append((v, r[0] if r[0] != r[si] else "", r[si].strip()[:110], ie))
        lines.sort(reverse=True)
        print("== source hot lines (samples, share) ==", rr[0][:1])
        for v, ln, s, ie in lines[:top]:
            print("  %6.0f %5.1f%%  inst=%s  %s" % (v, 100 * v / max(tot, 1), ie, s))
