"""Summarise an ncu report's source page: instructions executed per CUDA source
line (needs -lineinfo and --import-source on).
   python scripts/ncu_hot.py report.ncu-rep kernel_regex [top] [launch_index]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:k_"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; files = []; cur = None; fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        cur = {"fn": r[1], "file": fname, "lines": []}; files.append(cur); continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or cur is None or not r or not r[0].isdigit(): continue
    i_inst = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed"); i_smp = hdr.index("# Samples")
    try:
        cur["lines"].append((int(r[0]), r[1], int(r[i_inst]), int(r[i_thr]), int(r[i_smp])))
    except ValueError:
        pass
# group per kernel function occurrence: the report lists each (file, function) once per launch
import re
fns = sorted(set(f["fn"] for f in files if re.search(kern, f["fn"])))
sel_fn = fns[0]
launches = {}
for f in files:
    if f["fn"] != sel_fn: continue
    launches.setdefault(f["file"], []).append(f)
lines = []
for fn_file, lst in launches.items():
    if which < len(lst):
        lines += [(fn_file,) + l for l in lst[which]["lines"]]
tot = sum(l[3] for l in lines); tots = sum(l[5] for l in lines)
print(f"{sel_fn}: total warp instr {tot:.3e} samples {tots}")
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{l[3]/tot*100:5.1f}% inst {l[5]/max(tots,1)*100:5.1f}% smp thr/inst {l[4]/max(l[3],1):5.1f} | {l[0]}:{l[1]} {l[2].strip()[:100]}")
