"""Hot source lines of a kernel: joins the SASS-level samples of an .ncu-rep
(--page source --csv) with nvdisasm -g line info of the in-tree cubin.
usage: ncu_lines.py rep.ncu-rep <kernel-substr> [launch-index] [top]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, ksub = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "swift_b200", "libswiftgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split ncu output into kernels
blocks = []
cur = None
for line in src.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]; blocks.append(cur)
    elif cur is not None:
        cur.append(line)
sel = [b for b in blocks if ksub in b[0]]
b = sel[which]
name = list(csv.reader([b[0]]))[0][1]
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hd = rows[0]
ia, isamp, iinst, ithr = hd.index("Address"), hd.index("# Samples"), hd.index("Instructions Executed"), hd.index("Thread Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(hd) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[1][ia], 16)
# mangled name: find function section in disasm whose instruction count matches
# map offsets -> line for all functions, pick by sass text match of first instrs
funcs = {}
curf = None; curline = None
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: curf = m.group(1); funcs[curf] = {}; curline = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: curline = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and curf: funcs[curf][int(m.group(1), 16)] = (curline, m.group(2).strip())
n = len(rows) - 1
def norm(s): return re.sub(r"\s+", " ", s.strip().rstrip(";").strip())
best = None
kbase = re.sub(r'^void\s+', '', name).split('<')[0].split('::')[-1]
for f, mp in sorted(funcs.items(), key=lambda kv: (kbase not in kv[0])):
    if len(mp) != n: continue
    ok = all(norm(mp.get(16 * k, (None, ""))[1]) == norm(rows[1 + k][1]) for k in range(0, min(n, 40)))
    if ok: best = f; break
if best is None:
    cands = [f for f, mp in funcs.items() if len(mp) == n]
    best = cands[0] if cands else None
print("kernel:", name[:80], "->", best, "instrs", n)
mp = funcs[best]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = 0; toti = 0
for r in rows[1:]:
    off = int(r[ia], 16) - base
    ln = mp.get(off, (None, ""))[0]
    s = int(r[isamp] or 0); tot += s
    a = agg[ln]; a[0] += s; a[1] += int(r[iinst] or 0); a[2] += int(r[ithr] or 0); toti += int(r[iinst] or 0)
    for i, c in stall_cols:
        v = int(r[i] or 0)
        if v: a[3][c[6:]] += v
srcs = {}
print("total samples", tot, "warp-instr", toti)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if ln:
        p = os.path.join(root, "swift_b200", "csrc", ln[0])
        if p not in srcs and os.path.exists(p): srcs[p] = open(p).read().splitlines()
        if p in srcs and ln[1] - 1 < len(srcs[p]): text = srcs[p][ln[1] - 1].strip()[:70]
    st = ",".join("%s:%d" % (k, v) for k, v in a[3].most_common(3))
    print("%5.1f%% smp %4.1f%% ins  %-22s %-70s %s" % (100 * a[0] / max(tot, 1), 100 * a[1] / max(toti, 1), "%s:%d" % ln if ln else "?", text, st))
# region breakdown (instruction share by source file / line range), optional: REGIONS="file:lo-hi=name,..."
reg = os.environ.get("REGIONS")
if reg:
    regs = []
    for it in reg.split(","):
        k, nm = it.split("=")
        f, rng = k.split(":")
        lo, hi = rng.split("-")
        regs.append((f, int(lo), int(hi), nm))
    out = collections.Counter(); outs = collections.Counter()
    for ln, a in agg.items():
        nm = "other"
        if ln:
            for f, lo, hi, n_ in regs:
                if ln[0] == f and lo <= ln[1] <= hi: nm = n_; break
            else:
                nm = ln[0]
        out[nm] += a[1]; outs[nm] += a[0]
    for k, v in out.most_common():
        print("region %-28s ins %5.1f%%  samples %5.1f%%" % (k, 100 * v / toti, 100 * outs[k] / tot))
