"""Per-kernel SASS mnemonic counts of swift_b200/libswiftgpu.so (cuobjdump -sass): the evidence that
the loops use bulk TMA (UBLKCP) with mbarrier completion (SYNCS), and where FP64 / MUFU / conversion
instructions remain.  python scripts/sass_counts.py > profiles/<tag>_sass.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "swift_b200", "libswiftgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = ["UBLKCP", "SYNCS", "DADD", "DMUL", "DFMA", "F2F", "MUFU", "FFMA", "FMUL", "FADD", "LDS", "STS", "LDG", "STG",
        "ATOMG", "RED", "SHFL", "VOTE", "BAR", "BRA"]
kern, counts, total, arch = None, {}, {}, {}
cur_arch = ""
for line in txt.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m: cur_arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); total[kern] = 0; arch[kern] = cur_arch
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1); total[kern] += 1
        for w in WANT:
            if op == w or op.startswith(w + ".") or (w == "UBLKCP" and op.startswith("UBLKCP")):
                counts[kern][w] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(so, ROOT)} (cuobjdump -sass; scripts/sass_counts.py)")
print("# UBLKCP = cp.async.bulk (TMA bulk copy); SYNCS = mbarrier arrive/expect_tx/try_wait; DADD/F2F = the reference's double frame subtraction + float cast")
print("| kernel | arch | instr | " + " | ".join(WANT) + " |")
print("|---|---|---|" + "---|" * len(WANT))
for k, d in zip(counts, demangle):
    d = (d.split(">(")[0] + ">") if ">(" in d else re.sub(r"\(.*", "", d)
    d = d.replace("(int)", "").replace("void ", "").replace("swiftgpu::", "").replace("(anonymous namespace)::", "")
    print(f"| `{d}` | {arch[k]} | {total[k]} | " + " | ".join(str(counts[k][w]) for w in WANT) + " |")
