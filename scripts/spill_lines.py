"""Which source lines carry local-memory (spill) instructions of a kernel."""
import subprocess, sys, re, os, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ksub = sys.argv[1]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "swift_b200", "libswiftgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cur = None; line = None; cnt = collections.Counter(); n = 0
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m: cur = m.group(1); continue
    if cur is None or ksub not in cur: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        n += 1
        if re.search(r"\b(LDL|STL)", m.group(2)): cnt[(line, m.group(2).split()[0] if not m.group(2).startswith('@') else m.group(2).split()[1])] += 1
print("instructions", n)
for (ln, op), c in sorted(cnt.items(), key=lambda kv: (kv[0][0] or ("", 0))):
    print(ln, op, c)
