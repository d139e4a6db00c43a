"""Compact ptxas -v summary: registers / spills of the kernels matching a substring."""
import re, sys
sub = sys.argv[2] if len(sys.argv) > 2 else ""
txt = open(sys.argv[1]).read()
for m in re.finditer(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
    if sub in m.group(1):
        print("%-60s regs %s stack %s spill st/ld %s/%s" % (m.group(1)[:60], m.group(5), m.group(2), m.group(3), m.group(4)))
