"""Prints the parity report (max relative errors against the oracle) of a full
step for the three schemes: how much of the 1e-5 bar is used."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from swift_b200 import abi, host
for scheme in ("minimal", "gadget2", "sphenix"):
    for L, seed in ((20, 7), (24, 3)):
        ic = host.jittered_box(L, abi.SCHEMES[scheme], jitter=0.2, h_scatter=0.05, seed=seed)
        c = util.make_case(scheme, ic, (3, 3, 3))
        g = util.run_gpu(c)
        got = g.download_parts()
        o, kind = util.run_oracle(c)
        rep = util.parity_report(got, o.parts(), c.layout, scheme, c.cfg.h_tolerance)
        print(scheme, L, kind, {k: (float("%.3g" % v) if isinstance(v, float) else v) for k, v in rep.items()})
        g.close()
