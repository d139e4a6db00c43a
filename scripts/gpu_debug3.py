"""Debug: repeat the strongly clustered SPHENIX step, count particles whose force outputs move between runs."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from swift_b200 import abi, host
ic = host.clustered_box(32, abi.SCHEME_SPHENIX, seed=2025, sigma=2.5)
ic["h"] = (ic["h"] * np.float32(0.6)).astype(np.float32)
c = util.make_case("sphenix", ic, (3, 3, 3))
p = util.run_port(c)
pnd, png, pnf = p.counts()
ref = p.parts()
g = util.run_gpu(c)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    g.upload_parts(c.parts); g.run_step(abi.PHASE_ALL)
    nd, ng, nf = g.download_counts()
    got = g.download_parts()
    out = []
    for name in ("u_dt", "h_dt", "laplace_u", "rho"):
        a = host.field(got, c.layout, name).astype(np.float64); b = host.field(ref, c.layout, name).astype(np.float64)
        e = np.abs(a - b) / np.maximum(np.abs(b), 1e-2 * np.abs(b).mean())
        out.append("%s bad %d max %.2g" % (name, (e > 1e-2).sum(), e.max()))
    print("rep", rep, "nd", (nd != pnd).sum(), "ng", (ng != png).sum(), "nf", (nf != pnf).sum(), "|", " | ".join(out))
