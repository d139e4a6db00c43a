"""Debug: force-count mismatches of the gradient-list test case, by depth / cell size."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
from swift_b200 import abi, host
ic = host.clustered_box(32, abi.SCHEME_SPHENIX, seed=2025, sigma=2.5)
ic["h"] = (ic["h"] * np.float32(0.6)).astype(np.float32)
c = util.make_case("sphenix", ic, (3, 3, 3))
g = util.run_gpu(c)
st = g.stats()
print("rebuilds force", st.force_list_rebuilds, "gradient", st.gradient_list_rebuilds, "ghost it", st.ghost_iterations)
nd, ng, nf = g.download_counts()
p = util.run_port(c)
pnd, png, pnf = p.counts()
got = g.download_parts()
dh = host.field(got, c.layout, "depth_h")
print("nd diff", (nd != pnd).sum(), "ng diff", (ng != png).sum(), "nf diff", (nf != pnf).sum(), "of", nd.size)
bad = np.nonzero(nf != pnf)[0]
cells = c.tree.cells
leaf_of = np.zeros(nd.size, np.int64)
for ci in np.nonzero(cells["split"] == 0)[0]:
    f, n = int(cells["first_part"][ci]), int(cells["count"][ci])
    leaf_of[f:f + n] = ci
for k in bad[:40]:
    l = leaf_of[k]
    print(k, "nf", nf[k], "port", pnf[k], "depth_h", dh[k], "leaf depth", cells["depth"][l], "leaf count", cells["count"][l], "h", host.field(got, c.layout, "h")[k])
print("bad by depth_h:", np.unique(dh[bad], return_counts=True), "all:", np.unique(dh, return_counts=True))
print("diff sign:", np.unique(np.sign(nf[bad].astype(int) - pnf[bad]), return_counts=True))
bad = np.nonzero(ng != png)[0]
for k in bad[:20]:
    l = leaf_of[k]
    print("ng", k, ng[k], "port", png[k], "depth_h", dh[k], "leaf depth", cells["depth"][l], "leaf count", cells["count"][l])
ref = p.parts()
for name in ("h_dt", "u_dt"):
    a = host.field(got, c.layout, name).astype(np.float64); b = host.field(ref, c.layout, name).astype(np.float64)
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * np.abs(b).mean())
    w = np.argsort(e)[-12:]
    print(name, "worst:")
    for k in w:
        l = leaf_of[k]
        print("  ", k, a[k], b[k], "depth_h", dh[k], "leaf depth", cells["depth"][l], "count", cells["count"][l], "nf", nf[k])
    print(name, "n bad(>1e-3):", (e > 1e-3).sum(), "by depth_h", np.unique(dh[e > 1e-3], return_counts=True))
