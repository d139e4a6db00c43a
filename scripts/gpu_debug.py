"""Verbose GPU-vs-oracle run used while bringing the kernels up (gpurun)."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util
from swift_b200 import abi, host

def one(scheme, L=16, cdim=(3,3,3), **kw):
    ic = host.jittered_box(L, abi.SCHEMES[scheme], **kw)
    c = util.make_case(scheme, ic, cdim)
    masks = [("density", abi.PHASE_SORT | abi.PHASE_DENSITY), ("all", abi.PHASE_ALL)]
    for name, mask in masks:
        try:
            t = time.time()
            g = util.run_gpu(c, mask)
            dt = time.time() - t
            st = g.stats()
            p = util.run_port(c, mask)
            nd, ng, nf = g.download_counts(); pnd, png, pnf = p.counts()
            got = g.download_parts()
            print(f"[{scheme} L={L} {name}] gpu {dt:.3f}s ms: sort {st.ms_sort:.3f} dens {st.ms_density:.3f} ghost {st.ms_ghost:.3f} grad {st.ms_gradient:.3f} force {st.ms_force:.3f} "
                  f"n_d {st.n_density} n_g {st.n_gradient} n_f {st.n_force} iters {st.ghost_iterations}/{p.ghost_iterations()} launches {st.n_launches}")
            print("   count mismatches: d", int((nd != pnd).sum()), "g", int((ng != png).sum()), "f", int((nf != pnf).sum()),
                  " sums", nd.sum(), pnd.sum(), nf.sum(), pnf.sum())
            if mask == abi.PHASE_ALL:
                o, kind = util.run_oracle(c, mask)
                print("   vs", kind, util.parity_report(got, o.parts(), c.layout, scheme))
            else:
                for f in ("rho", "wcount", "rho_dh", "div_v"):
                    a = host.field(got, c.layout, f).astype(np.float64); b = host.field(p.parts(), c.layout, f).astype(np.float64)
                    print("   ", f, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * np.abs(b).max()))))
            g.close()
        except Exception:
            traceback.print_exc()

if __name__ == "__main__":
    for s in ("minimal", "gadget2", "sphenix"):
        one(s, 16, jitter=0.2, h_scatter=0.05, seed=7)
    one("sphenix", 32, (4,4,4), jitter=0.2, h_scatter=0.2, seed=9)
