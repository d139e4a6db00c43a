"""Development probe: GPU path vs the oracles, phase by phase, verbose."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from swift_b200 import abi, host
from swift_b200.engine import SwiftGPU
from oracle import ref, port
import util

def run(variant, L, kind, cd=None):
    sname = variant.split('_')[0]; scheme = abi.SCHEMES[sname]
    lay = ref.layout(variant) if ref.available(variant) else util.golden_layout(variant)
    if kind == "uni": ic = host.uniform_box(L, scheme)
    elif kind == "jit": ic = host.jittered_box(L, scheme, jitter=0.2, h_scatter=0.1)
    elif kind == "clu": ic = host.clustered_box(L, scheme)
    cd = cd or host.default_top_grid(L)
    c = util.make_case(sname, ic, cd, layout=lay)
    print(f"=== {variant} L={L} {kind} cdim={cd} cells={c.tree.cells.shape[0]} maxdepth={c.tree.cells['depth'].max()}")
    phases = [("sort+density", abi.PHASE_SORT | abi.PHASE_DENSITY), ("ghost", abi.PHASE_GHOST),
              ("gradient+extra", abi.PHASE_GRADIENT | abi.PHASE_EXTRA_GHOST), ("force", abi.PHASE_FORCE),
              ("end_force", abi.PHASE_END_FORCE)]
    r = ref.Reference(variant, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
    p = port.Port(sname, c.cfg, c.step, c.tree.cells, c.tree.top, c.parts)
    g = SwiftGPU(c.cfg)
    g.upload_cells(c.tree.cells, c.tree.top); g.upload_parts(c.parts); g.set_step(c.step)
    for name, ph in phases:
        t = time.time(); r.run(ph, threads=8); tr = time.time() - t
        p.run(ph)
        t = time.time(); g.run_step(ph); tg = time.time() - t
        ro, po, go = r.parts(), p.parts(), g.download_parts()
        cp, cg = p.counts(), g.download_counts()
        st = g.stats()
        names = ["h", "rho"]
        if ph & abi.PHASE_DENSITY: names += ["wcount", "wcount_dh", "rho_dh", "div_v", "rot_v"]
        if ph >= abi.PHASE_GHOST: names += ["f", "soundspeed", "balsara"] + (["pressure"] if scheme != 1 else ["P_over_rho2"])
        if ph >= abi.PHASE_FORCE: names += ["a_hydro", "h_dt", "v_sig"] + (["u_dt"] if scheme != 1 else ["entropy_dt"])
        if scheme == 2 and ph >= abi.PHASE_GRADIENT: names += ["v_sig", "laplace_u", "visc_alpha", "diff_alpha", "alpha_visc_max_ngb", "div_v"]
        e_ref = util.compare_fields(go, ro, lay, names, 1e-5)
        e_port = util.compare_fields(go, po, lay, names, 1e-5)
        print(f"[{name}] ref {tr:.3f}s gpu(wall) {tg:.3f}s  counts vs port: nd={np.array_equal(cp[0], cg[0])} ng={np.array_equal(cp[1], cg[1])} nf={np.array_equal(cp[2], cg[2])}"
              f"  nd_sum gpu={cg[0].sum()} port={cp[0].sum()}  nf_sum gpu={cg[2].sum()} port={cp[2].sum()}")
        print("   vs ref : ", {k: f"{v:.1e}" for k, v in e_ref.items()})
        print("   vs port: ", {k: f"{v:.1e}" for k, v in e_port.items()})
        if not np.array_equal(cp[0], cg[0]):
            bad = np.nonzero(cp[0] != cg[0])[0]; print("   nd bad:", bad[:10], cp[0][bad[:10]], cg[0][bad[:10]], "nbad", bad.size)
        if ph >= abi.PHASE_FORCE and not np.array_equal(cp[2], cg[2]):
            bad = np.nonzero(cp[2] != cg[2])[0]; print("   nf bad:", bad[:10], cp[2][bad[:10]], cg[2][bad[:10]], "nbad", bad.size)
        print(f"   depth_h eq ref: {np.array_equal(host.field(go, lay, 'depth_h'), host.field(ro, lay, 'depth_h'))}"
              f"  stats: ms sort {st.ms_sort:.3f} dens {st.ms_density:.3f} ghost {st.ms_ghost:.3f} grad {st.ms_gradient:.3f} force {st.ms_force:.3f} iters {st.ghost_iterations}")
    hr, hg = host.field(ro, lay, "h"), host.field(go, lay, "h")
    print("   h flips (>1e-5):", int((np.abs(hr - hg) / hr > 1e-5).sum()), "of", hr.size)
    cr, cgc = r.cells(), g.download_cells()
    print("   cell h_max eq:", np.array_equal(cr["h_max"], cgc["h_max"]), np.array_equal(cr["h_max_active"], cgc["h_max_active"]))
    g.close()

if __name__ == "__main__":
    args = sys.argv[1:]
    run(args[0], int(args[1]), args[2])
