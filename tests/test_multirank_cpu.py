"""CPU tests of the multi-GPU host logic (no GPU, no NCCL): the rank
extraction (local cells + foreign neighbours, src/engine_proxy.c), the halo
plan of libswiftgpu (swiftgpu_halo_plan) and - with two real processes over
gloo - that the send list of one rank IS the receive list of the other, cell
by cell and particle by particle, which is what lets the NCCL exchange run
without negotiation."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from swift_b200 import abi, host

ROOT = util.ROOT


def plan(cfg, cells, top, peer):
    lib = abi.load()
    cells = np.ascontiguousarray(cells)
    top = np.ascontiguousarray(top, np.int32)
    ntop = len(top)
    sc = np.zeros(ntop, np.int32); rc = np.zeros(ntop, np.int32)
    ns = C.c_int32(); nr = C.c_int32(); ps = C.c_int64(); pr = C.c_int64()
    ret = lib.swiftgpu_halo_plan(C.byref(cfg), cells.ctypes.data, len(cells), top.ctypes.data, ntop, peer,
                                 sc.ctypes.data, C.addressof(ns), rc.ctypes.data, C.addressof(nr),
                                 C.addressof(ps), C.addressof(pr))
    assert ret == 0
    return sc[:ns.value], rc[:nr.value], ps.value, pr.value


def make_rank_case(scheme, L, cdim, rank_grid, rank, seed=5):
    ic = host.jittered_box(L, abi.SCHEMES[scheme], jitter=0.2, h_scatter=0.05, seed=seed)
    c = util.make_case(scheme, ic, cdim, rank_grid=rank_grid, rank=rank)
    sub, parts, sel, is_local = host.extract_rank(c.tree, c.parts, c.layout, rank)
    return c, sub, parts, sel, is_local


def test_extract_rank_consistency():
    c, sub, parts, sel, is_local = make_rank_case("minimal", 16, (4, 4, 4), (2, 1, 1), 0)
    cells = sub.cells
    # every kept cell's particles are the same bytes as in the global array
    size = c.layout.size
    g = c.parts.reshape(-1, size)
    s = parts.reshape(-1, size)
    assert np.array_equal(s, g[sel])
    # local particles = those of nodeID == rank cells; with a 2x1x1 grid of 4^3
    # top cells and periodic wrap every foreign cell touches a local one
    assert is_local.sum() == (c.tree.cells["count"][c.tree.top][c.tree.cells["nodeID"][c.tree.top] == 0]).sum()
    assert len(sub.top) == 64
    # tree invariants after re-indexing
    for k, cell in enumerate(cells):
        if cell["split"]:
            prog = [p for p in cell["progeny"] if p >= 0]
            assert sum(int(cells[p]["count"]) for p in prog) == cell["count"]
            assert all(cells[p]["parent"] == k for p in prog)
        assert cells[cell["top"]]["depth"] == 0


def test_halo_plan_symmetric_single_process():
    """Rank 0's send list to rank 1 equals rank 1's receive list from rank 0."""
    plans = {}
    for rank in (0, 1):
        c, sub, parts, sel, is_local = make_rank_case("minimal", 16, (4, 4, 4), (2, 1, 1), rank)
        sc, rc, ps, pr = plan(c.cfg, sub.cells, sub.top, 1 - rank)
        x = host.field(parts, c.layout, "x")
        def pts(cell_list):
            return np.concatenate([x[int(sub.cells["first_part"][k]):int(sub.cells["first_part"][k]) + int(sub.cells["count"][k])]
                                   for k in cell_list])
        plans[rank] = (pts(sc), pts(rc), ps, pr)
    assert plans[0][2] == plans[1][3] and plans[0][3] == plans[1][2]
    assert np.array_equal(plans[0][0], plans[1][1])
    assert np.array_equal(plans[1][0], plans[0][1])
    assert plans[0][2] > 0


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SWIFT_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SWIFT_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
import test_multirank_cpu as T
from swift_b200 import host
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
c, sub, parts, sel, is_local = T.make_rank_case("sphenix", 16, (4, 4, 4), (2, 1, 1), rank)
peer = 1 - rank
sc, rc, ps, pr = T.plan(c.cfg, sub.cells, sub.top, peer)
size = c.layout.size
P = parts.reshape(-1, size)
def idx(cell_list):
    return np.concatenate([np.arange(int(sub.cells["first_part"][k]), int(sub.cells["first_part"][k]) + int(sub.cells["count"][k])) for k in cell_list])
si, ri = idx(sc), idx(rc)
# scramble the foreign copies, then refresh them from their owner over gloo
before = P[ri].copy()
P[ri] = 0
send = torch.from_numpy(np.ascontiguousarray(P[si]))
recv = torch.empty((len(ri), size), dtype=torch.uint8)
ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
for r in dist.batch_isend_irecv(ops): r.wait()
P[ri] = recv.numpy()
assert np.array_equal(P[ri], before), "halo refresh does not reproduce the foreign copies"
assert not is_local[ri].any() and is_local[si].all()
dist.barrier()
print("RANK_OK", rank, len(si), len(ri))
'''


def test_halo_plan_two_processes_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SWIFT_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"RANK_OK {r}" in o, o


def test_brick_generator_is_the_same_box_on_every_grid():
    """bench.py's strong-scaling workload: every rank generates only its brick + one top-level cell of
    halo (host.brick_box), yet the union over the ranks is the single-GPU box bit for bit, whatever the
    grid; the rank's own particles come first in its array (what swiftgpu_upload_parts_local needs)
    and every halo cell holds the same particles as its owner's copy (same count - the exchange ships
    cells in the sender's device order, so only the SET has to agree)."""
    L, top = 64, 4
    scheme = abi.SCHEME_SPHENIX
    whole = host.brick_box(L, scheme, top=top)
    by_id = np.argsort(whole["id"])
    for grid in ((2, 1, 1), (2, 2, 1), (2, 2, 2)):
        world = grid[0] * grid[1] * grid[2]
        seen = np.zeros(L ** 3 + 1, np.int32)
        per_rank = []
        for rank in range(world):
            ic = host.brick_box(L, scheme, grid=grid, rank=rank, top=top)
            pos = by_id[np.searchsorted(whole["id"][by_id], ic["id"])]
            assert np.array_equal(whole["x"][pos], ic["x"]) and np.array_equal(whole["u"][pos], ic["u"])
            c = util.make_case("sphenix", ic, (top,) * 3, rank_grid=grid, rank=rank, pack=False)
            sub, _, sel, is_local = host.extract_rank(c.tree, None, c.layout, rank, local_first=True)
            nl = int(is_local.sum())
            assert is_local[:nl].all() and not is_local[nl:].any()
            ids = np.asarray(ic["id"])[sub.perm]
            seen[ids[:nl]] += 1
            per_rank.append((sub, ids))
            # local top-level cells own exactly [0, nl)
            tops = sub.cells[sub.top]
            loc = tops["nodeID"] == rank
            assert (tops["first_part"][loc] + tops["count"][loc]).max() == nl
            assert tops["first_part"][~loc].min() >= nl
        assert (seen[1:] == 1).all(), "every particle is local on exactly one rank"
        # a proxy cell holds the same particle set as the owner's cell at the same location
        def cell_sets(sub, ids):
            out = {}
            for t in sub.top:
                cc = sub.cells[t]
                f, n = int(cc["first_part"]), int(cc["count"])
                out[tuple(np.round(cc["loc"] * top).astype(int))] = (int(cc["nodeID"]), frozenset(ids[f:f + n].tolist()))
            return out
        sets = [cell_sets(s, i) for s, i in per_rank]
        for r, cs in enumerate(sets):
            for loc, (owner, members) in cs.items():
                if owner != r:
                    assert sets[owner][loc][1] == members
