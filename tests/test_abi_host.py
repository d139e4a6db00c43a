"""CPU tests of the boundary and the host logic: the C-ABI library loads
without a GPU and exports every symbol include/swiftgpu.h declares; struct
mirrors have the C sizes; init fails LOUDLY (no CPU fallback) when there is no
device; the worklist flattener reproduces the reference's task decomposition
counts; the tree builder's invariants (space_split.c:50-330)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from swift_b200 import abi, host

ROOT = util.ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "swiftgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swiftgpu_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = abi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/swiftgpu.h but not exported"
    assert sorted(n for n, _, _ in abi.EXPORTS) == names


def test_struct_mirrors_match_c_sizes():
    assert C.sizeof(abi.PartLayout) == 33 * 4
    assert C.sizeof(abi.Cell) == abi.cell_dtype().itemsize == 160
    assert C.sizeof(abi.Step) == 32
    lib = abi.load()
    for s, name in enumerate(("minimal", "gadget2", "sphenix")):
        L = abi.PartLayout()
        assert lib.swiftgpu_default_layout(s, C.byref(L)) == 0
        assert L.as_dict() == util.golden_layout(name).as_dict()
        cfg = abi.Config()
        assert lib.swiftgpu_default_config(s, C.byref(cfg)) == 0
        assert cfg.max_smoothing_iterations == 30 and abs(cfg.h_tolerance - 1e-4) < 1e-9


def test_no_cpu_fallback():
    """Without a CUDA device swiftgpu_init must fail with a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = abi.load()
    cfg = abi.Config()
    lib.swiftgpu_default_config(0, C.byref(cfg))
    h = abi.VP()
    rc = lib.swiftgpu_init(C.byref(h), C.byref(cfg))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.swiftgpu_last_error(None)
    from swift_b200.engine import SwiftGPU
    with pytest.raises(RuntimeError):
        SwiftGPU(cfg)


def _stats(c, loop):
    lib = abi.load()
    out = np.zeros(6, np.int64)
    top = np.ascontiguousarray(c.tree.top, np.int32)
    cells = np.ascontiguousarray(c.tree.cells)
    rc = lib.swiftgpu_worklist_stats(C.byref(c.cfg), C.byref(c.step), cells.ctypes.data, len(cells),
                                     top.ctypes.data, len(top), loop, out.ctypes.data)
    assert rc == 0
    return out


def test_worklist_unsplit_grid():
    """3x3x3 unsplit periodic top cells: 27 selfs + 27*26/2 pairs (every couple
    of cells touches through the wrap), two directed items per pair."""
    ic = host.uniform_box(12, abi.SCHEME_MINIMAL)
    c = util.make_case("minimal", ic, (3, 3, 3))
    assert not c.tree.cells["split"].any()
    for loop in (0, 2):
        items, groups, cand, sorts, selfs, lim = _stats(c, loop)
        assert selfs == 27 and groups == 27
        assert items == 27 + 2 * (27 * 26 // 2)
        assert lim == 0
    items, groups, cand, sorts, selfs, lim = _stats(c, 3)
    assert selfs == 27 and items == 27 * 27


def test_worklist_split_recursion():
    """4096-particle top cells split twice (64 per leaf); DOSUB_PAIR1 recurses
    to the leaves (functions_hydro.h:2957: count < 100 stops): every leaf has
    one self + 26 neighbour leaves."""
    ic = host.uniform_box(48, abi.SCHEME_MINIMAL)
    c = util.make_case("minimal", ic, (3, 3, 3))
    cells = c.tree.cells
    leaves = cells[cells["split"] == 0]
    assert (leaves["count"] == 64).all() and cells["depth"].max() == 2
    items, groups, cand, sorts, selfs, lim = _stats(c, 0)
    assert groups == len(leaves) and selfs == len(leaves)
    assert items == len(leaves) * 27
    assert cand == len(leaves) * 27 * 64 * 64


def test_worklist_inactive_cells_produce_no_targets():
    ic = host.jittered_box(12, abi.SCHEME_MINIMAL, seed=1, active_fraction=0.1)
    c = util.make_case("minimal", ic, (3, 3, 3), max_active_bin=1)
    active_cells = (c.tree.cells["ti_end_min"] == c.step.ti_current).sum()
    items, groups, *_ = _stats(c, 0)
    assert groups <= active_cells


def test_tree_invariants():
    ic = host.clustered_box(20, abi.SCHEME_SPHENIX, seed=3, sigma=1.0)
    c = util.make_case("sphenix", ic, (3, 3, 3))
    cells = c.tree.cells
    x = host.field(c.parts, c.layout, "x")
    h = host.field(c.parts, c.layout, "h")
    assert sorted(c.tree.perm.tolist()) == list(range(c.n))
    for k, cell in enumerate(cells):
        f, n = int(cell["first_part"]), int(cell["count"])
        if n == 0:
            continue
        xs = x[f:f + n]
        assert (xs >= cell["loc"] - 1e-12).all() and (xs < cell["loc"] + cell["width"] + 1e-12).all()
        assert np.isclose(cell["h_max"], h[f:f + n].max())
        if cell["split"]:
            prog = [p for p in cell["progeny"] if p >= 0]
            assert sum(int(cells[p]["count"]) for p in prog) == n
            assert n > 400   # space_splitsize
            for p in prog:
                assert cells[p]["parent"] == k and cells[p]["depth"] == cell["depth"] + 1
    # depth_h: h_min_allowed <= h < h_max_allowed at the assigned level (cell.h:1787)
    dh = host.field(c.parts, c.layout, "depth_h")
    assert (dh >= 0).all() and (dh <= cells["depth"].max()).all()


_DIGEST_SCRIPT = """
import sys, ctypes as C, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import util
from swift_b200 import abi, host
ic = host.clustered_box(32, abi.SCHEME_SPHENIX, seed=5, sigma=2.0)
c = util.make_case("sphenix", ic, (4, 4, 4))
lib = abi.load()
cells = np.ascontiguousarray(c.tree.cells); top = np.ascontiguousarray(c.tree.top, np.int32)
out = []
for loop in (0, 2, 3):
    d = C.c_uint64(0)
    assert lib.swiftgpu_worklist_digest(C.byref(c.cfg), C.byref(c.step), cells.ctypes.data, len(cells),
                                        top.ctypes.data, len(top), loop, C.byref(d)) == 0
    out.append(d.value)
print("DIGEST", *out)
"""


def test_worklists_do_not_depend_on_the_host_thread_count():
    """The flattening of the reference's recursion runs on host threads over ranges of top-level
    cells (worklist.hpp). The lists - every item and group, in order, hence also the order of the
    device's sums - must be the same for 1, 3 and 8 threads (multi-level clustered tree, 64 top cells)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = _DIGEST_SCRIPT.format(root=root, tests=os.path.join(root, "tests"))
    seen = set()
    for nt in ("1", "3", "8"):
        e = dict(os.environ, SWIFTGPU_HOST_THREADS=nt)
        r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0]
        seen.add(line)
    assert len(seen) == 1, seen


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): ONE JSON line
    with the contract's keys, the reference's own CPU path timed on the host, nothing of the product."""
    import json
    import os
    import subprocess
    import sys
    from oracle import ref
    if not ref.available("minimal"):
        pytest.skip("needs oracle/_ref")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload",
                        "uniform32", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["workload"] == "uniform32"
