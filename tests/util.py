"""Shared helpers of the test-suite: build a synthetic case (config, step
scalars, cell tree, AoS particles) and compare particle fields."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from swift_b200 import abi, host  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_layout(variant):
    d = json.load(open(os.path.join(GOLDEN, "part_layouts.json")))[variant]
    return abi.PartLayout.from_dict(d)


def make_config(scheme, layout, rank=0, nranks=1, h_tolerance=1e-4, max_iter=30, h_max=1e10):
    cfg = abi.Config()
    cfg.abi_version = 1
    cfg.scheme = scheme
    cfg.device = 0
    cfg.periodic = 1
    cfg.dim[:] = [1.0, 1.0, 1.0]
    cfg.eta_neighbours = 1.2348
    cfg.h_tolerance = h_tolerance
    cfg.h_max = h_max
    cfg.h_min = 0.0
    cfg.max_smoothing_iterations = max_iter
    cfg.use_mass_weighted_num_ngb = 0
    cfg.CFL_condition = 0.1
    cfg.viscosity_alpha = 0.1 if scheme == abi.SCHEME_SPHENIX else 0.8
    cfg.viscosity_alpha_max, cfg.viscosity_alpha_min, cfg.viscosity_length = 2.0, 0.0, 0.05
    cfg.diffusion_alpha, cfg.diffusion_beta = 0.0, 1.0
    cfg.diffusion_alpha_max, cfg.diffusion_alpha_min = 1.0, 0.0
    cfg.rank, cfg.nranks = rank, nranks
    cfg.layout = layout
    return cfg


def make_step(max_active_bin=56):
    st = host.step_scalars(max_active_bin)
    s = abi.Step()
    s.ti_current, s.max_active_bin, s.time_base = st["ti_current"], st["max_active_bin"], st["time_base"]
    s.with_cosmology, s.a, s.H = 0, 1.0, 0.0
    return s


class Case:
    pass


def make_case(scheme_name, ic, cdim, layout=None, max_active_bin=56, rank_grid=(1, 1, 1), rank=0, **cfg_kw):
    c = Case()
    c.scheme_name = scheme_name
    c.scheme = abi.SCHEMES[scheme_name]
    c.layout = layout if layout is not None else golden_layout(scheme_name)
    nranks = rank_grid[0] * rank_grid[1] * rank_grid[2]
    c.cfg = make_config(c.scheme, c.layout, rank=rank, nranks=nranks, **cfg_kw)
    c.step = make_step(max_active_bin)
    c.tree = host.build_tree(ic["x"], ic["h"], ic["time_bin"], (1.0, 1.0, 1.0), cdim,
                             c.step.max_active_bin, c.step.ti_current, rank_grid=rank_grid)
    c.parts = host.pack_parts(c.layout, c.scheme, c.tree, ic)
    c.n = ic["x"].shape[0]
    c.ic = ic
    return c


def rel_err(a, b, floor):
    """|a-b| / max(|b|, floor): the `ignore-below` logic of the reference's
    tests/difffloat.py (values below `floor` are compared absolutely)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def compare_fields(got_u8, ref_u8, layout, names, rtol, report=None, floors=None):
    """Returns dict name -> max relative error (floor = 1e-3 * rms(ref) unless given)."""
    out = {}
    for n in names:
        g = host.field(got_u8, layout, n).astype(np.float64)
        r = host.field(ref_u8, layout, n).astype(np.float64)
        fl = (floors or {}).get(n)
        if fl is None:
            fl = max(1e-3 * float(np.sqrt(np.mean(r * r))), 1e-30)
        e = rel_err(g, r, fl)
        out[n] = float(e.max())
    return out
