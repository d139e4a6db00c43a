"""Host-side harness around the C ABI: synthetic initial conditions, the cell
tree (libswiftgpu_host.so, mirrors space_regrid/space_split) and AoS packing.

In a real deployment SWIFT owns all of this (struct space); the harness exists
for the tests and the benchmark only.
"""
import ctypes as C
import math
import os

import numpy as np

from . import abi

KERNEL_GAMMA = np.float32(1.825742)          # kernel_hydro.h:52
HYDRO_GAMMA = 5.0 / 3.0                       # adiabatic_index.h:43


_FIELD_TYPES = {"id": ("<i8", ()), "x": ("<f8", (3,)), "v": ("<f4", (3,)),
                "a_hydro": ("<f4", (3,)), "rot_v": ("<f4", (3,)),
                "time_bin": ("i1", ()), "depth_h": ("i1", ()),
                "min_ngb_time_bin": ("i1", ())}


def field(parts_u8, layout, name):
    """Strided numpy view of one field of the AoS `struct part` byte array."""
    L = layout.as_dict() if hasattr(layout, "as_dict") else dict(layout)
    off, size = L[name], L["size"]
    if off < 0:
        raise KeyError(f"field {name} not present in this scheme's struct part")
    fmt, shape = _FIELD_TYPES.get(name, ("<f4", ()))
    dt = np.dtype(fmt)
    n = parts_u8.size // size
    strides = (size,) + tuple(dt.itemsize for _ in shape)
    return np.ndarray(shape=(n,) + shape, dtype=dt, buffer=parts_u8, offset=off, strides=strides)


def has_field(layout, name):
    L = layout.as_dict() if hasattr(layout, "as_dict") else dict(layout)
    return L.get(name, -1) >= 0


class Tree:
    def __init__(self, cells, top, perm, depth_h):
        self.cells, self.top, self.perm, self.depth_h = cells, top, perm, depth_h


def build_tree(x, h, time_bin, dim, cdim, max_active_bin, ti_current,
               splitsize=400, rank_grid=(1, 1, 1)):
    """space_regrid + space_split equivalent. Returns Tree; tree.perm[new]=old."""
    host = abi.load_host()
    n = x.shape[0]
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = np.ascontiguousarray(h, dtype=np.float32)
    tb = np.ascontiguousarray(time_bin, dtype=np.int8)
    perm = np.empty(n, dtype=np.int64)
    depth_h = np.empty(n, dtype=np.int8)
    dim_a = (C.c_double * 3)(*dim)
    cdim_a = (C.c_int * 3)(*cdim)
    rg = (C.c_int * 3)(*rank_grid)
    t = host.swifthost_build_tree(x.ctypes.data, h.ctypes.data, tb.ctypes.data, n,
                                  C.addressof(dim_a), C.addressof(cdim_a), splitsize,
                                  max_active_bin, ti_current, C.addressof(rg),
                                  perm.ctypes.data, depth_h.ctypes.data)
    nc, nt = host.swifthost_tree_ncells(t), host.swifthost_tree_ntop(t)
    cells = np.zeros(nc, dtype=abi.cell_dtype())
    top = np.zeros(nt, dtype=np.int32)
    host.swifthost_tree_copy(t, cells.ctypes.data, top.ctypes.data)
    host.swifthost_tree_free(t)
    assert cells.dtype.itemsize == C.sizeof(abi.Cell)
    return Tree(cells, top, perm, depth_h)


def pack_parts(layout, scheme, tree, ic):
    """AoS struct part[] in cell order (what space->parts holds)."""
    host = abi.load_host()
    n = int(tree.perm.shape[0])  # a rank's sub-tree packs only its own particles
    out = np.zeros(n * layout.size, dtype=np.uint8)

    def p(a, dt):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=dt)
        return a, a.ctypes.data
    keep = []
    args = []
    for key, dt in (("x", np.float64), ("v", np.float32), ("mass", np.float32),
                    ("h", np.float32), ("u", np.float32), ("id", np.int64),
                    ("time_bin", np.int8)):
        a, ptr = p(ic.get(key), dt); keep.append(a); args.append(ptr)
    a, ptr = p(tree.depth_h, np.int8); keep.append(a); args.append(ptr)
    for key in ("visc_alpha", "diff_alpha", "div_v_previous_step", "rho"):
        a, ptr = p(ic.get(key), np.float32); keep.append(a); args.append(ptr)
    perm = np.ascontiguousarray(tree.perm)
    host.swifthost_pack_parts(C.byref(layout), scheme, n, perm.ctypes.data, *args, out.ctypes.data)
    return out


# ---------------------------------------------------------------------------
# Synthetic initial conditions. Everything f32 except positions.
# `u` holds the thermal variable of the scheme: internal energy (Minimal,
# SPHENIX) or entropic function A = P / rho^gamma (Gadget2).
# ---------------------------------------------------------------------------

def _finish(ic, scheme, n, time_bin=None):
    ic["id"] = np.arange(1, n + 1, dtype=np.int64)
    ic["time_bin"] = np.full(n, 1, dtype=np.int8) if time_bin is None else time_bin.astype(np.int8)
    if scheme == abi.SCHEME_GADGET2:
        # gas_entropy_from_internal_energy (ideal_gas/equation_of_state.h)
        rho0 = ic["_rho0"]
        ic["u"] = ((HYDRO_GAMMA - 1.0) * ic["u"].astype(np.float64) * rho0 ** (1.0 - HYDRO_GAMMA)).astype(np.float32)
    if scheme == abi.SCHEME_SPHENIX:
        # hydro_first_init_part + hydro_convert_quantities (SPHENIX/hydro.h:1163-1226)
        ic["visc_alpha"] = np.full(n, 0.1, dtype=np.float32)
        ic["diff_alpha"] = np.zeros(n, dtype=np.float32)
        ic["div_v_previous_step"] = np.zeros(n, dtype=np.float32)
    return ic


def uniform_box(L=32, scheme=abi.SCHEME_MINIMAL, rho=2.0, P=1.0, eta=1.2349, box=1.0):
    """examples/HydroTests/UniformBox_3D/makeIC.py:28-105: lattice at cell
    centres, rho=2, P=1, v=0, h = eta * spacing."""
    n = L ** 3
    g = (np.arange(L) + 0.5) / L * box
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3)
    ic = {"x": x, "v": np.zeros((n, 3), np.float32),
          "mass": np.full(n, rho * box ** 3 / n, np.float32),
          "h": np.full(n, eta * box / L, np.float32),
          "u": np.full(n, P / ((HYDRO_GAMMA - 1.0) * rho), np.float32), "_rho0": rho}
    return _finish(ic, scheme, n)


def jittered_box(L, scheme, jitter=0.2, seed=42, eta=1.2348, box=1.0, rho=1.0,
                 u0=1.0, vamp=0.05, h_scatter=0.0, active_fraction=1.0, bricks=(1, 1, 1)):
    """Lattice + uniform jitter (+-jitter spacing); smooth solenoidal-ish
    velocity field; optionally scattered h (to exercise the ghost) and a
    multi-time-step active subset clustered in space. bricks=(bx,by,bz) makes
    a box of bx*by*bz unit bricks of L^3 particles each (weak scaling)."""
    rng = np.random.default_rng(seed)
    Ls = [L * b for b in bricks]
    n = Ls[0] * Ls[1] * Ls[2]
    gs = [(np.arange(m) + 0.5) / L for m in Ls]
    x = np.stack(np.meshgrid(*gs, indexing="ij"), axis=-1).reshape(n, 3)
    x = x + rng.uniform(-jitter, jitter, size=(n, 3)) / L
    x = np.mod(x, np.array(bricks, dtype=np.float64)) * box
    k = 2 * np.pi / box
    v = np.stack([np.sin(k * x[:, 1]) + 0.5 * np.cos(2 * k * x[:, 2]),
                  np.sin(k * x[:, 2]) + 0.5 * np.cos(2 * k * x[:, 0]),
                  np.sin(k * x[:, 0]) + 0.5 * np.cos(2 * k * x[:, 1])], axis=1) * vamp
    h = np.full(n, eta * box / L)
    if h_scatter > 0:
        h = h * np.exp(rng.uniform(-h_scatter, h_scatter, size=n))
    u = u0 * (1.0 + 0.1 * np.sin(k * x[:, 0]) * np.cos(k * x[:, 1]))
    ic = {"x": x, "v": v.astype(np.float32), "mass": np.full(n, rho * box ** 3 / L ** 3, np.float32),
          "h": h.astype(np.float32), "u": u.astype(np.float32), "_rho0": rho}
    tb = None
    if active_fraction < 1.0:
        # active fraction clustered in space: a smooth field picks the region
        f = np.sin(k * x[:, 0]) * np.sin(k * x[:, 1]) * np.sin(k * x[:, 2])
        thr = np.quantile(f, 1.0 - active_fraction)
        tb = np.where(f >= thr, 1, 3)
    return _finish(ic, scheme, n, tb)


def sedov_box(L=128, scheme=abi.SCHEME_GADGET2, seed=1234, eta=1.2348, E0=1.0, P0=1e-6, rho0=1.0,
              bricks=(1, 1, 1)):
    """SedovBlast_3D/makeIC.py:24-58 on a perturbed lattice: E0 shared by the
    15 particles nearest the centre."""
    ic = jittered_box(L, abi.SCHEME_MINIMAL, jitter=0.1, seed=seed, eta=eta, rho=rho0, vamp=0.0,
                      bricks=bricks)
    n = ic["x"].shape[0]
    u = np.full(n, P0 / ((HYDRO_GAMMA - 1.0) * rho0))
    r2 = ((ic["x"] - 0.5 * np.array(bricks, dtype=np.float64)) ** 2).sum(axis=1)
    centre = np.argsort(r2)[:15]
    u[centre] = E0 / (15 * ic["mass"][0])
    ic["u"] = u.astype(np.float32)
    ic["v"][:] = 0
    ic["_rho0"] = rho0
    for k in ("visc_alpha", "diff_alpha", "div_v_previous_step"):
        ic.pop(k, None)
    return _finish(ic, scheme, n)


def clustered_box(L=256, scheme=abi.SCHEME_SPHENIX, seed=2025, sigma=1.5, eta=1.2348):
    """Lognormal-clustered box: lattice displaced along the gradient of a
    Gaussian random potential (P(k) ~ k^-2 density field, sigma_ln(rho) ~ sigma);
    h from the local density estimate, deliberately imperfect so the ghost has
    to iterate."""
    rng = np.random.default_rng(seed)
    n = L ** 3
    ng = min(L, 64)
    kf = np.fft.fftfreq(ng) * ng
    kx, ky, kz = np.meshgrid(kf, kf, np.fft.rfftfreq(ng) * ng, indexing="ij")
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    k2[0, 0, 0] = 1.0
    delta_k = (rng.normal(size=k2.shape) + 1j * rng.normal(size=k2.shape)) / k2 ** 0.5
    delta_k[0, 0, 0] = 0
    delta_k *= np.exp(-k2 / (0.25 * ng) ** 2)
    delta = np.fft.irfftn(delta_k, s=(ng, ng, ng), axes=(0, 1, 2))
    delta *= sigma / delta.std()
    # displacement field psi = -grad(phi), lap(phi) = delta
    psi = [np.fft.irfftn(-1j * kk * delta_k / k2, s=(ng, ng, ng), axes=(0, 1, 2)) for kk in (kx, ky, kz)]
    scale = sigma / (np.fft.irfftn(delta_k, s=(ng, ng, ng), axes=(0, 1, 2)).std() + 1e-30)
    g = (np.arange(L) + 0.5) / L
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(n, 3)
    gi = np.minimum((x * ng).astype(np.int64), ng - 1)
    disp = np.stack([p[gi[:, 0], gi[:, 1], gi[:, 2]] for p in psi], axis=1) * scale
    disp *= 0.35 / (2 * np.pi)   # keep shell crossing mild
    x = np.mod(x + disp + rng.uniform(-0.2, 0.2, size=(n, 3)) / L, 1.0)
    # local density from a CIC-like count on a grid of ~8 particles per cell
    nb = max(L // 2, 4)
    bi = np.minimum((x * nb).astype(np.int64), nb - 1)
    cnt = np.zeros((nb, nb, nb))
    np.add.at(cnt, (bi[:, 0], bi[:, 1], bi[:, 2]), 1.0)
    dens = np.maximum(cnt[bi[:, 0], bi[:, 1], bi[:, 2]], 1.0) * nb ** 3 / n
    h = eta / L * dens ** (-1.0 / 3.0) * np.exp(rng.uniform(-0.15, 0.15, size=n))
    k = 2 * np.pi
    v = 0.05 * np.stack([np.sin(k * x[:, 1]), np.sin(k * x[:, 2]), np.sin(k * x[:, 0])], axis=1)
    ic = {"x": x, "v": v.astype(np.float32), "mass": np.full(n, 1.0 / n, np.float32),
          "h": h.astype(np.float32), "u": np.ones(n, np.float32), "_rho0": 1.0}
    return _finish(ic, scheme, n)


def _hash_uniform(idx, seed, stream):
    """Counter-based uniforms in [0, 1): splitmix64 of (global lattice index, seed, stream). The same
    particle gets the same numbers whichever rank generates it."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) * np.uint64(3) + np.uint64(stream) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def brick_box(L, scheme, grid=(1, 1, 1), rank=0, top=None, jitter=0.2, seed=42, eta=1.2348, rho=1.0, u0=1.0,
              vamp=0.05, active_fraction=1.0):
    """The rank's part of ONE periodic unit box of L^3 jittered-lattice particles split over
    grid[0] x grid[1] x grid[2] ranks (partition_uniform_grid, src/partition.c:104-121): the rank's own
    top-level cells plus one top-level cell of halo on every side (the foreign cells it holds proxies
    of). Jitter comes from a counter-based hash of the global lattice index, so every rank generates
    identical particles for the cells it shares with its neighbours and the union over the ranks is the
    same box whatever the grid is (grid (1,1,1): the whole box). Returns the ic dict; ic["id"] is the
    global lattice index + 1."""
    top = top if top is not None else default_top_grid(L)[0]
    assert L % top == 0 and all(top % g == 0 for g in grid)
    per = L // top  # lattice points per top-level cell and axis
    r3 = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
    axes = []
    for a in range(3):
        if grid[a] == 1:
            axes.append(np.arange(L, dtype=np.int64))
        else:
            lo = r3[a] * (top // grid[a]) * per - per
            hi = (r3[a] + 1) * (top // grid[a]) * per + per
            axes.append(np.mod(np.arange(lo, hi, dtype=np.int64), L))
    n0, n1, n2 = (len(ax) for ax in axes)
    n = n0 * n1 * n2
    x = np.empty((n, 3), np.float64)
    v = np.empty((n, 3), np.float32)
    u = np.empty(n, np.float32)
    gidx = np.empty(n, np.int64)
    tb = None
    thr = None
    k = 2 * np.pi
    if active_fraction < 1.0:
        # active region clustered in space: where a smooth field exceeds the level that holds the
        # requested fraction of the volume (|sin sin sin| > t; the fraction is measured on this sample)
        ref = np.sin(k * np.linspace(0, 1, 64, endpoint=False) + 0.01)
        fr = (ref[:, None, None] * ref[None, :, None] * ref[None, None, :]).reshape(-1)
        thr = np.quantile(fr, 1.0 - active_fraction)
        tb = np.empty(n, np.int64)
    JK_J = np.repeat(axes[1], n2)
    JK_K = np.tile(axes[2], n1)

    def fill(a, b):
        # slabs [a, b) of the first axis: element-wise work only, so the chunking changes nothing
        sl = slice(a * n1 * n2, b * n1 * n2)
        I = np.repeat(axes[0][a:b], n1 * n2)
        J = np.tile(JK_J, b - a)
        K = np.tile(JK_K, b - a)
        g = (I * L + J) * L + K
        gidx[sl] = g
        xs = x[sl]
        for c, ia in enumerate((I, J, K)):
            xs[:, c] = (ia + 0.5 + (2.0 * _hash_uniform(g, seed, c) - 1.0) * jitter) / L
        vs = v[sl]
        vs[:, 0] = (np.sin(k * xs[:, 1]) + 0.5 * np.cos(2 * k * xs[:, 2])) * vamp
        vs[:, 1] = (np.sin(k * xs[:, 2]) + 0.5 * np.cos(2 * k * xs[:, 0])) * vamp
        vs[:, 2] = (np.sin(k * xs[:, 0]) + 0.5 * np.cos(2 * k * xs[:, 1])) * vamp
        u[sl] = (u0 * (1.0 + 0.1 * np.sin(k * xs[:, 0]) * np.cos(k * xs[:, 1]))).astype(np.float32)
        if tb is not None:
            f = np.sin(k * xs[:, 0]) * np.sin(k * xs[:, 1]) * np.sin(k * xs[:, 2])
            tb[sl] = np.where(f >= thr, 1, 3)

    # numpy releases the GIL inside these kernels: slabs of the first axis on a few host threads
    nthreads = max(1, min(16, os.cpu_count() or 1, n0 // 4))
    step = max(1, min(8, -(-n0 // nthreads)))
    chunks = [(a, min(n0, a + step)) for a in range(0, n0, step)]
    if nthreads == 1:
        for a, b in chunks:
            fill(a, b)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda ab: fill(*ab), chunks))
    ic = {"x": x, "v": v, "mass": np.full(n, rho / L ** 3, np.float32),
          "h": np.full(n, eta / L, np.float32), "u": u, "_rho0": rho}
    ic = _finish(ic, scheme, n, tb)
    ic["id"] = gidx + 1
    return ic


def default_top_grid(L, max_top=32):
    """Top-level grid: cells at least 2*gamma*h*space_stretch wide
    (space_regrid.c:50-127) and at least 3 per axis; prefer ~4096 parts/cell."""
    c = max(3, min(max_top, L // 16))
    return (c, c, c)


def step_scalars(max_active_bin=56):
    """ti_current such that exactly the bins <= max_active_bin end their step
    now (timeline.h:126): an odd multiple of 2^(bin+1)."""
    if max_active_bin >= 56:
        return dict(ti_current=8, max_active_bin=56, time_base=1e-6)
    return dict(ti_current=3 * (1 << (max_active_bin + 1)), max_active_bin=max_active_bin, time_base=1e-6)


# ---------------------------------------------------------------------------
# Multi-GPU: what one rank holds. The reference keeps, on every rank, its own
# top-level cells plus a proxy copy of every foreign top-level cell that
# touches one of them (src/engine_proxy.c; a pair task exists iff at least one
# side is local, engine_maketasks.c:3562-3569).
# ---------------------------------------------------------------------------
def extract_rank(tree, parts_u8, layout, rank, periodic=True, local_first=False):
    # parts_u8 may be None: pack the returned sub-tree with pack_parts() instead
    """Sub-tree + particles of `rank`: local top-level cells and their foreign
    neighbours. Returns (Tree, parts_u8, sel, is_local) where sel[k] is the
    index in the global (cell-ordered) particle array of local particle k.
    local_first: the rank's own top-level cells (and hence its own particles)
    come first, the proxies after them - how SWIFT lays out space->parts - so
    that only [0, n_local) has to cross the host boundary."""
    cells, top = tree.cells, np.asarray(tree.top)
    ntop = top.shape[0]
    tc = cells[top]
    idx3 = np.floor(tc["loc"] / tc["width"] + 0.5).astype(np.int64)
    cdim = idx3.max(axis=0) + 1
    grid = -np.ones(tuple(cdim), dtype=np.int64)
    grid[idx3[:, 0], idx3[:, 1], idx3[:, 2]] = np.arange(ntop)
    local = tc["nodeID"] == rank
    keep = local.copy()
    li = idx3[local]
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                j = li + np.array([dx, dy, dz])
                if periodic:
                    j = np.mod(j, cdim)
                    ok = np.ones(len(j), bool)
                else:
                    ok = ((j >= 0) & (j < cdim)).all(axis=1)
                    j = np.clip(j, 0, cdim - 1)
                g = grid[j[:, 0], j[:, 1], j[:, 2]]
                keep[g[ok & (g >= 0)]] = True
    pos = -np.ones(cells.shape[0], dtype=np.int64)
    pos[top] = np.arange(ntop)
    cell_keep = keep[pos[cells["top"]]]
    new_index = np.cumsum(cell_keep) - 1
    new_index[~cell_keep] = -1
    sub = cells[cell_keep].copy()
    # particle ranges of the kept top-level cells, in `top` order (local ones first if asked)
    kept_top = top[keep]
    if local_first:
        kl = local[keep]
        kept_top = np.concatenate([kept_top[kl], kept_top[~kl]])
    counts = cells["count"][kept_top].astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    top_off = np.zeros(ntop, dtype=np.int64)
    top_off[pos[kept_top]] = offs
    top_first = cells["first_part"][top]
    tpos = pos[sub["top"]]
    sub["first_part"] = top_off[tpos] + (sub["first_part"] - top_first[tpos])
    for name in ("parent", "top"):
        v = sub[name]
        sub[name] = np.where(v >= 0, new_index[np.maximum(v, 0)], -1)
    pr = sub["progeny"]
    sub["progeny"] = np.where(pr >= 0, new_index[np.maximum(pr, 0)], -1)
    sel = np.concatenate([np.arange(f, f + c) for f, c in zip(cells["first_part"][kept_top], counts)]) if len(counts) else np.zeros(0, np.int64)
    size = layout.size
    sub_parts = None
    if parts_u8 is not None:
        sub_parts = np.ascontiguousarray(parts_u8.reshape(-1, size)[sel]).reshape(-1)
    new_top = new_index[kept_top].astype(np.int32)
    is_local = np.repeat(cells["nodeID"][kept_top] == rank, counts)
    return Tree(sub, new_top, tree.perm[sel], tree.depth_h[sel]), sub_parts, sel, is_local
