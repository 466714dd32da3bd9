"""Shared helpers for the parity tests: build matching Oracle / UdalesGPU pairs on seeded inputs."""
import numpy as np

from oracle.oracle import Oracle, stretched_zf

STATE = ("u0", "v0", "w0", "um", "vm", "wm", "pres0")


def make_pair(itot, jtot, ktot, stretched=True, seed_ir=43, gpu_flags=0, ubase=1.0, ltempeq_gpu=False, **kw):
    import udales_b200 as U
    zsize = ktot * (itot / 2.0) / itot
    zf = stretched_zf(ktot, zsize, 1.04) if stretched else None
    o = Oracle(itot, jtot, ktot, zf=zf, **kw)
    g = U.UdalesGPU(itot, jtot, ktot, zf=o.zf, flags=gpu_flags, ltempeq=ltempeq_gpu, **kw)
    o.init_channel(ubase=ubase, ir=seed_ir)
    # a non-trivial pressure field with consistent periodic halos
    rng = np.random.default_rng(seed_ir)
    o.pres0[1:-1, 1:-1, 1:-1] = 0.1 * rng.standard_normal((itot, jtot, ktot))
    o.pres0[0, :, :] = o.pres0[-2, :, :]; o.pres0[-1, :, :] = o.pres0[1, :, :]
    o.pres0[:, 0, :] = o.pres0[:, -2, :]; o.pres0[:, -1, :] = o.pres0[:, 1, :]
    push_state(o, g)
    return o, g


def push_state(o, g, names=STATE):
    for n in names:
        g.push(n, getattr(o, n))
    for n4 in range(o.nsv):
        g.push("sv0", o.sv0[..., n4], n4)
        g.push("svm", o.svm[..., n4], n4)


def interior(a):
    return a[1:-1, 1:-1, 1:-1]


def tend_interior(a):
    """(ib:ie, jb:je, kb:ke) of a tendency-shaped array (k starts at kb, one ghost level on top)."""
    return a[1:-1, 1:-1, :-1]


def sv_interior(a, hc):
    return a[hc:-hc, hc:-hc, hc:-hc]


def svp_interior(a, hc):
    return a[hc:-hc, hc:-hc, :-hc]


def relerr(a, b):
    s = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / s


def ibm_lists(I, J, K, boxes):
    """synthetic building blocks -> the eight point lists of src/modibm.f90 (same construction as
    oracle/f90run/make_golden.py:ibm_geometry): solid_c = cells inside a box (i0..i1, j0..j1, 1..k1); solid_u/v/w =
    staggered points touching a solid cell; bound_* = fluid points with a masked neighbour in the directions the
    diff*_corr routines look at."""
    sc = np.zeros((I + 2, J + 2, K + 2), dtype=bool)
    for (i0, i1, j0, j1, k1) in boxes:
        sc[i0:i1 + 1, j0:j1 + 1, 1:k1 + 1] = True
    su = sc | np.roll(sc, 1, axis=0)
    sv = sc | np.roll(sc, 1, axis=1)
    sw = sc.copy(); sw[:, :, 1:] |= sc[:, :, :-1]
    lists, masks = {}, {}
    for nm, sol in (("u", su), ("v", sv), ("w", sw), ("c", sc)):
        lists["solid_" + nm] = (np.argwhere(sol[1:I + 1, 1:J + 1, 1:K + 1]) + 1).astype(np.int32)
        m = np.ones((I + 2, J + 2, K + 2)); m[:, :, 0] = 0.0
        if nm == "w":
            m[:, :, 1] = 0.0
        m[1:I + 1, 1:J + 1, 1:K + 1][sol[1:I + 1, 1:J + 1, 1:K + 1]] = 0.0
        m[0] = m[I]; m[I + 1] = m[1]; m[:, 0] = m[:, J]; m[:, J + 1] = m[:, 1]
        masks[nm] = m
    dirs = {"u": ((0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)), "v": ((1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1)),
            "w": ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)),
            "c": ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))}
    for nm in "uvwc":
        m = masks[nm]
        fluid = m[1:I + 1, 1:J + 1, 1:K + 1] == 1.0
        touch = np.zeros_like(fluid)
        for a, b, c in dirs[nm]:
            touch |= m[1 + a:I + 1 + a, 1 + b:J + 1 + b, 1 + c:K + 1 + c] == 0.0
        pts = np.argwhere(fluid & touch) + 1
        if nm == "w":
            pts = pts[pts[:, 2] >= 2]
        lists["bound_" + nm] = pts.astype(np.int32)
    return lists


def add_thermo(o, g, seed=7, slab=None, **kw):
    """switch the dry temperature tier on in a matching Oracle / UdalesGPU pair (g constructed with ltempeq=True): same
    namelist values, a stably stratified thl0 with noise (deterministic in the GLOBAL index, so slabs agree), consistent
    halos / ghost levels, thlm = thl0, and the first thermodynamics() as the reference's startup does.
    slab: None, or (U.slab_of, world, rank) for an x-slab of the global oracle state."""
    args = dict(lbuoyancy=True, grav=9.81, thls=288.0, BCtopT=1, wttop=-0.01, thl_top=289.0, BCbotT=1, wtsurf=0.01)
    args.update(kw)
    I, J, K = o.itot, o.jtot, o.ktot
    rng = np.random.default_rng(seed)
    car = 1e-4 * rng.standard_normal(K + 1)
    o.set_thermo(thlpcar=car, **args)
    g.set_thermo(thlpcar=car, **args)
    o.ekm[...] = 1.5e-5; o.ekh[...] = 1.5e-5 / 0.71      # startup state: molecular values (fluxtop divides by ekh)
    o.thl0[...] = 0.0
    o.thl0[1:-1, 1:-1, 1:-1] = 288.0 + 0.02 * np.arange(1, K + 1)[None, None, :] + 0.2 * rng.standard_normal((I, J, K))
    o.thl0[:, :, 0] = o.thl0[:, :, 1]                    # src/modstartup.f90:1208
    o.halos(); o.boundary()
    o.thlm[...] = o.thl0
    cut = (lambda a: a) if slab is None else (lambda a: slab[0](a, slab[1], slab[2]))
    for n in ("ekm", "ekh", "thl0", "thlm"):
        g.push(n, cut(getattr(o, n)))
    o.thermodynamics(); g.thermodynamics()
