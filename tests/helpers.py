"""Shared helpers for the parity tests: build matching Oracle / UdalesGPU pairs on seeded inputs."""
import numpy as np

from oracle.oracle import Oracle, stretched_zf

STATE = ("u0", "v0", "w0", "um", "vm", "wm", "pres0")


def make_pair(itot, jtot, ktot, stretched=True, seed_ir=43, gpu_flags=0, **kw):
    import udales_b200 as U
    zsize = ktot * (itot / 2.0) / itot
    zf = stretched_zf(ktot, zsize, 1.04) if stretched else None
    o = Oracle(itot, jtot, ktot, zf=zf, **kw)
    g = U.UdalesGPU(itot, jtot, ktot, zf=o.zf, flags=gpu_flags, **kw)
    o.init_channel(ir=seed_ir)
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
