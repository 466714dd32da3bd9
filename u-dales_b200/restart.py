"""Restart-file adapter: the reference's unformatted `initd` / `inits` files <-> the resident GPU state.

Layout = what `writerestartfiles` writes and `readrestartfiles` reads (src/modsave.f90:78-122,
src/modstartup.f90:2156-2221): Fortran sequential unformatted records (4-byte little-endian length before and after
each record), one file per rank, named ``initd<ntrun:8>_<myidx:3>_<myidy:3>.<expnr:3>``:

    1  mindist (ib:ie, jb:je, kb:ke)                          real64
    2  wall    (ib:ie, jb:je, kb:ke, 1:5)                     int32
    3-12  u0 v0 w0 pres0 thl0 e120 ekm qt0 ql0 ql0h           real64, each (ib-ih:ie+ih, jb-jh:je+jh, kb:ke+kh)
    13 timee, dt                                              2 x real64
    inits: sv0 (ib-ih:ie+ih, jb-jh:je+jh, kb:ke+kh, 1:nsv) ; timee

With these a GPU run can be warm-started from a reference run (any nprocx x nprocy of the writer is re-assembled into
this rank's x-slab) and can leave files the reference warm-starts from.  Host-side I/O only: nothing here is on the
timed path."""
from __future__ import annotations

import os
import struct

import numpy as np

FIELDS_D = ("u0", "v0", "w0", "pres0", "thl0", "e120", "ekm", "qt0", "ql0", "ql0h")


def _records(path):
    raw = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(raw):
        n = struct.unpack("<i", raw[pos:pos + 4])[0]
        body = raw[pos + 4:pos + 4 + n]
        if len(body) != n or struct.unpack("<i", raw[pos + 4 + n:pos + 8 + n])[0] != n:
            raise ValueError(f"{path}: broken Fortran record at byte {pos}")
        out.append(body)
        pos += 8 + n
    return out


def _write_records(path, recs):
    with open(path, "wb") as f:
        for r in recs:
            b = r if isinstance(r, (bytes, bytearray)) else np.asfortranarray(r).tobytes(order="F")
            f.write(struct.pack("<i", len(b))); f.write(b); f.write(struct.pack("<i", len(b)))


def restart_name(kind, ntrun, myidx, myidy, expnr):
    """'initd' / 'inits' file name of one rank (src/modsave.f90:81-85)"""
    return f"init{kind}{int(ntrun):08d}_{int(myidx):03d}_{int(myidy):03d}.{int(expnr):03d}"


def read_initd(path, imax, jmax, ktot, ih=1, jh=1, kh=1):
    """one rank's initd file -> dict: the ten 3-D fields with shape (imax+2ih, jmax+2jh, ktot+kh) (k = kb .. ke+kh),
    mindist (imax, jmax, ktot), wall (imax, jmax, ktot, 5) int32, timee, dt"""
    rec = _records(path)
    if len(rec) != 13:
        raise ValueError(f"{path}: {len(rec)} records, expected 13 (src/modsave.f90:87-99)")
    shp = (imax + 2 * ih, jmax + 2 * jh, ktot + kh)
    n3 = int(np.prod(shp)) * 8
    if len(rec[0]) != imax * jmax * ktot * 8 or any(len(r) != n3 for r in rec[2:12]):
        raise ValueError(f"{path}: record sizes do not match a {imax}x{jmax}x{ktot} pencil with halo {ih},{jh},{kh}")
    out = {"mindist": np.frombuffer(rec[0], "<f8").reshape((imax, jmax, ktot), order="F").copy(),
           "wall": np.frombuffer(rec[1], "<i4").reshape((imax, jmax, ktot, 5), order="F").copy()}
    for q, nm in enumerate(FIELDS_D):
        out[nm] = np.frombuffer(rec[2 + q], "<f8").reshape(shp, order="F").copy()
    out["timee"], out["dt"] = struct.unpack("<2d", rec[12])
    return out


def write_initd(path, fields, timee, dt):
    """fields: dict with the ten 3-D arrays (missing ones are written as zeros, like a neutral run leaves thl0 / qt0 ...),
    optional mindist / wall"""
    ref = np.asarray(fields["u0"])
    I, J, K = ref.shape[0] - 2, ref.shape[1] - 2, ref.shape[2] - 1
    recs = [np.asarray(fields.get("mindist", np.zeros((I, J, K))), dtype="<f8"),
            np.asarray(fields.get("wall", np.zeros((I, J, K, 5), dtype=np.int32)), dtype="<i4")]
    for nm in FIELDS_D:
        a = np.asarray(fields.get(nm, np.zeros(ref.shape)), dtype="<f8")
        if a.shape != ref.shape:
            raise ValueError(f"{nm}: shape {a.shape}, expected {ref.shape}")
        recs.append(a)
    recs.append(struct.pack("<2d", float(timee), float(dt)))
    _write_records(path, recs)


def read_inits(path, imax, jmax, ktot, nsv, ih=1, jh=1, kh=1):
    rec = _records(path)
    shp = (imax + 2 * ih, jmax + 2 * jh, ktot + kh, nsv)
    return np.frombuffer(rec[0], "<f8").reshape(shp, order="F").copy(), struct.unpack("<d", rec[1])[0]


def write_inits(path, sv0, timee):
    _write_records(path, [np.asarray(sv0, dtype="<f8"), struct.pack("<d", float(timee))])


def assemble(directory, ntrun, expnr, itot, jtot, ktot, nprocx, nprocy, names=("u0", "v0", "w0", "pres0", "ekm", "thl0")):
    """the files of an nprocx x nprocy reference run -> global arrays (itot+2, jtot+2, ktot+1) with the periodic /
    domain halo taken from the edge ranks (interior from every rank).  Returns (dict, timee, dt)."""
    imax, jmax = itot // nprocx, jtot // nprocy
    glob = {nm: np.zeros((itot + 2, jtot + 2, ktot + 1), order="F") for nm in names}
    meta = (0.0, 0.0)
    for px in range(nprocx):
        for py in range(nprocy):
            d = read_initd(os.path.join(directory, restart_name("d", ntrun, px, py, expnr)), imax, jmax, ktot)
            i0, j0 = 1 + px * imax, 1 + py * jmax
            for nm in names:
                a, gl = d[nm], glob[nm]
                gl[i0:i0 + imax, j0:j0 + jmax, :] = a[1:-1, 1:-1, :]
                if px == 0: gl[0, j0:j0 + jmax, :] = a[0, 1:-1, :]
                if px == nprocx - 1: gl[itot + 1, j0:j0 + jmax, :] = a[-1, 1:-1, :]
                if py == 0: gl[i0:i0 + imax, 0, :] = a[1:-1, 0, :]
                if py == nprocy - 1: gl[i0:i0 + imax, jtot + 1, :] = a[1:-1, -1, :]
            meta = (d["timee"], d["dt"])
    return glob, meta[0], meta[1]


def load_into(g, glob, timee=0.0, dt=0.0):
    """push a global restart state (arrays (itot+2, jtot+2, ktot+1), k = kb .. ke+kh) into a UdalesGPU handle's x-slab:
    level k = kb-1 is zero (never written by the reference either, src/modstartup.f90:1155-1176), halos() / boundary()
    re-establish the ghost cells, and um, vm, wm start as copies (src/modstartup.f90:1233-1244 after a warm start)."""
    imax, lo = g.imax, g.myidx * g.imax
    thermo = bool(g.cfg.ltempeq) and "thl0" in glob
    for nm in ("u0", "v0", "w0", "pres0", "ekm") + (("thl0",) if thermo else ()):
        if nm not in glob:
            continue
        a = np.zeros(g.shape(nm), order="F")
        a[:, :, 1:] = glob[nm][lo:lo + imax + 2, :, :]
        if nm == "thl0":
            a[:, :, 0] = a[:, :, 1]       # thl0(kb-1) = thl0(kb), src/modstartup.f90:1208
        g.push(nm, a)
    if thermo and "ekm" not in glob:
        # boundary()'s fluxtop divides by ekh: the reference's startup has run closure by then; molecular values stand in
        ek = np.full(g.shape("ekh"), float(g.cfg.numol) * float(g.cfg.prandtlmoli), order="F")
        g.push("ekh", ek)
    g.halos(); g.boundary()
    for nm in ("u0", "v0", "w0") + (("thl0",) if thermo else ()):
        g.push(nm.replace("0", "m"), g.pull(nm))
    if thermo:
        g.thermodynamics()                # src/modstartup.f90 calls thermodynamics before the time loop
    g.dt = dt
    g.rk3step = 0
    return timee


def save_from(g, directory, ntrun, expnr, timee, extra=None):
    """write this rank's initd file (x-slab: myidy = 0) in the reference's layout; thl0, e120, qt0, ql0, ql0h are zeros
    (neutral run) unless given in `extra`"""
    f = {nm: g.pull(nm)[:, :, 1:] for nm in ("u0", "v0", "w0", "pres0", "ekm") + (("thl0",) if g.cfg.ltempeq else ())}
    f.update(extra or {})
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, restart_name("d", ntrun, g.myidx, 0, expnr))
    write_initd(path, f, timee, g.dt)
    if g.nsv:
        # scalar arrays carry halo ihc (2 with kappa); the file holds (ib-ih:ie+ih, jb-jh:je+jh, kb:ke+kh) with ih = jh = kh = 1
        hc = (g.shape("sv0")[0] - g.imax) // 2
        sv = np.stack([g.pull("sv0", n)[hc - 1:hc + g.imax + 1, hc - 1:hc + g.jtot + 1, hc:hc + g.ktot + 1] for n in range(g.nsv)], axis=3)
        write_inits(os.path.join(directory, restart_name("s", ntrun, g.myidx, 0, expnr)), sv, timee)
    return path
