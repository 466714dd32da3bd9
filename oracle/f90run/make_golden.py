#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S FORTRAN SOURCE TEXT with oracle/f90run/interp.py.

Run in the build container only (needs /root/reference):   python oracle/f90run/make_golden.py
Nothing from the reference is copied: the .f90 files are read where they lie, interpreted, and only
seeded inputs + numerical outputs are stored.  Module state that the hot path reads (grid metrics,
switches) is set up here following src/modglobal.f90:708-870 / namelist defaults; MPI, 2decomp-fft
and FFTW calls are replaced by single-pencil Python callbacks (halo exchange = no-op on one rank with
non-periodic communicators, transposes = copies, FFTW r2c/c2r = numpy.fft).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
from interp import FArray, Interp  # noqa: E402

REF = os.environ.get("UDALES_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(ROOT, "tests", "golden")


def fa(shape_bounds, fill=0.0, dtype=float):
    return FArray.alloc(shape_bounds, dtype, fill)


def stretched_zf(ktot, zsize, ratio):
    dz = ratio ** np.arange(ktot)
    dz *= zsize / dz.sum()
    zh = np.concatenate([[0.0], np.cumsum(dz)])
    return 0.5 * (zh[:-1] + zh[1:])


class World:
    """module variables of modglobal / modfields / modsubgriddata / modpois / decomp_2d for one pencil."""

    def __init__(self, itot, jtot, ktot, xlen, ylen, zf, nsv=0, iadv_sv=7, BCtopm=1, lvreman=True, lsmagorinsky=False,
                 Uinf=0.0, Vinf=0.0, seed=0):
        g = {}
        self.g = g
        ib = jb = kb = 1
        ie, je, ke = itot, jtot, ktot
        ih = jh = kh = 1
        hc = 2 if (nsv > 0 and iadv_sv == 7) else 1
        g.update(ib=ib, ie=ie, jb=jb, je=je, kb=kb, ke=ke, ih=ih, jh=jh, kh=kh, ihc=hc, jhc=hc, khc=hc,
                 imax=itot, jmax=jtot, kmax=ktot, itot=itot, jtot=jtot, ktot=ktot, nsv=nsv)
        K = ktot
        # ---- metrics: src/modglobal.f90:708-870 -------------------------------------------------
        dx, dy = xlen / float(itot), ylen / float(jtot)
        zfA = fa([(kb, ke + kh)]); zhA = fa([(kb, ke + kh)])
        dzf = fa([(kb - kh, ke + kh)]); dzh = fa([(kb, ke + kh)])
        for k in range(1, K + 1):
            zfA.a[k - 1] = zf[k - 1]
        zhA.a[0] = 0.0
        for k in range(1, K + 1):
            zhA.a[k] = zhA.a[k - 1] + 2.0 * (zfA.a[k - 1] - zhA.a[k - 1])
        zfA.a[K] = zfA.a[K - 1] + 2.0 * (zhA.a[K] - zfA.a[K - 1])
        for k in range(1, K + 1):
            dzf.a[k] = zhA.a[k] - zhA.a[k - 1]
        dzf.a[K + 1] = dzf.a[K]
        dzf.a[0] = dzf.a[1]
        dzh.a[0] = 2 * zfA.a[0]
        for k in range(2, K + 2):
            dzh.a[k - 1] = zfA.a[k - 1] - zfA.a[k - 2]
        dxf = fa([(ib - ih, itot + ih)], dx)
        delta = fa([(ib - ih, itot + ih), (kb, ke + kh)])
        for k in range(1, K + 2):
            delta.a[:, k - 1] = (dx * dy * dzf.a[k]) ** (1. / 3.)
        dzhi = FArray(1. / dzh.a, dzh.lb); dzfi = FArray(1. / dzf.a, dzf.lb)
        g.update(dx=dx, dy=dy, zf=zfA, zh=zhA, dzf=dzf, dzh=dzh, dxf=dxf, delta=delta,
                 dzhi=dzhi, dzfi=dzfi, dzf2=FArray(dzf.a * dzf.a, dzf.lb),
                 dxi=1. / dx, dyi=1. / dy, dx2=dx * dx, dy2=dy * dy,
                 dzhiq=FArray(0.25 * dzhi.a, dzh.lb), dzfiq=FArray(0.25 * dzfi.a, dzf.lb),
                 dzh2i=FArray(dzhi.a * dzhi.a, dzh.lb), dzfi5=FArray(0.5 * dzfi.a, dzf.lb))
        g["dxiq"] = 0.25 * g["dxi"]; g["dyiq"] = 0.25 * g["dyi"]
        g["dx2i"] = g["dxi"] * g["dxi"]; g["dy2i"] = g["dyi"] * g["dyi"]
        g["dxi5"] = 0.5 * g["dxi"]; g["dyi5"] = 0.5 * g["dyi"]
        # kappa grids (:841-870)
        dzfc = fa([(kb - hc, ke + hc)]); dzhci = fa([(kb - 1, ke + hc)])
        dxfc = fa([(ib - hc, itot + hc)], dx); dxhci = fa([(ib - 1, itot + hc)], 1. / dx)
        dzfc.a[hc - kh:hc - kh + K + 2 * kh] = dzf.a
        dzfc.a[0] = dzfc.a[hc - kh]; dzfc.a[-1] = dzfc.a[-1 - (hc - kh)]
        dzhci.a[1:1 + K + kh] = dzhi.a
        dzhci.a[0] = dzhci.a[1]
        if hc > kh:
            dzhci.a[-1] = dzhci.a[-2]
        g.update(dzfc=dzfc, dzfci=FArray(1. / dzfc.a, dzfc.lb), dzhci=dzhci, dxfc=dxfc, dxfci=FArray(1. / dxfc.a, dxfc.lb), dxhci=dxhci)
        # ---- constants & switches (namelist defaults, src/modglobal.f90, src/modsubgriddata.f90) -----
        g.update(numol=1.5e-5, prandtlmoli=1. / 0.71, eps1=1.e-10, pi=3.141592653589793116, e12min=5.e-5, grav=9.81,
                 lles=bool(lvreman or lsmagorinsky), lmoist=False, ltempeq=False, lbuoyancy=False, lchem=False,
                 iadv_mom=2, iadv_tke=2, iadv_thl=2, iadv_qt=2, iadv_cd2=2, iadv_kappa=7, iadv_upw=1,
                 iadv_sv=FArray(np.full(100, iadv_sv if nsv else -1, dtype=int), [1]),
                 bcxm=1, bcym=1, bcxt=1, bcyt=1, bcxq=1, bcyq=1, bcxs=1, bcys=1, bctopm=BCtopm, bctopt=1, bctopq=1, bctops=1, bczp=1,
                 bcxm_periodic=1, bcxm_profile=2, bcxm_driver=3, bcym_periodic=1, bcym_profile=2,
                 bcxt_periodic=1, bcxt_profile=2, bcxt_driver=3, bcyt_periodic=1, bcyt_profile=2,
                 bcxq_periodic=1, bcxq_profile=2, bcxq_driver=3, bcyq_periodic=1, bcyq_profile=2,
                 bcxs_periodic=1, bcxs_profile=2, bcxs_driver=3, bcxs_custom=4, bcys_periodic=1,
                 bctopm_freeslip=1, bctopm_noslip=2, bctopm_pressure=3, bctopt_flux=1, bctopt_value=2,
                 bctopq_flux=1, bctopq_value=2, bctops_flux=1, bctops_value=2,
                 ibrank=True, ierank=True, jbrank=True, jerank=True, uinf=Uinf, vinf=Vinf,
                 ipoiss=0, poiss_fft2d=0, poiss_cyc=1, poiss_fft3d=2, poiss_fft2d_2decomp=3,
                 rk3step=0, dt=0.0, dtmax=1e9, courant=1.0, diffnr=0.25, ladaptive=False, lwarmstart=True,
                 timee=0.0, timeleft=1e9, ntimee=0, ntrun=0, dt_lim=0.0, ifixuinf=0, iinletgen=0, idriver=0,
                 luoutflowr=True, luvolflowr=False, lchunkread=False, thlsrc=0.0, dgdt=0.0, myid=0, cmyid="000",
                 nrank=0, comm3d=0, mpierr=0, my_real=0, mpi_max=1, mpi_sum=2,
                 lsmagorinsky=lsmagorinsky, lvreman=lvreman, loneeqn=False, lbuoycorr=False, ldelta=False,
                 c_vreman=0.07, cs=-1.0, prandtli=1. / 0.333, dampmin=1e-10, rigc=0.25,
                 cm=0.12, cn=0.76, ch1=1., ch2=2., thvs=300.0,
                 fftw_measure=0, fftw_redft10=5, fftw_redft01=4,
                 thl_top=0.0, qt_top=0.0, wttop=0.0, wqtop=0.0, ubulk=0.0, vbulk=0.0, uouttot=0.0, vouttot=0.0)
        cf, alpha_kolm = 2.5, 1.5
        cm = cf / (2. * g["pi"]) * (1.5 * alpha_kolm) ** (-1.5)
        ceps = 2. * g["pi"] / cf * (1.5 * alpha_kolm) ** (-1.5)
        g["csz"] = fa([(ib - ih, ie + ih), (kb, ke + kh)], (cm ** 3 / ceps) ** 0.25)      # src/modsubgrid.f90:72-76
        g["wsvtop"] = fa([(1, max(nsv, 1))], 0.0); g["sv_top"] = fa([(1, max(nsv, 1))], 0.0)
        g["zsize"] = FArray(np.array([itot, jtot, ktot]), [1]); g["zstart"] = FArray(np.array([1, 1, 1]), [1])
        g["xsize"] = g["zsize"]; g["ysize"] = g["zsize"]
        # ---- fields: src/modfields.f90:440-520 -----------------------------------------------------
        full = [(ib - ih, ie + ih), (jb - jh, je + jh), (kb - kh, ke + kh)]
        tend = [(ib - ih, ie + ih), (jb - jh, je + jh), (kb, ke + kh)]
        for nm in ("u0", "v0", "w0", "um", "vm", "wm", "pres0", "ekm", "ekh", "thl0", "thlm", "qt0", "qtm", "e120", "e12m", "dthvdz"):
            g[nm] = fa(full)
        g["thl0c"] = fa([(ib - hc, ie + hc), (jb - hc, je + hc), (kb - hc, ke + hc)])
        for nm in ("up", "vp", "wp", "thlp", "qtp", "e12p", "zlt"):
            g[nm] = fa(tend)
        g["thlpc"] = fa([(ib - hc, ie + hc), (jb - hc, je + hc), (kb, ke + hc)])
        n4 = max(nsv, 1)
        g["sv0"] = fa([(ib - hc, ie + hc), (jb - hc, je + hc), (kb - hc, ke + hc), (1, n4)])
        g["svm"] = fa([(ib - hc, ie + hc), (jb - hc, je + hc), (kb - hc, ke + hc), (1, n4)])
        g["svp"] = fa([(ib - hc, ie + hc), (jb - hc, je + hc), (kb, ke + hc), (1, n4)])
        g["damp"] = fa([(ib, ie), (jb, je), (kb, ke)], 1.0)
        g["rhobf"] = fa([(kb, ke + kh)], 1.0); g["rhobh"] = fa([(kb, ke + kh)], 1.0)   # src/modfields.f90:571-572
        g["u0av"] = fa([(kb, ke + kh)]); g["v0av"] = fa([(kb, ke + kh)]); g["dpdxl"] = fa([(kb, ke + kh)])
        # modpois module variables (allocated by initpois)
        for nm in ("p", "pup", "pvp", "pwp", "rhs", "dpupdx", "dpvpdy", "dpwpdz", "fxy", "fxyz", "xrt", "yrt", "zrt", "xyzrt", "bxyzrt",
                   "a", "b", "c", "sxr", "sxfr", "syr", "syfr", "szr", "szfr", "sxfc", "syfc", "dpdztop", "pij",
                   "plan_r2fc_x", "plan_r2fc_y", "plan_fc2r_x", "plan_fc2r_y", "kbc1", "kbc2"):
            g[nm] = None
        self.plans = {}

    # -------------------------------------------------------------------------------------------
    def externals(self):
        w = self

        def alloc_pencil(it, frame, vals, setters, kw):
            lev = [0, 0, 0]
            for k in ("opt_xlevel", "opt_ylevel", "opt_zlevel"):
                if k in kw:
                    lev = [int(x) for x in kw[k]]
            if not kw:
                g = w.g
                lev = [g["ih"], g["jh"], g["kh"]]          # decomp_main%zlevel (src/modglobal.f90:638)
            g = w.g
            b = [(1 - lev[0], g["itot"] + lev[0]), (1 - lev[1], g["jtot"] + lev[1]), (1 - lev[2], g["ktot"] + lev[2])]
            setters[0](fa(b, 0.0))

        def transpose(it, frame, vals, setters, kw):
            vals[1].a[...] = vals[0].a

        def nop(it, frame, vals, setters, kw):
            pass

        def plan_r2c(it, frame, vals, setters, kw):
            pid = len(w.plans) + 1
            w.plans[pid] = ("r2c", vals[1], vals[2], vals[3])
            setters[0](pid)

        def plan_c2r(it, frame, vals, setters, kw):
            pid = len(w.plans) + 1
            w.plans[pid] = ("c2r", vals[1], vals[2], vals[3])
            setters[0](pid)

        def execute(it, frame, vals, setters, kw):
            kind, n, a, b = w.plans[vals[0]]
            if kind == "r2c":
                b.a[...] = np.fft.rfft(a.a)                       # FFTW r2c: unnormalised, sign -
            else:
                b.a[...] = np.fft.irfft(a.a, n) * n               # FFTW c2r: unnormalised

        def allreduce(it, frame, vals, setters, kw):
            setters[1](vals[0])

        return {"alloc_x": alloc_pencil, "alloc_y": alloc_pencil, "alloc_z": alloc_pencil,
                "transpose_x_to_y": transpose, "transpose_y_to_x": transpose, "transpose_y_to_z": transpose,
                "transpose_z_to_y": transpose, "exchange_halo_z": nop, "mpi_bcast": nop, "barrou": nop,
                "dfftw_plan_dft_r2c_1d": plan_r2c, "dfftw_plan_dft_c2r_1d": plan_c2r, "dfftw_execute": execute,
                "mpi_allreduce": allreduce}


def make_interp(w):
    it = Interp(w.g, w.externals())
    it.gtypes = {"sxfc": complex, "syfc": complex}
    for f in ("modadvection.f90", "modsubgrid.f90", "modpois.f90", "modtstep.f90", "modboundary.f90", "modchecksim.f90"):
        it.load(os.path.join(SRC, f))
    return it


def seed_fields(w, seed, nsv):
    g = w.g
    rng = np.random.default_rng(seed)
    I, J, K = g["itot"], g["jtot"], g["ktot"]
    for nm, base, amp in (("u0", 1.0, 0.3), ("v0", 0.2, 0.3), ("w0", 0.0, 0.3), ("pres0", 0.0, 0.5)):
        g[nm].a[...] = 0.0
        g[nm].a[1:-1, 1:-1, 1:-1] = base + amp * rng.standard_normal((I, J, K))
    g["w0"].a[:, :, 1] = 0.0
    hc = g["ihc"]
    for n in range(nsv):
        g["sv0"].a[..., n] = 0.0
        g["sv0"].a[hc:-hc, hc:-hc, hc:-hc, n] = 1.0 + 0.2 * rng.standard_normal((I, J, K))


def snapshot(w, names):
    return {n: np.array(w.g[n].a, copy=True) for n in names}


def case_substeps(tag, nsub=3, **kw):
    """the in-scope part of src/program.f90:132-207, executed from the reference text."""
    I, J, K = kw.pop("shape")
    nsv = kw.get("nsv", 0)
    zf = stretched_zf(K, 0.5 * K * 1.1, 1.07)
    w = World(I, J, K, xlen=0.55 * I, ylen=0.45 * J, zf=zf, **kw)
    it = make_interp(w)
    g = w.g
    it.call("initpois")
    seed_fields(w, 11, nsv)
    g["ekm"].a[...] = g["numol"]                      # startup state of the reference: molecular values
    g["ekh"].a[...] = g["numol"] * g["prandtlmoli"]   # (fluxtopscal divides by ekh, src/modboundary.f90:1530)
    it.call("halos"); it.call("boundary")
    for a, b in (("um", "u0"), ("vm", "v0"), ("wm", "w0")):
        g[a].a[...] = g[b].a
    g["svm"].a[...] = g["sv0"].a
    out = {"zf": zf, "shape": np.array([I, J, K]), "xlen": 0.55 * I, "ylen": 0.45 * J, "nsv": nsv,
           "BCtopm": g["bctopm"], "lvreman": int(g["lvreman"]), "lsmagorinsky": int(g["lsmagorinsky"]),
           "iadv_sv": int(g["iadv_sv"].a[0]), "Uinf": g["uinf"], "Vinf": g["vinf"]}
    state = ["u0", "v0", "w0", "um", "vm", "wm", "pres0", "sv0", "svm"]
    for k_, v_ in snapshot(w, state).items():
        out["in_" + k_] = v_
    for nm in ("xrt", "yrt", "a", "b", "c", "bxyzrt"):
        out["pois_" + nm] = np.array(g[nm].a, copy=True)
    dt = 0.03
    g["dt"] = dt
    g["dtmax"] = dt
    g["ladaptive"] = False
    for s in range(nsub):
        it.call("tstep_update")
        it.call("advection")
        if s == 0:
            for k_, v_ in snapshot(w, ["up", "vp", "wp", "svp"]).items():
                out["adv_" + k_] = v_
        it.call("subgrid")
        if s == 0:
            for k_, v_ in snapshot(w, ["up", "vp", "wp", "svp", "ekm", "ekh"]).items():
                out["sub_" + k_] = v_
        it.call("poisson")
        if s == 0:
            for k_, v_ in snapshot(w, ["p", "up", "vp", "wp", "pres0", "rhs"]).items():
                out["pois_" + k_] = v_
        it.call("tstep_integrate")
        it.call("halos")
        it.call("boundary")
        for k_, v_ in snapshot(w, state).items():
            out[f"s{s + 1}_" + k_] = v_
        out[f"s{s + 1}_rk3step"] = g["rk3step"]
    # chkdiv formula (src/modchecksim.f90:182-188) evaluated by the reference text: capture through allreduce
    caught = {}
    ext = it.ext

    def allreduce(it_, frame, vals, setters, kw_):
        setters[1](vals[0])
        caught[len(caught)] = vals[0]
    ext["mpi_allreduce"] = allreduce
    it.call("chkdiv")
    out["divtot"], out["divmax"] = caught[0], caught[1]
    # adaptive time step from the reference text
    g["ladaptive"] = True; g["rk3step"] = 0; g["dt"] = 0.05; g["dtmax"] = 2.0; g["courant"] = 1.1; g["diffnr"] = 0.25
    caught.clear()
    it.call("tstep_update")
    out["adapt_courtot"], out["adapt_diffnrtot"], out["adapt_dt"] = caught[0], caught[1], g["dt"]
    for k_, v_ in out.items():
        if isinstance(v_, np.ndarray) and v_.dtype.kind == "f" and not k_.startswith("in_") and "bxyzrt" not in k_:
            # interior of every output must be finite
            assert np.isfinite(v_[tuple(slice(2, -2) for _ in range(min(3, v_.ndim)))]).all(), k_
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
    print(f"wrote ref_{tag}.npz  ({len(out)} arrays)  final |u0|max {np.abs(g['u0'].a).max():.6f} divmax {out['divmax']:.2e}")


def ibm_geometry(I, J, K, boxes):
    """Synthetic building blocks -> the eight point lists of src/modibm.f90 (local 1-based i,j,k).
    solid_c: cell centres inside a box; solid_u / v / w: staggered points touching a solid cell;
    fluid-boundary lists: interior fluid points with at least one masked neighbour in the directions the
    corresponding diff*_corr routine looks at (incl. the ground level kb-1 and, for w, level kb)."""
    sc = np.zeros((I + 2, J + 2, K + 2), dtype=bool)
    for (i0, i1, j0, j1, k1) in boxes:
        sc[i0:i1 + 1, j0:j1 + 1, 1:k1 + 1] = True
    su = sc | np.roll(sc, 1, axis=0)            # u(i) sits between cells i-1 and i
    sv = sc | np.roll(sc, 1, axis=1)
    sw = sc.copy(); sw[:, :, 1:] |= sc[:, :, :-1]
    lists, masks = {}, {}
    for nm, sol in (("u", su), ("v", sv), ("w", sw), ("c", sc)):
        pts = np.argwhere(sol[1:I + 1, 1:J + 1, 1:K + 1]) + 1
        lists["solid_" + nm] = pts.astype(np.int32)
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
        out = []
        klo = 2 if nm == "w" else 1
        for k in range(klo, K + 1):
            for j in range(1, J + 1):
                for i in range(1, I + 1):
                    if m[i, j, k] == 1.0 and any(m[i + a, j + b, k + c] == 0.0 for a, b, c in dirs[nm]):
                        out.append((i, j, k))
        lists["bound_" + nm] = np.array(out, dtype=np.int32).reshape(-1, 3)
    return lists


def case_ibm(tag, shape=(12, 10, 8), nsv=1):
    """solid / ibmnorm / diffu,v,w,c_corr executed from src/modibm.f90."""
    I, J, K = shape
    zf = stretched_zf(K, 0.5 * K * 1.1, 1.07)
    w = World(I, J, K, xlen=0.55 * I, ylen=0.45 * J, zf=zf, nsv=nsv, iadv_sv=7)
    it = make_interp(w)
    it.load(os.path.join(SRC, "modibm.f90"), only=["solid", "diffu_corr", "diffv_corr", "diffw_corr", "diffc_corr", "ibmnorm"])
    g = w.g
    lists = ibm_geometry(I, J, K, [(4, 6, 3, 5, 3), (9, 10, 7, 8, 2), (1, 2, 9, 10, 4)])
    full = [(0, I + 1), (0, J + 1), (0, K + 1)]
    tend = [(0, I + 1), (0, J + 1), (1, K + 1)]
    for nm in "uvwc":
        pts = lists["solid_" + nm]
        g["solid_info_" + nm] = {"nsolptsrank": int(pts.shape[0]), "solpts_loc": FArray(np.asfortranarray(pts.astype(int)), [1, 1])}
        b = lists["bound_" + nm]
        g["bound_info_" + nm] = {"nbndptsrank": int(b.shape[0]), "bndpts_loc": FArray(np.asfortranarray(b.astype(int)), [1, 1])}
    # masks exactly as initibm builds them (src/modibm.f90:153-192), `solid` interpreted from the reference text
    dummy = fa(tend)
    for nm in "uvwc":
        m = fa(full, 1.0)
        m.a[:, :, 0] = 0.0
        if nm == "w":
            m.a[:, :, 1] = 0.0
        it.call("solid", g["solid_info_" + nm], m, dummy, 0.0, 1, 1, 1)
        a = m.a
        a[0] = a[I]; a[I + 1] = a[1]; a[:, 0] = a[:, J]; a[:, J + 1] = a[:, 1]     # exchange_halo_z on one periodic pencil
        g["mask_" + nm] = m
    g.update(libm=True, lconservativeibm=False, thl0av=fa([(1, K + 1)]))
    rng = np.random.default_rng(23)
    seed_fields(w, 17, nsv)
    for nm in ("um", "vm", "wm"):
        g[nm].a[...] = g[nm.replace("m", "0")].a + 0.05 * rng.standard_normal(g[nm].a.shape)
    g["svm"].a[...] = g["sv0"].a + 0.05 * rng.standard_normal(g["sv0"].a.shape)
    g["sv0"].a[...] += 0.01 * rng.standard_normal(g["sv0"].a.shape)     # halos included: neighbours of boundary points are read
    for nm in ("up", "vp", "wp", "svp"):
        g[nm].a[...] = rng.standard_normal(g[nm].a.shape)
    g["ekm"].a[...] = 1e-3 * (1.0 + rng.random(g["ekm"].a.shape))
    g["ekh"].a[...] = 3e-3 * (1.0 + rng.random(g["ekh"].a.shape))
    out = {"zf": zf, "shape": np.array([I, J, K]), "xlen": 0.55 * I, "ylen": 0.45 * J, "nsv": nsv}
    for k_, v_ in lists.items():
        out["pts_" + k_] = v_
    for nm in "uvwc":
        out["mask_" + nm] = np.array(g["mask_" + nm].a, copy=True)
    state = ["u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh", "sv0", "svm", "svp"]
    for k_, v_ in snapshot(w, state).items():
        out["in_" + k_] = v_
    # the in-scope part of ibmwallfun (src/modibm.f90:1211-1213, 1240-1242)
    it.call("diffu_corr"); it.call("diffv_corr"); it.call("diffw_corr")
    hc = g["ihc"]
    for n in range(1, nsv + 1):
        sv0n = FArray(g["sv0"].a[:, :, :, n - 1], g["sv0"].lb[:3])
        svpn = FArray(g["svp"].a[:, :, :, n - 1], g["svp"].lb[:3])
        it.call("diffc_corr", sv0n, svpn, hc, hc, hc)
    for k_, v_ in snapshot(w, ["up", "vp", "wp", "svp"]).items():
        out["corr_" + k_] = v_
    it.call("ibmnorm")
    for k_, v_ in snapshot(w, ["um", "vm", "wm", "up", "vp", "wp", "svm", "svp"]).items():
        out["norm_" + k_] = v_
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
    print(f"wrote ref_{tag}.npz  solid pts u/v/w/c {[lists['solid_' + n].shape[0] for n in 'uvwc']}  "
          f"boundary pts {[lists['bound_' + n].shape[0] for n in 'uvwc']}")


def case_forces(tag, shape=(8, 6, 7)):
    """forces (src/modforces.f90:46-133, neutral branch) executed from the reference text."""
    I, J, K = shape
    zf = stretched_zf(K, 0.5 * K * 1.1, 1.07)
    w = World(I, J, K, xlen=0.55 * I, ylen=0.45 * J, zf=zf)
    it = make_interp(w)
    it.load(os.path.join(SRC, "modforces.f90"), only=["forces"])
    g = w.g
    rng = np.random.default_rng(5)
    g["dpdxl"] = FArray(-1e-3 * (1.0 + rng.random(K + 1)), [1])
    g["dpdyl"] = FArray(2e-4 * rng.standard_normal(K + 1), [1])
    g["thlpcar"] = fa([(1, K + 1)], 0.0)
    g["thv0h"] = fa([(0, I + 1), (0, J + 1), (0, K + 1)], 300.0); g["thvh"] = fa([(1, K + 1)], 300.0)
    g["lbuoyancy"] = False
    for nm in ("up", "vp", "wp"):
        g[nm].a[...] = rng.standard_normal(g[nm].a.shape)
    g["thlp"].a[...] = 0.0
    out = {"shape": np.array(shape), "zf": zf, "xlen": 0.55 * I, "ylen": 0.45 * J,
           "dpdxl": np.array(g["dpdxl"].a), "dpdyl": np.array(g["dpdyl"].a)}
    for k_, v_ in snapshot(w, ["up", "vp", "wp"]).items():
        out["in_" + k_] = v_
    it.call("forces")
    for k_, v_ in snapshot(w, ["up", "vp", "wp"]).items():
        out["out_" + k_] = v_
    np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
    print(f"wrote ref_{tag}.npz")


def case_channelglue(tag, shape=(10, 8, 6), nsv=1):
    """bottom -> wfmneutral case 91 (src/modibm.f90:1998-2100, src/modwallfunctions.f90:262-352) and masscorr, volume-flow
    branch (src/modforces.f90:394-420, 470-495), executed from the reference text: the rest of the resident-channel
    glue of examples/999.  avexy_ibm (src/modmpi.f90:623-664) is an MPI helper and is given as a
    single-pencil Python callback like the other MPI calls."""
    I, J, K = shape
    zf = stretched_zf(K, 0.5 * K * 1.1, 1.07)
    w = World(I, J, K, xlen=0.55 * I, ylen=0.45 * J, zf=zf, nsv=nsv, iadv_sv=7)
    g = w.g
    ext = w.externals()

    def avexy_ibm(it_, frame, vals, setters, kw_):
        # aver, var, ib, ie, jb, je, kb, ke, kh, II, IIs, lnan
        var, II, IIs, lnan = vals[1], vals[9], vals[10], vals[11]
        nk = var.a.shape[2]
        averl = np.array([np.sum(var.a[:, :, k] * II.a[:, :, k]) for k in range(nk)])
        IId = np.array(IIs.a, copy=True)
        if (not lnan) and IId[0] == 0:
            averl[0] = np.sum(var.a[:, :, 0])
            IId[0] = IId[nk - 2]            # IId(ke)
        aver = np.where(IId == 0, -999.0, averl / np.where(IId == 0, 1, IId))
        vals[0].a[...] = aver
    ext["avexy_ibm"] = avexy_ibm
    it = Interp(g, ext)
    for f in ("modadvection.f90", "modsubgrid.f90", "modpois.f90", "modtstep.f90", "modboundary.f90", "modchecksim.f90"):
        it.load(os.path.join(SRC, f))
    it.load(os.path.join(SRC, "modibm.f90"), only=["bottom"])
    it.load(os.path.join(SRC, "modwallfunctions.f90"), only=["wfmneutral"])
    it.load(os.path.join(SRC, "modforces.f90"), only=["masscorr"])
    rng = np.random.default_rng(31)
    seed_fields(w, 29, nsv)
    it.call("halos"); it.call("boundary")
    for a, b in (("um", "u0"), ("vm", "v0"), ("wm", "w0")):
        g[a].a[...] = g[b].a + 0.02 * rng.standard_normal(g[b].a.shape)
    for nm in ("up", "vp", "wp", "svp"):
        g[nm].a[...] = rng.standard_normal(g[nm].a.shape)
    g["ekm"].a[...] = 1e-3 * (1.0 + rng.random(g["ekm"].a.shape))
    g["ekh"].a[...] = 3e-3 * (1.0 + rng.random(g["ekh"].a.shape))
    full = [(0, I + 1), (0, J + 1), (0, K + 1)]
    tend = [(0, I + 1), (0, J + 1), (1, K + 1)]
    hc = g["ihc"]
    # IIu / IIv: 1 = fluid (createmasks, src/modibm.f90:2103-2160); a few "solid" points so that the masked means differ from plain ones
    IIu = FArray.alloc([(1 - hc, I + hc), (1 - hc, J + hc), (1, K + hc)], int)
    IIv = FArray.alloc([(1 - hc, I + hc), (1 - hc, J + hc), (1, K + hc)], int)
    IIu.a[...] = 1; IIv.a[...] = 1
    IIu.a[hc + 2:hc + 5, hc + 1:hc + 4, 0:2] = 0
    IIv.a[hc + 2:hc + 4, hc + 1:hc + 5, 0:2] = 0
    IIus = FArray(np.array([int(IIu.a[hc:hc + I, hc:hc + J, k].sum()) for k in range(K + hc)]), [1])
    IIvs = FArray(np.array([int(IIv.a[hc:hc + I, hc:hc + J, k].sum()) for k in range(K + hc)]), [1])
    g.update(iiu=IIu, iiv=IIv, iius=IIus, iivs=IIvs, uflowrate=1.37, vflowrate=-0.21, linoutflow=False,
             luoutflowr=False, lvoutflowr=False, luvolflowr=True, lvvolflowr=True, udef=0.0, vdef=0.0,
             uoutarea=1.0, voutarea=1.0, dxh=fa([(1, I + 1)], g["dx"]), dxhi=fa([(1, I + 1)], 1. / g["dx"]),
             fkar=0.41, lbottom=True, bcbotm=3, bcbott=1, bcbotq=1, bcbots=1, z0=0.01, z0h=0.000067,
             momfluxb=fa(full), tfluxb=fa(full), tau_x=fa(full), tau_y=fa(full), tau_z=fa(full), thl_flux=fa(full),
             wtsurf=0.0, wqsurf=0.0, thls=300.0)
    out = {"zf": zf, "shape": np.array([I, J, K]), "xlen": 0.55 * I, "ylen": 0.45 * J, "nsv": nsv, "z0": 0.01, "fkar": 0.41,
           "uflowrate": 1.37, "vflowrate": -0.21, "IIu": np.array(IIu.a[hc:hc + I, hc:hc + J, :K + 1], dtype=np.int32),
           "IIv": np.array(IIv.a[hc:hc + I, hc:hc + J, :K + 1], dtype=np.int32), "IIus": np.array(IIus.a[:K + 1], dtype=np.int32),
           "IIvs": np.array(IIvs.a[:K + 1], dtype=np.int32)}
    state = ["u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh", "sv0", "svp"]
    for k_, v_ in snapshot(w, state).items():
        out["in_" + k_] = v_
    it.call("bottom")
    for k_, v_ in snapshot(w, ["up", "vp", "wp", "svp", "momfluxb"]).items():
        out["bottom_" + k_] = v_
    for rk in (1, 2, 3):
        g["rk3step"] = rk; g["dt"] = 0.037
        it.call("masscorr")
        for k_, v_ in snapshot(w, ["up", "vp"]).items():
            out[f"mc{rk}_" + k_] = v_
        out[f"mc{rk}_udef"], out[f"mc{rk}_vdef"] = g["udef"], g["vdef"]
    out["dt"] = 0.037
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
    print(f"wrote ref_{tag}.npz  udef {out['mc1_udef']:.6e} {out['mc3_udef']:.6e} vdef {out['mc1_vdef']:.6e}")


def _avexy_ibm_cb(it_, frame, vals, setters, kw_):
    """avexy_ibm (src/modmpi.f90:623-664) on one pencil: an MPI helper, given as a callback like the other MPI calls.
    args: aver, var, ib, ie, jb, je, kb, ke, kh, II, IIs, lnan"""
    var, II, IIs, lnan = vals[1], vals[9], vals[10], vals[11]
    nk = var.a.shape[2]
    averl = np.array([np.sum(var.a[:, :, k] * II.a[:, :, k]) for k in range(nk)])
    IId = np.array(IIs.a, copy=True)
    if (not lnan) and IId[0] == 0:
        averl[0] = np.sum(var.a[:, :, 0])
        IId[0] = IId[nk - 2]            # IId(ke)
    aver = np.where(IId == 0, -999.0, averl / np.where(IId == 0, 1, IId))
    vals[0].a[...] = aver


def case_thermo(tag, shape=(8, 6, 7), BCtopT=1, wttop=-0.01, wtsurf=0.01, with_ibm=False, nsub=3, lbuoycorr=False, wfuno=False):
    """Temperature on the resident path, dry: advecc_2nd + diffc on thl0, bottom (BCbotm = 3 wfmneutral +
    fixed-flux temperature, src/modibm.f90:2033-2046), forces with buoyancy (src/modforces.f90:70-109), tstep_integrate,
    halos, boundary (BCtopT), thermodynamics (src/modthermodynamics.f90:55-121 incl. diagfld, fromztop, calc_halflev, calthv),
    executed from the reference text over `nsub` RK3 substeps in the order of src/program.f90:132-212.  With with_ibm the IBM
    calls on temperature run too: diffc_corr(thl0, thlp) (ibmwallfun, :1225) and ibmnorm's solid(.., mask_c) +
    advecc2nd_corr_liberal (:714-722)."""
    I, J, K = shape
    zf = stretched_zf(K, 0.5 * K * 1.1, 1.07)
    w = World(I, J, K, xlen=0.55 * I, ylen=0.45 * J, zf=zf)
    g = w.g
    ext = w.externals()
    ext["avexy_ibm"] = _avexy_ibm_cb
    it = Interp(g, ext)
    it.gtypes = {"sxfc": complex, "syfc": complex}
    for f in ("modadvection.f90", "modsubgrid.f90", "modpois.f90", "modtstep.f90", "modboundary.f90", "modchecksim.f90"):
        it.load(os.path.join(SRC, f))
    it.load(os.path.join(SRC, "modthermodynamics.f90"), only=["thermodynamics", "diagfld", "fromztop", "calc_halflev", "calthv"])
    it.load(os.path.join(SRC, "modforces.f90"), only=["forces"])
    it.load(os.path.join(SRC, "modibm.f90"), only=["bottom", "solid", "diffu_corr", "diffv_corr", "diffw_corr", "diffc_corr", "ibmnorm",
                                                   "advecc2nd_corr_liberal"])
    it.load(os.path.join(SRC, "modwallfunctions.f90"), only=["wfmneutral", "wfuno", "unom", "unoh"])
    it.call("initpois")
    full = [(0, I + 1), (0, J + 1), (0, K + 1)]
    tend = [(0, I + 1), (0, J + 1), (1, K + 1)]
    prof = [(1, K + 1)]
    thls = 288.1 if wfuno else 288.0      # wfuno: wall temperature inside the range of thl0(kb): stable and unstable points
    g.update(ltempeq=True, lbuoyancy=True, bctopt=BCtopT, wttop=wttop, thl_top=289.5, wtsurf=wtsurf, wqsurf=0.0, thls=thls, thvs=thls,
             qts=0.0, ps=101500.0, pref0=1.e5, rd=287.04, rv=461.5, cp=1004., rlv=2.26e6, chi_half=0.5, khc=1,
             lbottom=True, bcbotm=2 if wfuno else 3, bcbott=2 if wfuno else 1, bcbotq=1, bcbots=1, z0=0.01, z0h=0.000067, fkar=0.41,
             prandtlturb=0.71, bctfluxa=0.0,
             dxh=fa([(1, I + 1)], g["dx"]), dxhi=fa([(1, I + 1)], 1. / g["dx"]),
             momfluxb=fa(full), tfluxb=fa(full), tau_x=fa(full), tau_y=fa(full), tau_z=fa(full), thl_flux=fa(full),
             thl0h=fa(full), qt0h=fa(full), ql0=fa(full), ql0h=fa(full), thv0h=fa(tend), thv0=fa([(1, I), (1, J), (1, K + 1)]),
             th0av=fa(prof), thl0av=fa(prof), qt0av=fa(prof), ql0av=fa(prof), sv0av=fa([(1, K + 1), (1, 1)]),
             thvh=fa(prof), thvf=fa(prof), presf=fa(prof), presh=fa(prof), exnf=fa(prof), exnh=fa(prof), rhof=fa(prof),
             thlpcar=fa(prof), dpdyl=fa(prof), libm=with_ibm, lconservativeibm=False, lwritefac=False, lbuoycorr=lbuoycorr)
    g["dthvdz"] = fa(tend)
    rng = np.random.default_rng(41)
    g["thlpcar"].a[...] = 1e-4 * rng.standard_normal(K + 1)
    g["dpdxl"].a[...] = -1e-3 * (1.0 + rng.random(K + 1))
    g["dpdyl"].a[...] = 2e-4 * rng.standard_normal(K + 1)
    # masks / integer masks (createmasks, src/modibm.f90:2103-2190)
    IIc = FArray.alloc([(1, I), (1, J), (1, K + 1)], int); IIc.a[...] = 1
    IIu = FArray.alloc([(1, I), (1, J), (1, K + 1)], int); IIu.a[...] = 1
    IIv = FArray.alloc([(1, I), (1, J), (1, K + 1)], int); IIv.a[...] = 1
    IIw = FArray.alloc([(1, I), (1, J), (1, K + 1)], int); IIw.a[...] = 1
    lists = None
    if with_ibm:
        lists = ibm_geometry(I, J, K, [(3, 4, 2, 3, 2), (6, 7, 5, 6, 3), (1, 1, 4, 5, 2)])
        for nm in "uvwc":
            pts = lists["solid_" + nm]
            g["solid_info_" + nm] = {"nsolptsrank": int(pts.shape[0]), "solpts_loc": FArray(np.asfortranarray(pts.astype(int)), [1, 1])}
            b = lists["bound_" + nm]
            g["bound_info_" + nm] = {"nbndptsrank": int(b.shape[0]), "bndpts_loc": FArray(np.asfortranarray(b.astype(int)), [1, 1])}
        dummy = fa(tend)
        for nm, II in zip("uvwc", (IIu, IIv, IIw, IIc)):
            m = fa(full, 1.0)
            m.a[:, :, 0] = 0.0
            if nm == "w":
                m.a[:, :, 1] = 0.0
            it.call("solid", g["solid_info_" + nm], m, dummy, 0.0, 1, 1, 1)
            a = m.a
            a[0] = a[I]; a[I + 1] = a[1]; a[:, 0] = a[:, J]; a[:, J + 1] = a[:, 1]
            g["mask_" + nm] = m
            for q in lists["solid_" + nm]:
                II.a[q[0] - 1, q[1] - 1, q[2] - 1] = 0
        IIw.a[:, :, 0] = 0
    for nm, II in (("c", IIc), ("u", IIu), ("v", IIv), ("w", IIw)):
        g["ii" + nm] = II
        g["ii" + nm + "s"] = FArray(np.array([int(II.a[:, :, k].sum()) for k in range(K + 1)]), [1])
    seed_fields(w, 37, 0)
    g["thl0"].a[...] = 0.0
    g["thl0"].a[1:-1, 1:-1, 1:-1] = thls + 0.05 * np.arange(1, K + 1)[None, None, :] + 0.3 * rng.standard_normal((I, J, K))
    g["thl0"].a[:, :, 0] = g["thl0"].a[:, :, 1]                 # startup: thl0(kb-1) = thl0(kb), src/modstartup.f90:1208
    g["ekm"].a[...] = g["numol"]
    g["ekh"].a[...] = g["numol"] * g["prandtlmoli"]
    it.call("halos"); it.call("boundary")
    for a, b in (("um", "u0"), ("vm", "v0"), ("wm", "w0"), ("thlm", "thl0")):
        g[a].a[...] = g[b].a
    it.call("thermodynamics")
    out = {"zf": zf, "shape": np.array([I, J, K]), "xlen": 0.55 * I, "ylen": 0.45 * J, "BCtopT": BCtopT, "wttop": wttop, "thl_top": 289.5,
           "wtsurf": wtsurf, "thls": thls, "grav": g["grav"], "z0": 0.01, "fkar": 0.41, "with_ibm": int(with_ibm),
           "lbuoycorr": int(lbuoycorr), "Rigc": g["rigc"], "wfuno": int(wfuno), "z0h": 0.000067, "prandtlturb": 0.71,
           "thlpcar": np.array(g["thlpcar"].a), "dpdxl": np.array(g["dpdxl"].a), "dpdyl": np.array(g["dpdyl"].a)}
    if lists:
        for k_, v_ in lists.items():
            out["pts_" + k_] = v_
    state = ["u0", "v0", "w0", "um", "vm", "wm", "pres0", "thl0", "thlm"]
    diag = ["thl0h", "thv0h", "dthvdz", "thvh", "thl0av", "th0av"]
    for k_, v_ in snapshot(w, state + diag).items():
        out["in_" + k_] = v_
    dt = 0.03
    g["dt"] = dt; g["dtmax"] = dt; g["ladaptive"] = False
    for s in range(nsub):
        it.call("tstep_update")
        it.call("advection")
        if s == 0:
            out["adv_thlp"] = np.array(g["thlp"].a, copy=True)
        it.call("subgrid")
        if s == 0:
            out["sub_thlp"] = np.array(g["thlp"].a, copy=True)
            out["sub_ekm"] = np.array(g["ekm"].a, copy=True); out["sub_ekh"] = np.array(g["ekh"].a, copy=True)
        it.call("bottom")
        if s == 0:
            for k_, v_ in snapshot(w, ["up", "vp", "thlp"]).items():
                out["bottom_" + k_] = v_
        it.call("forces")
        if s == 0:
            for k_, v_ in snapshot(w, ["up", "vp", "wp", "thlp"]).items():
                out["forces_" + k_] = v_
        if with_ibm:
            it.call("diffu_corr"); it.call("diffv_corr"); it.call("diffw_corr")
            it.call("diffc_corr", g["thl0"], g["thlp"], 1, 1, 1)
            if s == 0:
                out["corr_thlp"] = np.array(g["thlp"].a, copy=True)
            it.call("ibmnorm")
            if s == 0:
                for k_, v_ in snapshot(w, ["thlp", "thlm", "wp", "wm"]).items():
                    out["norm_" + k_] = v_
        it.call("poisson")
        it.call("tstep_integrate")
        it.call("halos")
        it.call("boundary")
        it.call("thermodynamics")
        for k_, v_ in snapshot(w, state + diag).items():
            out[f"s{s + 1}_" + k_] = v_
    for k_, v_ in out.items():
        if isinstance(v_, np.ndarray) and v_.dtype.kind == "f" and v_.ndim == 3 and not k_.startswith("in_"):
            assert np.isfinite(v_[1:-1, 1:-1, 1:-1]).all(), k_
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"ref_{tag}.npz"), **out)
    print(f"wrote ref_{tag}.npz ({len(out)} arrays) thvh {np.array(g['thvh'].a)[:3]} |w0|max {np.abs(g['w0'].a).max():.4f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "thermo":
        case_thermo("thermo_flux")
        case_thermo("thermo_value_ibm", shape=(8, 8, 6), BCtopT=2, wttop=0.0, wtsurf=-0.005, with_ibm=True)
        case_thermo("thermo_buoycorr", shape=(6, 8, 8), lbuoycorr=True)
        case_thermo("thermo_wfuno", shape=(8, 6, 6), wfuno=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "forces":
        case_forces("forces")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "channelglue":
        case_channelglue("channelglue")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ibm":
        case_ibm("ibm")
        sys.exit(0)
    case_substeps("vreman_freeslip", shape=(10, 8, 6))
    case_substeps("smag_noslip", shape=(8, 12, 5), lvreman=False, lsmagorinsky=True, BCtopm=2, Uinf=1.2, Vinf=-0.3)
    case_substeps("dns", shape=(8, 6, 7), lvreman=False, lsmagorinsky=False)
    case_substeps("kappa2", shape=(8, 8, 6), nsv=2, iadv_sv=7)
    case_substeps("cd2scalar", shape=(6, 8, 5), nsv=1, iadv_sv=2)
    case_ibm("ibm")
    case_forces("forces")
    case_channelglue("channelglue")
    case_thermo("thermo_flux")
    case_thermo("thermo_value_ibm", shape=(8, 8, 6), BCtopT=2, wttop=0.0, wtsurf=-0.005, with_ibm=True)
    case_thermo("thermo_flux")
    case_thermo("thermo_value_ibm", shape=(8, 8, 6), BCtopT=2, wttop=0.0, wtsurf=-0.005, with_ibm=True)
    case_thermo("thermo_buoycorr", shape=(6, 8, 8), lbuoycorr=True)
    case_thermo("thermo_wfuno", shape=(8, 6, 6), wfuno=True)
