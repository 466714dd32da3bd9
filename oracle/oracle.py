"""ctypes front-end of the CPU parity oracle (oracle/udales_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  Never imported by the product package.

Arrays are exposed as numpy views in Fortran order with the reference's shapes
(src/modfields.f90:440-474): ``o.u0[i + ih - 1, j + jh - 1, k + kh - 1]`` is Fortran ``u0(i,j,k)``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LIB_PATH = None   # None = the parity build (liboracle.so, -ffp-contract=off); bench.py's CPU arm points this at its own
                   # -O3 -march=native build before the first Oracle is created (timing only, never parity)


def use_library(path):
    global _LIB, _LIB_PATH
    _LIB, _LIB_PATH = None, path


class _Cfg(C.Structure):
    _fields_ = [("itot", C.c_int), ("jtot", C.c_int), ("ktot", C.c_int), ("nsv", C.c_int),
                ("BCtopm", C.c_int), ("lles", C.c_int), ("lvreman", C.c_int), ("lsmagorinsky", C.c_int),
                ("iadv_sv", C.c_int),
                ("xlen", C.c_double), ("ylen", C.c_double),
                ("numol", C.c_double), ("prandtlmoli", C.c_double), ("prandtli", C.c_double),
                ("c_vreman", C.c_double), ("cs", C.c_double), ("Uinf", C.c_double), ("Vinf", C.c_double),
                ("zf", C.POINTER(C.c_double))]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "udales_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_LIB_PATH or build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Cfg)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_field.restype = C.POINTER(C.c_double)
        L.orc_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.orc_metric.restype = C.POINTER(C.c_double)
        L.orc_metric.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for f in ("orc_advection", "orc_closure", "orc_subgrid", "orc_tderive", "orc_halos", "orc_boundary"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        for f in ("orc_fillps", "orc_poisson", "orc_tstep_integrate"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double, C.c_int]
            getattr(L, f).restype = None
        L.orc_poisson_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_tstep_update.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_chkdiv.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3
        L.orc_randomize.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_double, C.c_int]
        L.orc_substep.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_double, C.c_int,
                                  C.c_double, C.c_double]
        L.orc_rfft_packed.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
        L.orc_set_forcing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_forces.argtypes = [C.c_void_p]
        L.orc_set_bottom.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_bottom.argtypes = [C.c_void_p]
        L.orc_set_wfuno.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.orc_momfluxb.restype = C.POINTER(C.c_double)
        L.orc_momfluxb.argtypes = [C.c_void_p]
        L.orc_set_masscorr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_masscorr.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.orc_masscorr_get.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_ibm_set_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_ibm_build_masks.argtypes = [C.c_void_p]
        L.orc_ibm_mask.restype = C.POINTER(C.c_double)
        L.orc_ibm_mask.argtypes = [C.c_void_p, C.c_int]
        L.orc_ibmnorm.argtypes = [C.c_void_p]
        L.orc_ibm_diffcorr.argtypes = [C.c_void_p]
        L.orc_set_thermo.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                     C.c_int, C.c_double, C.c_void_p]
        L.orc_thermodynamics.argtypes = [C.c_void_p]
        L.orc_set_buoycorr.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_thermo_profile.restype = C.POINTER(C.c_double)
        L.orc_thermo_profile.argtypes = [C.c_void_p, C.c_char_p]
        _LIB = L
    return _LIB


def equidistant_zf(ktot: int, zsize: float) -> np.ndarray:
    dz = zsize / ktot
    return (np.arange(ktot) + 0.5) * dz


def stretched_zf(ktot: int, zsize: float, ratio: float = 1.03) -> np.ndarray:
    """geometric stretching: dzf(k+1) = ratio*dzf(k); exercises every dzf/dzh weight."""
    dz = ratio ** np.arange(ktot)
    dz *= zsize / dz.sum()
    zh = np.concatenate([[0.0], np.cumsum(dz)])
    return 0.5 * (zh[:-1] + zh[1:])


FIELDS = ("u0", "v0", "w0", "um", "vm", "wm", "pres0", "p", "ekm", "ekh",
          "up", "vp", "wp", "pup", "pvp", "pwp", "rhs", "sv0", "svm", "svp")
THERMO_FIELDS = ("thl0", "thlm", "thlp", "thl0h", "thv0h", "dthvdz")


class Oracle:
    """One single-pencil uDALES dynamics state on the CPU."""

    def __init__(self, itot, jtot, ktot, xlen=None, ylen=None, zf=None, nsv=0, BCtopm=1,
                 lvreman=True, lsmagorinsky=False, lles=None, iadv_sv=7,
                 numol=1.5e-5, prandtlmol=0.71, prandtl=0.333, c_vreman=0.07, cs=-1.0,
                 Uinf=0.0, Vinf=0.0):
        self.L = lib()
        xlen = float(xlen if xlen is not None else itot / 2.0)
        ylen = float(ylen if ylen is not None else jtot / 2.0)
        if zf is None:
            zf = equidistant_zf(ktot, ktot * xlen / itot)
        self.zf = np.ascontiguousarray(zf, dtype=np.float64)
        assert self.zf.size == ktot
        if lles is None:
            lles = bool(lvreman or lsmagorinsky)
        self.cfg = _Cfg(itot, jtot, ktot, nsv, BCtopm, int(lles), int(lvreman), int(lsmagorinsky), iadv_sv,
                        xlen, ylen, numol, 1.0 / prandtlmol, 1.0 / prandtl, c_vreman, cs, Uinf, Vinf,
                        self.zf.ctypes.data_as(C.POINTER(C.c_double)))
        self.h = C.c_void_p(self.L.orc_create(C.byref(self.cfg)))
        self.itot, self.jtot, self.ktot, self.nsv = itot, jtot, ktot, nsv
        self.ih = self.jh = self.kh = 1
        self.ihc = self.jhc = self.khc = 2 if (nsv > 0 and iadv_sv == 7) else 1
        self.dx, self.dy = xlen / itot, ylen / jtot
        self._map_fields(FIELDS)
        self.ltempeq = False
        self.rk3step = 0
        self.dt = 0.0

    def _map_fields(self, names):
        for name in names:
            dims = (C.c_int * 4)()
            ptr = self.L.orc_field(self.h, name.encode(), dims)
            if not ptr:
                setattr(self, name, None)
                continue
            shape = tuple(dims[:3]) + ((dims[3],) if name.startswith("sv") else ())
            n = int(np.prod(shape))
            arr = np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape, order="F")
            setattr(self, name, arr)

    def metric(self, name):
        lo, n = C.c_int(), C.c_int()
        ptr = self.L.orc_metric(self.h, name.encode(), C.byref(lo), C.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(n.value,)), lo.value

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # reference call surface -------------------------------------------------
    def advection(self): self.L.orc_advection(self.h)
    def closure(self): self.L.orc_closure(self.h)
    def subgrid(self): self.L.orc_subgrid(self.h)
    def fillps(self, dt, rk3step): self.L.orc_fillps(self.h, dt, rk3step)
    def tderive(self): self.L.orc_tderive(self.h)
    def poisson(self, dt, rk3step): self.L.orc_poisson(self.h, dt, rk3step)
    def tstep_integrate(self, dt, rk3step): self.L.orc_tstep_integrate(self.h, dt, rk3step)
    def halos(self): self.L.orc_halos(self.h)
    def boundary(self): self.L.orc_boundary(self.h)

    def poisson_solve(self, rhs: np.ndarray) -> np.ndarray:
        pz = np.array(rhs, dtype=np.float64, order="F", copy=True)
        assert pz.shape == (self.itot, self.jtot, self.ktot)
        self.L.orc_poisson_solve(self.h, pz.ctypes.data_as(C.POINTER(C.c_double)))
        return pz

    def tstep_update(self, dt, rk3step, courant=1.0, diffnr=0.25, dtmax=1e9, ladaptive=True):
        d, r = C.c_double(dt), C.c_int(rk3step)
        ct, dn = C.c_double(0), C.c_double(0)
        self.L.orc_tstep_update(self.h, C.byref(d), courant, diffnr, dtmax, int(ladaptive), C.byref(r),
                                C.byref(ct), C.byref(dn))
        return d.value, r.value, ct.value, dn.value

    def chkdiv(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_chkdiv(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def randomize(self, name, ampl, ir=43, n=0):
        self.L.orc_randomize(self.h, name.encode(), n, ampl, ir)

    def substep(self, dtmax, ladaptive=False, courant=1.0, diffnr=0.25):
        d, r = C.c_double(self.dt), C.c_int(self.rk3step)
        self.L.orc_substep(self.h, C.byref(d), C.byref(r), dtmax, int(ladaptive), courant, diffnr)
        self.dt, self.rk3step = d.value, r.value

    # forces (src/modforces.f90:46) -------------------------------------------------
    def set_forcing(self, dpdxl, dpdyl):
        a = np.ascontiguousarray(dpdxl, dtype=np.float64); b = np.ascontiguousarray(dpdyl, dtype=np.float64)
        assert a.size == self.ktot + 1 and b.size == self.ktot + 1
        self.L.orc_set_forcing(self.h, a.ctypes.data, b.ctypes.data)

    def forces(self): self.L.orc_forces(self.h)

    # temperature, dry -----------------------------------------------
    def set_thermo(self, lbuoyancy=True, grav=9.81, thls=288.0, BCtopT=1, wttop=0.0, thl_top=288.0, BCbotT=1, wtsurf=0.0,
                   thlpcar=None):
        a = None if thlpcar is None else np.ascontiguousarray(thlpcar, dtype=np.float64)
        assert a is None or a.size == self.ktot + 1
        self.L.orc_set_thermo(self.h, int(lbuoyancy), grav, thls, BCtopT, wttop, thl_top, BCbotT, wtsurf,
                              None if a is None else a.ctypes.data)
        self._map_fields(THERMO_FIELDS)
        self.ltempeq = True

    def thermodynamics(self): self.L.orc_thermodynamics(self.h)
    def set_buoycorr(self, lbuoycorr=True, Rigc=0.25): self.L.orc_set_buoycorr(self.h, int(lbuoycorr), Rigc)

    def thermo_profile(self, name):
        return np.ctypeslib.as_array(self.L.orc_thermo_profile(self.h, name.encode()), shape=(self.ktot + 1,))

    # bottom -> wfmneutral (src/modibm.f90:1998, src/modwallfunctions.f90:307) and masscorr (src/modforces.f90:328) ----
    def set_bottom(self, z0, fkar=0.41, lbottom=True, BCbotm=3, BCbots=1):
        self.L.orc_set_bottom(self.h, int(lbottom), BCbotm, BCbots, z0, fkar)

    def set_wfuno(self, z0h=0.00035, prandtlturb=0.71, grav=9.81, thls=288.0, tcell=288.0):
        """parameters of the wall functions with stability correction (BCbotm = 2, BCbotT = 2)"""
        self.L.orc_set_wfuno(self.h, z0h, prandtlturb, grav, thls, tcell)

    def bottom(self): self.L.orc_bottom(self.h)

    def momfluxb(self):
        shape = (self.itot + 2, self.jtot + 2, self.ktot + 2)
        return np.ctypeslib.as_array(self.L.orc_momfluxb(self.h), shape=(int(np.prod(shape)),)).reshape(shape, order="F")

    def set_masscorr(self, uflowrate=None, vflowrate=None, IIu=None, IIv=None):
        """volume-flow forcing (luvolflowr / lvvolflowr); IIu, IIv: (itot, jtot, ktot+1) int arrays, 1 = fluid, or None"""
        a = None if IIu is None else np.asfortranarray(IIu, dtype=np.int32)
        b = None if IIv is None else np.asfortranarray(IIv, dtype=np.int32)
        self.L.orc_set_masscorr(self.h, int(uflowrate is not None), int(vflowrate is not None), float(uflowrate or 0.0), float(vflowrate or 0.0),
                                None if a is None else a.ctypes.data, None if b is None else b.ctypes.data)

    def masscorr(self, dt, rk3step):
        self.L.orc_masscorr(self.h, dt, rk3step)
        u, v = C.c_double(), C.c_double()
        self.L.orc_masscorr_get(self.h, C.byref(u), C.byref(v))
        return u.value, v.value

    # immersed boundary masking (src/modibm.f90) ---------------------------------
    IBM_KINDS = ("solid_u", "solid_v", "solid_w", "solid_c", "bound_u", "bound_v", "bound_w", "bound_c")

    def ibm_set(self, lists):
        """lists: dict kind -> (n,3) int array of local 1-based (i,j,k); missing kinds = empty.  Builds the masks."""
        for kind, name in enumerate(self.IBM_KINDS):
            pts = np.ascontiguousarray(np.asarray(lists.get(name, np.zeros((0, 3))), dtype=np.int32).reshape(-1, 3))
            self.L.orc_ibm_set_points(self.h, kind, pts.shape[0], pts.ctypes.data)
        self.L.orc_ibm_build_masks(self.h)

    def ibm_mask(self, m):
        shape = (self.itot + 2, self.jtot + 2, self.ktot + 2)
        return np.ctypeslib.as_array(self.L.orc_ibm_mask(self.h, m), shape=(int(np.prod(shape)),)).reshape(shape, order="F")

    def ibmnorm(self): self.L.orc_ibmnorm(self.h)
    def ibm_diffcorr(self): self.L.orc_ibm_diffcorr(self.h)

    # synthetic channel of SURVEY.md §8d --------------------------------------
    def init_channel(self, ubase=1.0, ampl=0.05, ir=43):
        """u = ubase + LCG noise, v, w = LCG noise; ghosts as A.5; then one projection-free ghost fill."""
        for f in (self.u0, self.v0, self.w0, self.pres0):
            f[...] = 0.0
        self.u0[1:-1, 1:-1, 1:-1] = ubase
        self.randomize("u0", ampl, ir)
        self.randomize("v0", ampl, ir + 1)
        self.randomize("w0", ampl, ir + 2)
        self.w0[:, :, 1] = 0.0
        for n in range(self.nsv):
            hc = self.ihc
            self.sv0[..., n] = 0.0
            self.sv0[hc:-hc, hc:-hc, hc:-hc, n] = 1.0
            self.randomize("sv0", 0.1, ir + 10 + n, n)
        self.halos()
        self.boundary()
        self.um[...] = self.u0
        self.vm[...] = self.v0
        self.wm[...] = self.w0
        if self.nsv:
            self.svm[...] = self.sv0
        for f in (self.up, self.vp, self.wp):
            f[...] = 0.0


def rfft_packed(line: np.ndarray, inverse: bool = False) -> np.ndarray:
    x = np.array(line, dtype=np.float64, copy=True)
    lib().orc_rfft_packed(x.size, x.ctypes.data_as(C.POINTER(C.c_double)), int(inverse))
    return x
