"""udales_b200 — host-side mirror of the uDALES dynamics call surface over the sm_100a C-ABI.

The product is ``libudales_gpu.so`` (include/udales_gpu.h).  This module is the thin ctypes
binding used by the tests, bench.py and Python drivers; its method names are the reference's
procedure names (src/program.f90:134-207): ``tstep_update, advection, subgrid, poisson,
tstep_integrate, halos, boundary``.  There is no CPU fallback: if the library is missing or no
B200 is visible, construction raises.

The directory name contains a hyphen (``u-dales_b200``), so the importable alias is the
top-level ``udales_b200.py`` shim.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libudales_gpu.so")
ABI_VERSION = 2

FIELD_IDS = {"u0": 0, "v0": 1, "w0": 2, "um": 3, "vm": 4, "wm": 5, "up": 6, "vp": 7, "wp": 8,
             "pres0": 9, "p": 10, "ekm": 11, "ekh": 12, "rhs": 13, "sv0": 14, "svm": 15, "svp": 16, "momfluxb": 17,
             "thl0": 18, "thlm": 19, "thlp": 20}

EXPORTS = [
    "udgpu_nccl_unique_id", "udgpu_init", "udgpu_finalize", "udgpu_last_error", "udgpu_abi_version",
    "udgpu_push", "udgpu_pull", "udgpu_pull_points", "udgpu_add_points", "udgpu_field_count", "udgpu_device_ptr", "udgpu_sync",
    "udgpu_host_register", "udgpu_host_unregister",
    "udgpu_tstep_update", "udgpu_advection", "udgpu_subgrid", "udgpu_closure", "udgpu_poisson",
    "udgpu_poisson_solve", "udgpu_poisson_solve_resident", "udgpu_fillps", "udgpu_tderive",
    "udgpu_tstep_integrate", "udgpu_halos", "udgpu_boundary", "udgpu_divergence", "udgpu_substep",
    "udgpu_rk3_step_host", "udgpu_set_forcing", "udgpu_forces", "udgpu_set_bottom", "udgpu_set_wfuno", "udgpu_bottom", "udgpu_set_masscorr", "udgpu_masscorr", "udgpu_ibm_set_points", "udgpu_ibm_commit", "udgpu_ibm_pull_mask", "udgpu_ibmnorm", "udgpu_ibm_diffcorr",
    "udgpu_set_thermo", "udgpu_thermodynamics", "udgpu_thermo_profile", "udgpu_set_buoycorr",
    "udgpu_profile_enable", "udgpu_profile_get", "udgpu_profile_reset", "udgpu_launch_count", "udgpu_stream", "udgpu_trace_dump",
]


class Cfg(C.Structure):
    """ctypes image of ``udgpu_cfg`` (include/udales_gpu.h)."""
    _fields_ = [("abi_version", C.c_int),
                ("itot", C.c_int), ("jtot", C.c_int), ("ktot", C.c_int),
                ("imax", C.c_int), ("jmax", C.c_int), ("kmax", C.c_int),
                ("ih", C.c_int), ("jh", C.c_int), ("kh", C.c_int),
                ("ihc", C.c_int), ("jhc", C.c_int), ("khc", C.c_int),
                ("nsv", C.c_int), ("zstart", C.c_int * 3),
                ("nprocx", C.c_int), ("nprocy", C.c_int), ("myidx", C.c_int), ("myidy", C.c_int),
                ("BCxm", C.c_int), ("BCym", C.c_int), ("BCtopm", C.c_int), ("BCzp", C.c_int),
                ("ipoiss", C.c_int), ("iadv_mom", C.c_int), ("iadv_sv", C.c_int),
                ("lles", C.c_int), ("lvreman", C.c_int), ("lsmagorinsky", C.c_int), ("loneeqn", C.c_int),
                ("ltempeq", C.c_int), ("lmoist", C.c_int),
                ("dx", C.c_double), ("dy", C.c_double),
                ("dzf", C.POINTER(C.c_double)), ("dzh", C.POINTER(C.c_double)), ("delta", C.POINTER(C.c_double)),
                ("numol", C.c_double), ("prandtlmoli", C.c_double), ("prandtli", C.c_double),
                ("c_vreman", C.c_double), ("cs", C.c_double), ("Uinf", C.c_double), ("Vinf", C.c_double),
                ("e12min", C.c_double), ("device", C.c_int), ("flags", C.c_int), ("iadv_thl", C.c_int)]


class UdalesGPUError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libudales_gpu.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", _HERE, "libudales_gpu.so"], capture_output=True, text=True)
    if out.returncode != 0:
        raise UdalesGPUError("nvcc build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


_LIB = None


def lib():
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise UdalesGPUError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.udgpu_last_error.restype = C.c_char_p
        L.udgpu_launch_count.restype = C.c_long
        L.udgpu_launch_count.argtypes = [C.c_void_p]
        L.udgpu_init.argtypes = [C.POINTER(Cfg), C.c_void_p, C.POINTER(C.c_void_p)]
        L.udgpu_finalize.argtypes = [C.c_void_p]
        L.udgpu_push.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.udgpu_pull.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.udgpu_field_count.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.udgpu_device_ptr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.udgpu_sync.argtypes = [C.c_void_p]
        L.udgpu_pull_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
        L.udgpu_add_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
        L.udgpu_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.udgpu_host_unregister.argtypes = [C.c_void_p]
        L.udgpu_tstep_update.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in ("udgpu_advection", "udgpu_subgrid", "udgpu_closure", "udgpu_poisson_solve_resident",
                  "udgpu_tderive", "udgpu_halos", "udgpu_boundary"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("udgpu_poisson", "udgpu_fillps", "udgpu_tstep_integrate"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.udgpu_poisson_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.udgpu_divergence.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3
        L.udgpu_substep.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_double, C.c_int,
                                    C.c_double, C.c_double]
        L.udgpu_rk3_step_host.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.POINTER(C.c_double), C.c_double, C.c_int,
                                                                              C.c_double, C.c_double]
        L.udgpu_set_forcing.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.udgpu_forces.argtypes = [C.c_void_p]
        L.udgpu_set_bottom.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.udgpu_bottom.argtypes = [C.c_void_p]
        L.udgpu_set_wfuno.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.udgpu_set_masscorr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.udgpu_masscorr.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.udgpu_ibm_set_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.udgpu_ibm_commit.argtypes = [C.c_void_p]
        L.udgpu_ibm_pull_mask.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.udgpu_ibmnorm.argtypes = [C.c_void_p]
        L.udgpu_ibm_diffcorr.argtypes = [C.c_void_p]
        L.udgpu_set_thermo.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                       C.c_void_p]
        L.udgpu_thermodynamics.argtypes = [C.c_void_p]
        L.udgpu_set_buoycorr.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.udgpu_thermo_profile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.udgpu_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.udgpu_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_long)]
        L.udgpu_profile_reset.argtypes = [C.c_void_p]
        L.udgpu_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.udgpu_nccl_unique_id.argtypes = [C.c_void_p]
        L.udgpu_trace_dump.argtypes = [C.c_void_p, C.c_char_p]
        _LIB = L
    return _LIB


def grid_metrics(zf: np.ndarray):
    """dzf(kb-kh:ke+kh), dzh(kb:ke+kh) from cell-centre heights, src/modglobal.f90:746-760."""
    K = zf.size
    zfx = np.zeros(K + 2)          # 1-based: zfx[k] = zf(k)
    zfx[1:K + 1] = zf
    zh1 = np.zeros(K + 3)
    zh1[1] = 0.0
    for k in range(1, K + 1):
        zh1[k + 1] = zh1[k] + 2.0 * (zfx[k] - zh1[k])
    zfx[K + 1] = zfx[K] + 2.0 * (zh1[K + 1] - zfx[K])
    dzf = np.zeros(K + 2)          # index k = 0..K+1
    for k in range(1, K + 1):
        dzf[k] = zh1[k + 1] - zh1[k]
    dzf[K + 1] = dzf[K]
    dzf[0] = dzf[1]
    dzh = np.zeros(K + 1)          # index k = 1..K+1 -> dzh[k-1]
    dzh[0] = 2 * zfx[1]
    for k in range(2, K + 2):
        dzh[k - 1] = zfx[k] - zfx[k - 1]
    return dzf, dzh


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 calls this, the host broadcasts it: MPI_Bcast in Fortran,
    torch.distributed in the tests/bench)."""
    buf = C.create_string_buffer(128)
    rc = lib().udgpu_nccl_unique_id(buf)
    if rc != 0:
        raise UdalesGPUError(f"udgpu error {rc}: {lib().udgpu_last_error().decode()}")
    return buf.raw


def slab_of(a_global: np.ndarray, nprocx: int, myidx: int, halo: int = 1) -> np.ndarray:
    """x-slab (with its halo columns) of a global halo'd Fortran-shaped array: local storage column c is
    global column myidx*imax + c (decomp_2d zstart arithmetic, 2decomp-fft/src/decomp_2d.f90:1172-1204)."""
    itot = a_global.shape[0] - 2 * halo
    imax = itot // nprocx
    lo = myidx * imax
    return np.asfortranarray(a_global[lo:lo + imax + 2 * halo])


class UdalesGPU:
    """One z-pencil of the uDALES dynamics core resident on one B200."""

    def __init__(self, itot, jtot, ktot, xlen=None, ylen=None, zf=None, nsv=0, BCtopm=1,
                 lvreman=True, lsmagorinsky=False, lles=None, iadv_sv=7,
                 numol=1.5e-5, prandtlmol=0.71, prandtl=0.333, c_vreman=0.07, cs=-1.0,
                 Uinf=0.0, Vinf=0.0, device=-1, flags=0, nprocx=1, myidx=0, nccl_uid=None, ltempeq=False, iadv_thl=2):
        self.L = lib()
        xlen = float(xlen if xlen is not None else itot / 2.0)
        ylen = float(ylen if ylen is not None else jtot / 2.0)
        if zf is None:
            dz = xlen / itot
            zf = (np.arange(ktot) + 0.5) * dz
        zf = np.ascontiguousarray(zf, dtype=np.float64)
        self._dzf, self._dzh = grid_metrics(zf)
        if lles is None:
            lles = bool(lvreman or lsmagorinsky)
        hc = 2 if (nsv > 0 and iadv_sv == 7) else 1
        c = Cfg()
        c.abi_version = ABI_VERSION
        c.itot, c.jtot, c.ktot = itot, jtot, ktot
        c.imax, c.jmax, c.kmax = itot // nprocx, jtot, ktot
        c.ih = c.jh = c.kh = 1
        c.ihc = c.jhc = c.khc = hc
        c.nsv = nsv
        c.zstart[0] = myidx * (itot // nprocx) + 1
        c.zstart[1] = c.zstart[2] = 1
        c.nprocx, c.nprocy = nprocx, 1
        c.myidx, c.myidy = myidx, 0
        c.BCxm = c.BCym = 1
        c.BCtopm = BCtopm
        c.BCzp = 1
        c.ipoiss = 0
        c.iadv_mom = 2
        c.iadv_sv = iadv_sv
        c.lles, c.lvreman, c.lsmagorinsky, c.loneeqn = int(lles), int(lvreman), int(lsmagorinsky), 0
        c.ltempeq, c.lmoist, c.iadv_thl = int(ltempeq), 0, iadv_thl
        c.dx, c.dy = xlen / itot, ylen / jtot
        c.dzf = self._dzf.ctypes.data_as(C.POINTER(C.c_double))
        c.dzh = self._dzh.ctypes.data_as(C.POINTER(C.c_double))
        c.delta = None
        c.numol, c.prandtlmoli, c.prandtli = numol, 1.0 / prandtlmol, 1.0 / prandtl
        c.c_vreman, c.cs, c.Uinf, c.Vinf, c.e12min = c_vreman, cs, Uinf, Vinf, 5e-5
        c.device, c.flags = device, flags
        self.cfg = c
        self.h = C.c_void_p()
        uid = C.create_string_buffer(nccl_uid, 128) if nccl_uid is not None else None
        self._chk(self.L.udgpu_init(C.byref(c), uid, C.byref(self.h)))
        self.itot, self.jtot, self.ktot, self.nsv = itot, jtot, ktot, nsv
        self.imax, self.nprocx, self.myidx = itot // nprocx, nprocx, myidx
        self.dt, self.rk3step = 0.0, 0

    def _chk(self, rc):
        if rc != 0:
            raise UdalesGPUError(f"udgpu error {rc}: {self.L.udgpu_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.udgpu_finalize(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # residency ---------------------------------------------------------------
    def shape(self, name):
        cnt, dims = C.c_size_t(), (C.c_int * 3)()
        self._chk(self.L.udgpu_field_count(self.h, FIELD_IDS[name], C.byref(cnt), dims))
        return tuple(dims)

    def push(self, name, arr, n4=0):
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.shape(name), (name, a.shape, self.shape(name))
        self._chk(self.L.udgpu_push(self.h, FIELD_IDS[name], n4, a.ctypes.data))
        self._chk(self.L.udgpu_sync(self.h))

    def pull(self, name, n4=0, out=None):
        if out is None:
            out = np.empty(self.shape(name), dtype=np.float64, order="F")
        self._chk(self.L.udgpu_pull(self.h, FIELD_IDS[name], n4, out.ctypes.data))
        return out

    def push_raw(self, name, ptr, n4=0):
        self._chk(self.L.udgpu_push(self.h, FIELD_IDS[name], n4, ptr))

    def pull_raw(self, name, ptr, n4=0):
        self._chk(self.L.udgpu_pull(self.h, FIELD_IDS[name], n4, ptr))

    def _offsets(self, name, ijk):
        """(n,3) 0-based storage indices (halo cells included) -> linear Fortran offsets into the field"""
        d = self.shape(name)
        ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
        return np.ascontiguousarray(ijk[:, 0] + d[0] * (ijk[:, 1] + d[1] * ijk[:, 2]))

    def pull_points(self, name, ijk, n4=0):
        """values of a field at a list of points (sparse residency for host add-ons such as the facet wall functions)"""
        off = self._offsets(name, ijk)
        out = np.empty(off.size)
        self._chk(self.L.udgpu_pull_points(self.h, FIELD_IDS[name], n4, off.size, off.ctypes.data, out.ctypes.data))
        return out

    def add_points(self, name, ijk, vals, n4=0):
        """tendency(name) += vals at a list of points"""
        off = self._offsets(name, ijk)
        v = np.ascontiguousarray(vals, dtype=np.float64)
        assert v.size == off.size
        self._chk(self.L.udgpu_add_points(self.h, FIELD_IDS[name], n4, off.size, off.ctypes.data, v.ctypes.data))

    def sync(self): self._chk(self.L.udgpu_sync(self.h))

    # reference call surface ----------------------------------------------------
    def tstep_update(self, dt, rk3step, courant=1.0, diffnr=0.25, dtmax=1e9, ladaptive=True):
        d, r, ct, dn = C.c_double(dt), C.c_int(rk3step), C.c_double(0), C.c_double(0)
        self._chk(self.L.udgpu_tstep_update(self.h, C.byref(d), courant, diffnr, dtmax, int(ladaptive),
                                            C.byref(r), C.byref(ct), C.byref(dn)))
        return d.value, r.value, ct.value, dn.value

    def advection(self): self._chk(self.L.udgpu_advection(self.h))
    def subgrid(self): self._chk(self.L.udgpu_subgrid(self.h))
    def closure(self): self._chk(self.L.udgpu_closure(self.h))
    def fillps(self, dt, rk3step): self._chk(self.L.udgpu_fillps(self.h, dt, rk3step))
    def tderive(self): self._chk(self.L.udgpu_tderive(self.h))
    def poisson(self, dt, rk3step): self._chk(self.L.udgpu_poisson(self.h, dt, rk3step))
    def tstep_integrate(self, dt, rk3step): self._chk(self.L.udgpu_tstep_integrate(self.h, dt, rk3step))
    def halos(self): self._chk(self.L.udgpu_halos(self.h))
    def boundary(self): self._chk(self.L.udgpu_boundary(self.h))
    def poisson_solve_resident(self): self._chk(self.L.udgpu_poisson_solve_resident(self.h))

    def poisson_solve(self, rhs):
        a = np.array(rhs, dtype=np.float64, order="F", copy=True)
        assert a.shape == (self.imax, self.jtot, self.ktot)
        self._chk(self.L.udgpu_poisson_solve(self.h, a.ctypes.data, a.ctypes.data))
        return a

    def divergence(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.L.udgpu_divergence(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def substep(self, dtmax, ladaptive=False, courant=1.0, diffnr=0.25):
        d, r = C.c_double(self.dt), C.c_int(self.rk3step)
        self._chk(self.L.udgpu_substep(self.h, C.byref(d), C.byref(r), dtmax, int(ladaptive), courant, diffnr))
        self.dt, self.rk3step = d.value, r.value

    def rk3_step_host(self, u0, v0, w0, pres0, dtmax, ladaptive=False, courant=1.0, diffnr=0.25):
        """one full RK3 time step on host arrays (numpy F-ordered, reference shapes, updated in place) or raw
        pointers (ints) to pinned memory"""
        ptrs = [a if isinstance(a, int) else a.ctypes.data for a in (u0, v0, w0, pres0)]
        d = C.c_double(self.dt)
        self._chk(self.L.udgpu_rk3_step_host(self.h, *ptrs, C.byref(d), dtmax, int(ladaptive), courant, diffnr))
        self.dt, self.rk3step = d.value, 3

    # forces (src/modforces.f90:46) -------------------------------------------------------
    def set_forcing(self, dpdxl, dpdyl):
        """large-scale pressure gradient profiles dpdxl(kb:ke+kh), dpdyl(kb:ke+kh): ktot+1 values each"""
        a = np.ascontiguousarray(dpdxl, dtype=np.float64); b = np.ascontiguousarray(dpdyl, dtype=np.float64)
        assert a.size == self.ktot + 1 and b.size == self.ktot + 1
        self._chk(self.L.udgpu_set_forcing(self.h, a.ctypes.data, b.ctypes.data))

    def forces(self): self._chk(self.L.udgpu_forces(self.h))

    # temperature, dry (construct with ltempeq=True) ---------------------------------------------------
    def set_thermo(self, lbuoyancy=True, grav=9.81, thls=288.0, BCtopT=1, wttop=0.0, thl_top=288.0, BCbotT=1, wtsurf=0.0,
                   thlpcar=None):
        a = None if thlpcar is None else np.ascontiguousarray(thlpcar, dtype=np.float64)
        assert a is None or a.size == self.ktot + 1
        self._chk(self.L.udgpu_set_thermo(self.h, int(lbuoyancy), grav, thls, BCtopT, wttop, thl_top, BCbotT, wtsurf,
                                          None if a is None else a.ctypes.data))

    def thermodynamics(self): self._chk(self.L.udgpu_thermodynamics(self.h))
    def set_buoycorr(self, lbuoycorr=True, Rigc=0.25): self._chk(self.L.udgpu_set_buoycorr(self.h, int(lbuoycorr), Rigc))

    def thermo_profile(self, name):
        out = np.empty(self.ktot + 1)
        self._chk(self.L.udgpu_thermo_profile(self.h, {"thl0av": 0, "thvh": 1}[name], out.ctypes.data))
        return out

    # bottom -> wfmneutral (src/modibm.f90:1998, src/modwallfunctions.f90:307) and masscorr (src/modforces.f90:328) --------
    def set_bottom(self, z0, fkar=0.41, lbottom=True, BCbotm=3, BCbots=1):
        self._chk(self.L.udgpu_set_bottom(self.h, int(lbottom), BCbotm, BCbots, z0, fkar))

    def set_wfuno(self, z0h=0.00035, prandtlturb=0.71, grav=9.81, thls=288.0, tcell=288.0):
        """parameters of the wall functions with stability correction (BCbotm = 2, BCbotT = 2)"""
        self._chk(self.L.udgpu_set_wfuno(self.h, z0h, prandtlturb, grav, thls, tcell))

    def bottom(self): self._chk(self.L.udgpu_bottom(self.h))

    def set_masscorr(self, uflowrate=None, vflowrate=None):
        """volume-flow forcing: luvolflowr / lvvolflowr are on for the components whose flow rate is given"""
        self._chk(self.L.udgpu_set_masscorr(self.h, int(uflowrate is not None), int(vflowrate is not None), float(uflowrate or 0.0),
                                            float(vflowrate or 0.0)))

    def masscorr(self, dt, rk3step, want_def=True):
        u, v = C.c_double(), C.c_double()
        self._chk(self.L.udgpu_masscorr(self.h, dt, rk3step, C.byref(u) if want_def else None, C.byref(v) if want_def else None))
        return u.value, v.value

    # immersed boundary masking (src/modibm.f90) ----------------------------------------
    IBM_KINDS = ("solid_u", "solid_v", "solid_w", "solid_c", "bound_u", "bound_v", "bound_w", "bound_c")

    def ibm_set(self, lists):
        """lists: dict kind -> (n,3) int array of LOCAL 1-based (i,j,k) (solid_info%solpts_loc / bound_info%bndpts_loc);
        missing kinds are empty.  Builds the masks on the device and enables the IBM calls of substep()."""
        for kind, name in enumerate(self.IBM_KINDS):
            pts = np.ascontiguousarray(np.asarray(lists.get(name, np.zeros((0, 3))), dtype=np.int32).reshape(-1, 3))
            self._chk(self.L.udgpu_ibm_set_points(self.h, kind, pts.shape[0], pts.ctypes.data, 0))
        self._chk(self.L.udgpu_ibm_commit(self.h))

    def ibm_mask(self, m):
        out = np.empty(self.shape("u0"), dtype=np.float64, order="F")
        self._chk(self.L.udgpu_ibm_pull_mask(self.h, m, out.ctypes.data))
        return out

    def ibmnorm(self): self._chk(self.L.udgpu_ibmnorm(self.h))
    def ibm_diffcorr(self): self._chk(self.L.udgpu_ibm_diffcorr(self.h))

    # measurement ---------------------------------------------------------------
    def profile_enable(self, on=True): self._chk(self.L.udgpu_profile_enable(self.h, int(on)))
    def profile_reset(self): self._chk(self.L.udgpu_profile_reset(self.h))

    def profile_get(self, which):
        ms, n = C.c_double(), C.c_long()
        self._chk(self.L.udgpu_profile_get(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self): return int(self.L.udgpu_launch_count(self.h))

    def trace(self, on=True): self._chk(self.L.udgpu_profile_enable(self.h, 2 if on else 0))
    def trace_dump(self, path): self._chk(self.L.udgpu_trace_dump(self.h, str(path).encode()))

    def stream(self):
        s = C.c_void_p()
        self._chk(self.L.udgpu_stream(self.h, C.byref(s)))
        return s.value
