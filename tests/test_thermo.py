"""Temperature on the resident path (dry: ltempeq, lbuoyancy, iadv_thl = cd2).

tests/golden/ref_thermo_*.npz were produced by EXECUTING the reference's Fortran text (oracle/f90run/make_golden.py
thermo): advecc_2nd + diffc on thl0 (src/modadvection.f90:67-69, src/modsubgrid.f90:146), bottom with the fixed-flux
temperature branch (src/modibm.f90:2033-2046), forces with buoyancy and radiative tendency (src/modforces.f90:70-109),
tstep_integrate / halos / boundary with BCtopT = flux | value (src/modboundary.f90:208-221), thermodynamics
(src/modthermodynamics.f90:55-121: diagfld, calc_halflev, calthv, thvh), and with IBM diffc_corr(thl0) + ibmnorm's
solid(.., mask_c) + advecc2nd_corr_liberal (src/modibm.f90:714-722, 936-987, 1225), over three RK3 substeps in the order of
src/program.f90:132-212.  The same staged driver checks the CPU oracle (not gpu) and the CUDA library (gpu)."""
import os

import numpy as np
import pytest

from oracle.oracle import Oracle

GOLD = [os.path.join(os.path.dirname(__file__), "golden", f"ref_thermo_{t}.npz") for t in ("flux", "value_ibm")]
STATE = ("u0", "v0", "w0", "um", "vm", "wm", "pres0", "thl0", "thlm")


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


class OracleView:
    """gives the Oracle the push / pull / profile vocabulary of the CUDA library so that one driver serves both"""

    def __init__(self, o):
        self.o = o

    def __getattr__(self, n):
        return getattr(self.o, n)

    def push(self, name, a):
        getattr(self.o, name)[...] = a

    def pull(self, name):
        return getattr(self.o, name)

    def momfluxb(self):
        return self.o.momfluxb()


def setup(d, x):
    K = int(d["shape"][2])
    x.set_thermo(lbuoyancy=True, grav=float(d["grav"]), thls=float(d["thls"]), BCtopT=int(d["BCtopT"]), wttop=float(d["wttop"]),
                 thl_top=float(d["thl_top"]), BCbotT=1, wtsurf=float(d["wtsurf"]), thlpcar=d["thlpcar"])
    x.set_forcing(d["dpdxl"], d["dpdyl"])
    x.set_bottom(float(d["z0"]), float(d["fkar"]))
    if int(d["with_ibm"]):
        x.ibm_set({k[4:]: d[k] for k in d.files if k.startswith("pts_")})
    ekm = np.full(d["in_u0"].shape, 1.5e-5, order="F")
    x.push("ekm", ekm); x.push("ekh", np.asfortranarray(ekm / 0.71))
    for n in STATE:
        x.push(n, d["in_" + n])
    assert K + 1 == d["thlpcar"].size


def drive(d, x, tol, tol_p):
    """the stages of oracle/f90run/make_golden.py:case_thermo on object x (Oracle view or UdalesGPU)"""
    ibm = bool(int(d["with_ibm"]))
    x.thermodynamics()
    for n in ("thvh", "thl0av"):
        assert rel(x.thermo_profile(n), d["in_" + n]) < tol, ("thermodynamics0", n)
    ti = (slice(1, -1), slice(1, -1), slice(0, -1))       # (ib:ie, jb:je, kb:ke) of a tendency-shaped array
    dt = 0.03
    x.dt, x.rk3step = dt, 0
    for s in range(3):
        x.dt, x.rk3step, _, _ = x.tstep_update(x.dt, x.rk3step, dtmax=dt, ladaptive=False)
        x.advection()
        if s == 0:
            assert rel(x.pull("thlp")[ti], d["adv_thlp"][ti]) < tol, "advecc_2nd(thl0)"
        x.subgrid()
        if s == 0:
            assert rel(x.pull("thlp")[ti], d["sub_thlp"][ti]) < tol, "diffc(thl0)"
        x.bottom()
        if s == 0:
            for n in ("up", "vp", "thlp"):
                assert rel(x.pull(n)[ti], d["bottom_" + n][ti]) < tol, ("bottom", n)
        x.forces()
        if s == 0:
            for n in ("up", "vp", "wp", "thlp"):
                assert rel(x.pull(n)[ti], d["forces_" + n][ti]) < tol, ("forces", n)
        if ibm:
            x.ibm_diffcorr()
            if s == 0:
                assert rel(x.pull("thlp")[ti], d["corr_thlp"][ti]) < tol, "diffc_corr(thl0)"
            x.ibmnorm()
            if s == 0:
                for n in ("thlp", "wp"):
                    assert rel(x.pull(n)[ti], d["norm_" + n][ti]) < tol, ("ibmnorm", n)
                assert rel(x.pull("thlm")[1:-1, 1:-1, 1:-1], d["norm_thlm"][1:-1, 1:-1, 1:-1]) < tol, "ibmnorm thlm"
        x.poisson(x.dt, x.rk3step)
        x.tstep_integrate(x.dt, x.rk3step)
        x.halos()
        x.boundary()
        x.thermodynamics()
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert rel(x.pull(n), d[f"s{s + 1}_{n}"]) < tol_p, (s, n)            # whole arrays incl. halos / ghost levels
        for n in ("thl0", "thlm"):
            a, b = x.pull(n), d[f"s{s + 1}_{n}"]
            assert rel(a[:, :, 1:], b[:, :, 1:]) < tol_p, (s, n)                 # incl. lateral halos and the top ghost level
            assert np.array_equal(a[1:-1, 1:-1, 0], d["in_" + n][1:-1, 1:-1, 0])  # thl0(kb-1): startup value, never touched
        for n in ("thvh", "thl0av"):
            assert rel(x.thermo_profile(n), d[f"s{s + 1}_{n}"]) < tol_p, (s, n)
    assert np.abs(d["s3_thl0"] - d["in_thl0"])[1:-1, 1:-1, 1:-1].max() > 1e-4   # something happened
    assert np.abs(d["forces_wp"] - d["bottom_up"] * 0)[ti].max() > 1e-3


@pytest.mark.parametrize("path", GOLD, ids=["flux", "value_ibm"])
def test_oracle_thermo_matches_reference_source(path):
    d = np.load(path)
    I, J, K = (int(v) for v in d["shape"])
    o = Oracle(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"])
    x = OracleView(o)
    setup(d, x)
    drive(d, x, 2e-13, 1e-11)
    # diagnostics only the oracle keeps as 3-D arrays
    for n in ("thl0h", "thv0h", "dthvdz"):
        a, b = getattr(o, n), d["s3_" + n]
        k0 = 1 if n == "thl0h" else 0
        assert rel(a[1:-1, 1:-1, k0:], b[1:-1, 1:-1, k0:]) < 1e-11, n
