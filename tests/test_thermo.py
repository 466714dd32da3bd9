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

GOLD = [os.path.join(os.path.dirname(__file__), "golden", f"ref_thermo_{t}.npz") for t in ("flux", "value_ibm", "buoycorr", "wfuno")]
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
    wf = int(d["wfuno"]) if "wfuno" in d.files else 0
    x.set_thermo(lbuoyancy=True, grav=float(d["grav"]), thls=float(d["thls"]), BCtopT=int(d["BCtopT"]), wttop=float(d["wttop"]),
                 thl_top=float(d["thl_top"]), BCbotT=2 if wf else 1, wtsurf=float(d["wtsurf"]), thlpcar=d["thlpcar"])
    if int(d["lbuoycorr"]):
        x.set_buoycorr(True, float(d["Rigc"]))
    x.set_forcing(d["dpdxl"], d["dpdyl"])
    if wf:       # BCbotm = 2 / BCbotT = 2: wfuno cases 91 / 92 (src/modwallfunctions.f90:24-260)
        x.set_wfuno(z0h=float(d["z0h"]), prandtlturb=float(d["prandtlturb"]), grav=float(d["grav"]), thls=float(d["thls"]))
        x.set_bottom(float(d["z0"]), float(d["fkar"]), BCbotm=2)
    else:
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
            for n in ("ekm", "ekh"):
                assert rel(x.pull(n), d["sub_" + n]) < tol, ("closure", n)      # whole arrays incl. halos and ghost levels
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
    assert np.abs(d["forces_wp"])[ti].max() > 1e-3


@pytest.mark.parametrize("path", GOLD, ids=["flux", "value_ibm", "buoycorr", "wfuno"])
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


# ---- the CUDA library -------------------------------------------------------------------------------------------------
F_NO_LAZY, F_V1, F_NO_HALO = 1, 8, 16


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_NO_HALO, F_V1])
@pytest.mark.parametrize("path", GOLD, ids=["flux", "value_ibm", "buoycorr", "wfuno"])
def test_cuda_thermo_matches_reference_source(path, flags):
    """the same staged comparison, CUDA library against the vectors from the executed reference source (no oracle in between)"""
    import udales_b200 as U
    d = np.load(path)
    I, J, K = (int(v) for v in d["shape"])
    g = U.UdalesGPU(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], ltempeq=True, flags=flags)
    setup(d, g)
    drive(d, g, 1e-12, 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 24, 20), (64, 64, 16)])
@pytest.mark.parametrize("case", ["plain", "value-top", "ibm", "scalars+masscorr", "buoycorr"])
@pytest.mark.parametrize("flags", [0, F_NO_LAZY])
def test_cuda_thermo_substeps_track_oracle(shape, case, flags):
    """six RK3 substeps through udgpu_substep with temperature, buoyancy, bottom (wall function + surface heat flux), forces
    (+ IBM masking / kappa scalars and the volume-flow correction) against the oracle on whole arrays"""
    from helpers import add_thermo, ibm_lists, make_pair
    nsv = 2 if case.startswith("scalars") else 0
    o, g = make_pair(*shape, nsv=nsv, ltempeq_gpu=True, gpu_flags=flags)
    K = shape[2]
    if case == "ibm":
        lists = ibm_lists(*shape, [(5, 9, 4, 8, 5), (20, 24, 15, 20, 7), (shape[0] - 1, shape[0], 1, 3, 4)])
        o.ibm_set(lists); g.ibm_set(lists)
    kw = dict(BCtopT=2, wttop=0.0, wtsurf=-0.004) if case == "value-top" else {}
    if case == "buoycorr":
        o.set_buoycorr(True, 0.25); g.set_buoycorr(True, 0.25)
    add_thermo(o, g, **kw)
    prof = -1e-3 * (1.0 + 0.1 * np.arange(K + 1))
    o.set_forcing(prof, 0.1 * prof); g.set_forcing(prof, 0.1 * prof)
    o.set_bottom(0.01); g.set_bottom(0.01)
    if case.startswith("scalars"):
        o.set_masscorr(1.0, None, None, None); g.set_masscorr(1.0, None)
    dt = 0.02
    o.dt = g.dt = dt
    for s in range(6):
        o.substep(dt); g.substep(dt)
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert rel(g.pull(n), getattr(o, n)) < 1e-11, (s, n)
        for n in ("thl0", "thlm"):
            assert rel(g.pull(n)[:, :, 1:], getattr(o, n)[:, :, 1:]) < 1e-12, (s, n)
        for n in ("thvh", "thl0av"):
            assert rel(g.thermo_profile(n), o.thermo_profile(n)) < 1e-12, (s, n)
        for n4 in range(nsv):
            hc = o.ihc
            assert rel(g.pull("sv0", n4)[:, :, hc:-hc], o.sv0[:, :, hc:-hc, n4]) < 1e-11, (s, "sv0", n4)
    assert g.divergence()[2] < 1e-12
    assert np.abs(o.w0).max() > 1e-3 and np.abs(o.thl0[1:-1, 1:-1, 1:-1] - 288.0).max() > 0.1


@pytest.mark.gpu
def test_thermo_state_errors():
    """forces / ibmnorm refuse to run on a stale thvh; set_thermo needs ltempeq; unsupported switches are EINVAL"""
    import udales_b200 as U
    g = U.UdalesGPU(16, 16, 8)
    with pytest.raises(U.UdalesGPUError):
        g.set_thermo()
    g.close()
    g = U.UdalesGPU(16, 16, 8, ltempeq=True)
    with pytest.raises(U.UdalesGPUError):
        g.set_thermo(BCbotT=3)
    g.set_thermo()
    g.advection(); g.subgrid()
    with pytest.raises(U.UdalesGPUError):
        g.forces()                       # no thermodynamics() since thl0 was (never) set
    g.thermodynamics()
    g.close()
    with pytest.raises(U.UdalesGPUError):
        U.UdalesGPU(16, 16, 8, ltempeq=True, iadv_thl=7)


def test_buoycorr_golden_exercises_the_correction():
    """the lbuoycorr vectors differ from a run without the correction (Rig > 0 somewhere), so the case pins :332-354"""
    d = np.load(GOLD[2])
    I, J, K = (int(v) for v in d["shape"])
    o = Oracle(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"])
    x = OracleView(o)
    setup(d, x)
    o.set_buoycorr(False)
    o.thermodynamics(); o.advection(); o.subgrid()
    assert rel(o.ekm, d["sub_ekm"]) > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("ltempeq", [False, True])
def test_cuda_wfuno_bottom_tracks_oracle(ltempeq):
    """BCbotm = 2 is the namelist default: wfuno case 91 (stability-corrected momentum wall function) in a run WITHOUT
    temperature equation evaluates its Richardson number with the uniform thl0 = thlprof and thls (examples/999: lbottom, no
    thls given) — and with ltempeq case 92 (BCbotT = 2) on top; six substeps against the oracle"""
    from helpers import add_thermo, make_pair
    shape = (32, 24, 20)
    o, g = make_pair(*shape, ltempeq_gpu=ltempeq)
    kw = dict(z0h=0.00035, prandtlturb=0.71, grav=9.81, thls=287.6 if ltempeq else -1.0, tcell=288.0)
    o.set_wfuno(**kw); g.set_wfuno(**kw)
    if ltempeq:
        add_thermo(o, g, BCbotT=2, thls=287.6)
    o.set_bottom(0.05, BCbotm=2); g.set_bottom(0.05, BCbotm=2)
    dt = 0.02
    o.dt = g.dt = dt
    for s in range(6):
        o.substep(dt); g.substep(dt)
        for n in ("u0", "v0", "w0"):
            assert rel(g.pull(n), getattr(o, n)) < 1e-11, (s, n)
        if ltempeq:
            assert rel(g.pull("thl0")[:, :, 1:], o.thl0[:, :, 1:]) < 1e-12, s
    assert rel(g.pull("momfluxb")[1:-1, 1:-1, 1], o.momfluxb()[1:-1, 1:-1, 1]) < 1e-11
    with pytest.raises(Exception):
        g2 = make_pair(16, 16, 8)[1]
        g2.set_bottom(0.05, BCbotm=2); g2.advection(); g2.subgrid(); g2.bottom()      # wfuno without its parameters
