"""Restart adapter (u-dales_b200/restart.py): the reference's unformatted initd / inits layout (src/modsave.f90:78-122,
src/modstartup.f90:2156-2221).  CPU tests: record round trip, file names, and — in the build container, where
/root/reference is mounted — the reference binary's own files of examples/102 against the committed fixture.
GPU tests: warm start of the resident state, write-out, and a parity run on the REAL turbulent field."""
import os

import numpy as np
import pytest

import udales_b200 as U
from udales_b200 import restart as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF102 = "/root/reference/examples/102/warmstart_files"


def test_initd_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    I, J, K = 6, 4, 5
    f = {nm: rng.standard_normal((I + 2, J + 2, K + 1)) for nm in R.FIELDS_D}
    f["mindist"] = rng.random((I, J, K)); f["wall"] = rng.integers(0, 9, (I, J, K, 5)).astype(np.int32)
    p = tmp_path / R.restart_name("d", 267, 1, 0, 102)
    assert p.name == "initd00000267_001_000.102"
    R.write_initd(p, f, 100.25, 0.43)
    # record structure of src/modsave.f90:87-99: 13 records with the reference's byte counts
    rec = R._records(p)
    assert [len(r) for r in rec] == [I * J * K * 8, I * J * K * 5 * 4] + [(I + 2) * (J + 2) * (K + 1) * 8] * 10 + [16]
    d = R.read_initd(p, I, J, K)
    for nm in R.FIELDS_D + ("mindist", "wall"):
        assert np.array_equal(d[nm], f[nm]), nm
    assert (d["timee"], d["dt"]) == (100.25, 0.43)
    sv = rng.standard_normal((I + 2, J + 2, K + 1, 3))
    ps = tmp_path / R.restart_name("s", 267, 1, 0, 102)
    R.write_inits(ps, sv, 100.25)
    sv2, t = R.read_inits(ps, I, J, K, 3)
    assert np.array_equal(sv, sv2) and t == 100.25
    with pytest.raises(ValueError):
        R.read_initd(p, I + 1, J, K)


@pytest.mark.skipif(not os.path.isdir(REF102), reason="reference tree not mounted (build container only)")
def test_reads_the_reference_binarys_own_restart_files():
    glob, timee, dt = R.assemble(REF102, 267, 102, 64, 64, 64, 2, 2)
    assert timee == pytest.approx(100.2639, abs=1e-3) and dt == pytest.approx(0.42978, abs=1e-4)
    fx = np.load(os.path.join(GOLD, "ref_restart102_turb32.npz"))
    i0, j0, n = int(fx["i0"]), int(fx["j0"]), int(fx["n"])
    for nm in ("u0", "v0", "w0", "pres0", "thl0"):
        assert np.array_equal(glob[nm][i0 - 1:i0 + n + 1, j0 - 1:j0 + n + 1, 0:n + 1], fx[nm]), nm
    th = glob["thl0"][1:-1, 1:-1, :-1]
    assert 287.0 < th.min() and th.max() < 289.5 and th.std() > 0.01          # examples/102: thl0 = 288 K plus the heated-surface signal
    u, v, w = glob["u0"], glob["v0"], glob["w0"]
    div = (u[2:, 1:-1, :-1] - u[1:-1, 1:-1, :-1]) + (v[1:-1, 2:, :-1] - v[1:-1, 1:-1, :-1]) + (w[1:-1, 1:-1, 1:] - w[1:-1, 1:-1, :-1])
    assert np.abs(div).max() < 5e-15


@pytest.mark.gpu
def test_warm_start_and_write_out_on_the_gpu(tmp_path):
    """load a restart state into the resident fields, step, write the rank file, read it back == what the device holds"""
    fx = np.load(os.path.join(GOLD, "ref_restart102_turb32.npz"))
    n = int(fx["n"])
    g = U.UdalesGPU(n, n, n, xlen=float(n), ylen=float(n), zf=np.arange(n) + 0.5, nsv=1)
    glob = {nm: fx[nm] for nm in ("u0", "v0", "w0", "pres0")}
    R.load_into(g, glob, timee=float(fx["timee"]), dt=0.05)
    assert np.array_equal(g.pull("u0")[1:-1, 1:-1, 1:-1], fx["u0"][1:-1, 1:-1, :-1])
    assert np.array_equal(g.pull("um"), g.pull("u0"))
    for _ in range(3):
        g.substep(0.05)
    p = R.save_from(g, tmp_path, 268, 102, 100.5)
    assert os.path.basename(p) == "initd00000268_000_000.102"
    d = R.read_initd(p, n, n, n)
    for nm in ("u0", "v0", "w0", "pres0", "ekm"):
        assert np.array_equal(d[nm], g.pull(nm)[:, :, 1:]), nm
    assert np.all(d["thl0"] == 0.0) and d["timee"] == 100.5
    sv, t = R.read_inits(os.path.join(tmp_path, R.restart_name("s", 268, 0, 0, 102)), n, n, n, 1)
    assert np.array_equal(sv[..., 0], g.pull("sv0", 0)[1:-1, 1:-1, 2:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=True)])
def test_parity_on_the_reference_binarys_turbulent_field(kw):
    """closure and three substeps on a 32^3 block of the REAL 64^3 LES state the reference binary wrote
    (examples/102/warmstart_files, tests/golden/ref_restart102_turb32.npz) instead of synthetic noise: CUDA == oracle."""
    from oracle.oracle import Oracle
    fx = np.load(os.path.join(GOLD, "ref_restart102_turb32.npz"))
    n = int(fx["n"])
    zf = np.arange(n) + 0.5                       # dx = dy = dz = 1 m (examples/102/namoptions.102)
    o = Oracle(n, n, n, xlen=float(n), ylen=float(n), zf=zf, **kw)
    g = U.UdalesGPU(n, n, n, xlen=float(n), ylen=float(n), zf=zf, **kw)
    for nm in ("u0", "v0", "w0", "pres0"):
        getattr(o, nm)[...] = 0.0
        getattr(o, nm)[:, :, 1:] = fx[nm]
    o.halos(); o.boundary()
    o.um[...] = o.u0; o.vm[...] = o.v0; o.wm[...] = o.w0
    for nm in ("u0", "v0", "w0", "um", "vm", "wm", "pres0"):
        g.push(nm, getattr(o, nm))
    o.closure(); g.closure()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(g.pull("ekm"), o.ekm) < 1e-12 and rel(g.pull("ekh"), o.ekh) < 1e-12
    assert o.ekm.max() > 50 * 1.5e-5              # the eddy viscosity of a developed LES field, not molecular noise
    dt = float(fx["dt"]) * 0.5
    o.dt = g.dt = dt
    for s in range(3):
        o.substep(dt); g.substep(dt)
        for nm in ("u0", "v0", "w0", "um"):
            assert rel(g.pull(nm), getattr(o, nm)) < 1e-11, (s, nm)
        assert g.divergence()[2] < 1e-12


@pytest.mark.gpu
def test_temperature_run_on_the_reference_binarys_turbulent_field(tmp_path):
    """examples/102 is a temperature case (ltempeq, lbuoyancy, wtsurf = 0.01, wttop = -0.01): warm-start velocity AND thl0 of the
    reference binary's own restart state (32^3 block), run three substeps with buoyancy, the surface heat flux and the wall
    function on the device and in the oracle, write the rank file: thl0 is in it."""
    from oracle.oracle import Oracle
    fx = np.load(os.path.join(GOLD, "ref_restart102_turb32.npz"))
    n = int(fx["n"])
    zf = np.arange(n) + 0.5
    g = U.UdalesGPU(n, n, n, xlen=float(n), ylen=float(n), zf=zf, ltempeq=True)
    o = Oracle(n, n, n, xlen=float(n), ylen=float(n), zf=zf)
    kw = dict(lbuoyancy=True, thls=288.0, BCtopT=1, wttop=-0.01, BCbotT=1, wtsurf=0.01)
    g.set_thermo(**kw); o.set_thermo(**kw)
    g.set_bottom(0.01); o.set_bottom(0.01)
    R.load_into(g, {nm: fx[nm] for nm in ("u0", "v0", "w0", "pres0", "thl0")}, timee=float(fx["timee"]), dt=0.1)
    for nm in ("u0", "v0", "w0", "um", "vm", "wm", "pres0", "thl0", "thlm", "ekm", "ekh"):
        getattr(o, nm)[...] = g.pull(nm)
    o.thermodynamics()
    assert np.array_equal(g.pull("thl0")[1:-1, 1:-1, 1:-1], fx["thl0"][1:-1, 1:-1, :-1])
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(g.thermo_profile("thvh"), o.thermo_profile("thvh")) < 1e-13
    dt = 0.1
    o.dt = g.dt = dt
    for s in range(3):
        o.substep(dt); g.substep(dt)
        for nm in ("u0", "v0", "w0"):
            assert rel(g.pull(nm), getattr(o, nm)) < 1e-11, (s, nm)
        assert rel(g.pull("thl0")[:, :, 1:], o.thl0[:, :, 1:]) < 1e-13, s
        assert g.divergence()[2] < 1e-12
    assert np.abs(o.thl0[1:-1, 1:-1, 1:-1] - fx["thl0"][1:-1, 1:-1, :-1]).max() > 1e-3     # the temperature field moved
    p = R.save_from(g, tmp_path, 268, 102, 100.6)
    d = R.read_initd(p, n, n, n)
    assert np.array_equal(d["thl0"], g.pull("thl0")[:, :, 1:])
