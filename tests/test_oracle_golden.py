"""Pin the CPU oracle to the REFERENCE: tests/golden/ref_*.npz were produced by executing the reference's
own Fortran source text (oracle/f90run/interp.py + make_golden.py, run in the build container where
/root/reference is mounted).  Every stage of the in-scope substep is compared:
advection -> subgrid (closure, closurebc, diffusion) -> poisson (initpois coefficients, fillps, FFT2D solve,
tderive) -> tstep_integrate -> halos -> boundary, for three RK3 substeps, plus chkdiv and the adaptive dt."""
import glob
import os

import numpy as np
import pytest

from oracle.oracle import Oracle

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")) if not p.endswith(("ref_ibm.npz", "ref_restart102_block.npz", "ref_restart102_turb32.npz", "ref_forces.npz", "ref_channelglue.npz", "ref_thermo_flux.npz", "ref_thermo_value_ibm.npz", "ref_thermo_buoycorr.npz", "ref_thermo_wfuno.npz")))
GOLD_IBM = os.path.join(os.path.dirname(__file__), "golden", "ref_ibm.npz")
TOL = 2e-13      # same arithmetic order, no FMA on either side; FFT library and pow() rounding differ


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def build(d):
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    o = Oracle(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv, BCtopm=int(d["BCtopm"]),
               lvreman=bool(d["lvreman"]), lsmagorinsky=bool(d["lsmagorinsky"]), iadv_sv=int(d["iadv_sv"]) if nsv else 7,
               Uinf=float(d["Uinf"]), Vinf=float(d["Vinf"]))
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "pres0"):
        getattr(o, n)[...] = d["in_" + n]
    if nsv:
        o.sv0[...] = d["in_sv0"][..., :nsv]
        o.svm[...] = d["in_svm"][..., :nsv]
    return o, nsv


def test_goldens_exist():
    assert len(GOLD) >= 5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[4:-4] for p in GOLD])
def test_oracle_matches_reference_source(path):
    d = np.load(path)
    o, nsv = build(d)
    hc = o.ihc
    # initpois coefficients (src/modpois.f90:98-220)
    for nm in ("xrt", "yrt", "a", "b", "c"):
        assert rel(o.metric(nm)[0], d["pois_" + nm]) < 1e-15, nm
    dt = 0.03
    o.dt, o.rk3step = dt, 0
    I, J, K = o.itot, o.jtot, o.ktot
    for s in range(3):
        o.dt, o.rk3step, _, _ = o.tstep_update(o.dt, o.rk3step, dtmax=dt, ladaptive=False)
        o.advection()
        if s == 0:
            for n in ("up", "vp", "wp"):
                assert rel(o.__dict__[n][1:-1, 1:-1, :-1], d["adv_" + n][1:-1, 1:-1, :-1]) < TOL, ("advection", n)
            for n4 in range(nsv):
                assert rel(o.svp[hc:-hc, hc:-hc, :-hc, n4], d["adv_svp"][hc:-hc, hc:-hc, :-hc, n4]) < TOL, ("advection svp", n4)
        o.subgrid()
        if s == 0:
            assert rel(o.ekm, d["sub_ekm"]) < TOL and rel(o.ekh, d["sub_ekh"]) < TOL          # whole arrays incl. ghosts
            for n in ("up", "vp", "wp"):
                assert rel(o.__dict__[n][1:-1, 1:-1, :-1], d["sub_" + n][1:-1, 1:-1, :-1]) < TOL, ("subgrid", n)
            for n4 in range(nsv):
                assert rel(o.svp[hc:-hc, hc:-hc, :-hc, n4], d["sub_svp"][hc:-hc, hc:-hc, :-hc, n4]) < TOL, ("subgrid svp", n4)
        o.poisson(o.dt, o.rk3step)
        if s == 0:
            assert rel(o.rhs, d["pois_rhs"]) < TOL
            assert rel(o.p[1:-1, 1:-1, 1:-1], d["pois_p"][1:-1, 1:-1, 1:-1]) < 1e-11, "p"
            for n in ("up", "vp", "wp"):
                assert rel(o.__dict__[n][1:-1, 1:-1, :-1], d["pois_" + n][1:-1, 1:-1, :-1]) < 1e-11, ("tderive", n)
            assert rel(o.pres0[:, 1:-1, 1:-1], d["pois_pres0"][:, 1:-1, 1:-1]) < 1e-11
            assert rel(o.pres0[1:-1, :, 1:-1], d["pois_pres0"][1:-1, :, 1:-1]) < 1e-11
        o.tstep_integrate(o.dt, o.rk3step)
        o.halos()
        o.boundary()
        assert o.rk3step == int(d[f"s{s + 1}_rk3step"])
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert rel(getattr(o, n), d[f"s{s + 1}_{n}"]) < 1e-11, (s, n)       # whole arrays incl. halos / ghost levels
        for n4 in range(nsv):
            for n in ("sv0", "svm"):
                a, b = getattr(o, n)[..., n4], d[f"s{s + 1}_{n}"][..., n4]
                assert rel(a[:, :, hc:-hc], b[:, :, hc:-hc]) < 1e-11, (s, n, n4)
                e = hc - 1
                assert rel(a[e:a.shape[0] - e, e:a.shape[1] - e, -hc:], b[e:b.shape[0] - e, e:b.shape[1] - e, -hc:]) < 1e-11
    divmax, divtot, _ = o.chkdiv()
    assert abs(divmax - float(d["divmax"])) < 1e-13 and abs(divtot - float(d["divtot"])) < 1e-12
    dtn, _, ct, dn = o.tstep_update(0.05, 0, courant=1.1, diffnr=0.25, dtmax=2.0, ladaptive=True)
    assert ct == pytest.approx(float(d["adapt_courtot"]), rel=1e-11)
    assert dn == pytest.approx(float(d["adapt_diffnrtot"]), rel=1e-11)
    assert dtn == pytest.approx(float(d["adapt_dt"]), rel=1e-11)


def build_ibm(d, cls=Oracle, **kw):
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    o = cls(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv, **kw)
    return o, nsv


def test_oracle_ibm_matches_reference_source():
    """solid / ibmnorm / diffu,v,w,c_corr (src/modibm.f90:748-826, 697-745, 990-1164) and the masks of initibm
    (:153-192): golden produced by executing the reference text on synthetic building blocks."""
    d = np.load(GOLD_IBM)
    o, nsv = build_ibm(d)
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh"):
        getattr(o, n)[...] = d["in_" + n]
    o.sv0[...] = d["in_sv0"][..., :nsv]; o.svm[...] = d["in_svm"][..., :nsv]; o.svp[...] = d["in_svp"][..., :nsv]
    o.ibm_set({k[4:]: d[k] for k in d.files if k.startswith("pts_")})
    for m, nm in enumerate("uvwc"):
        assert np.array_equal(o.ibm_mask(m), d["mask_" + nm]), nm
    o.ibm_diffcorr()
    for n in ("up", "vp", "wp"):
        assert np.array_equal(getattr(o, n), d["corr_" + n]), n          # same operations in the same order: identical bits
    assert np.array_equal(o.svp, d["corr_svp"][..., :nsv])
    o.ibmnorm()
    for n in ("um", "vm", "wm", "up", "vp", "wp"):
        assert np.array_equal(getattr(o, n), d["norm_" + n]), n
    assert np.array_equal(o.svm, d["norm_svm"][..., :nsv]) and np.array_equal(o.svp, d["norm_svp"][..., :nsv])
    # something actually happened
    assert np.abs(d["corr_up"] - d["in_up"]).max() > 1e-6 and np.abs(d["norm_um"] - d["in_um"]).max() > 1e-3


def test_oracle_forces_matches_reference_source():
    """forces, neutral branch (src/modforces.f90:88-125), executed from the reference text."""
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_forces.npz"))
    I, J, K = (int(x) for x in d["shape"])
    o = Oracle(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"])
    for n in ("up", "vp", "wp"):
        getattr(o, n)[...] = d["in_" + n]
    o.set_forcing(d["dpdxl"], d["dpdyl"])
    o.forces()
    for n in ("up", "vp", "wp"):
        assert np.array_equal(getattr(o, n), d["out_" + n]), n
    assert np.abs(d["out_up"] - d["in_up"]).max() > 1e-4


GOLD_GLUE = os.path.join(os.path.dirname(__file__), "golden", "ref_channelglue.npz")


def build_glue(d, cls=Oracle, **kw):
    """state of tests/golden/ref_channelglue.npz in an Oracle (or, with cls = UdalesGPU-like factory, in the CUDA library)"""
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    o = cls(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv, **kw)
    return o, nsv


def test_oracle_bottom_and_masscorr_match_reference_source():
    """bottom -> wfmneutral(91) + scalar bottom correction (src/modibm.f90:1998-2100, src/modwallfunctions.f90:307-349) and
    masscorr's volume-flow branches (src/modforces.f90:394-420, 470-495) against vectors produced by executing the
    reference text (make_golden.py channelglue)."""
    d = np.load(GOLD_GLUE)
    o, nsv = build_glue(d)
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh"):
        getattr(o, n)[...] = d["in_" + n]
    o.sv0[...] = d["in_sv0"][..., :nsv]
    o.svp[...] = d["in_svp"][..., :nsv]
    o.set_bottom(float(d["z0"]), float(d["fkar"]))
    o.bottom()
    for n in ("up", "vp", "wp"):
        assert rel(getattr(o, n), d["bottom_" + n]) < 1e-15, n           # same operation order, no FMA: bit-level agreement
    assert rel(o.svp, d["bottom_svp"][..., :nsv]) < 1e-15
    assert rel(o.momfluxb()[1:-1, 1:-1, 1], d["bottom_momfluxb"][1:-1, 1:-1, 1]) < 1e-15
    o.set_masscorr(float(d["uflowrate"]), float(d["vflowrate"]), d["IIu"], d["IIv"])
    for rk in (1, 2, 3):
        udef, vdef = o.masscorr(float(d["dt"]), rk)
        assert udef == pytest.approx(float(d[f"mc{rk}_udef"]), rel=1e-13)  # slab sums: summation order differs (numpy pairwise)
        assert vdef == pytest.approx(float(d[f"mc{rk}_vdef"]), rel=1e-13)
        assert rel(o.up[1:-1, 1:-1, :-1], d[f"mc{rk}_up"][1:-1, 1:-1, :-1]) < 1e-13
        assert rel(o.vp[1:-1, 1:-1, :-1], d[f"mc{rk}_vp"][1:-1, 1:-1, :-1]) < 1e-13
