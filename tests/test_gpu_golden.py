"""CUDA path vs the golden vectors produced from the reference's Fortran source text (no oracle in between)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")) if not p.endswith(("ref_ibm.npz", "ref_restart102_block.npz", "ref_restart102_turb32.npz", "ref_forces.npz", "ref_channelglue.npz", "ref_thermo_flux.npz", "ref_thermo_value_ibm.npz", "ref_thermo_buoycorr.npz", "ref_thermo_wfuno.npz")))
GOLD_IBM = os.path.join(os.path.dirname(__file__), "golden", "ref_ibm.npz")


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[4:-4] for p in GOLD])
@pytest.mark.parametrize("flags", [0, 1, 8])
def test_cuda_matches_reference_source(path, flags):
    import udales_b200 as U
    d = np.load(path)
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    g = U.UdalesGPU(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv, BCtopm=int(d["BCtopm"]),
                    lvreman=bool(d["lvreman"]), lsmagorinsky=bool(d["lsmagorinsky"]), iadv_sv=int(d["iadv_sv"]) if nsv else 7,
                    Uinf=float(d["Uinf"]), Vinf=float(d["Vinf"]), flags=flags)
    hc = 2 if (nsv and int(d["iadv_sv"]) == 7) else 1
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "pres0"):
        g.push(n, d["in_" + n])
    for n4 in range(nsv):
        g.push("sv0", d["in_sv0"][..., n4], n4)
        g.push("svm", d["in_svm"][..., n4], n4)
    dt = 0.03
    g.dt, g.rk3step = dt, 0
    for s in range(3):
        if s == 0:
            # stage-by-stage on the first substep
            g.dt, g.rk3step, _, _ = g.tstep_update(g.dt, g.rk3step, dtmax=dt, ladaptive=False)
            g.advection(); g.subgrid()
            assert rel(g.pull("ekm"), d["sub_ekm"]) < 1e-12 and rel(g.pull("ekh"), d["sub_ekh"]) < 1e-12
            for n in ("up", "vp", "wp"):
                assert rel(g.pull(n)[1:-1, 1:-1, :-1], d["sub_" + n][1:-1, 1:-1, :-1]) < 1e-12, n
            for n4 in range(nsv):
                assert rel(g.pull("svp", n4)[hc:-hc, hc:-hc, :-hc], d["sub_svp"][hc:-hc, hc:-hc, :-hc, n4]) < 1e-12
            g.poisson(g.dt, g.rk3step)
            assert rel(g.pull("p")[1:-1, 1:-1, 1:-1], d["pois_p"][1:-1, 1:-1, 1:-1]) < 1e-10
            for n in ("up", "vp", "wp"):
                assert rel(g.pull(n)[1:-1, 1:-1, :-1], d["pois_" + n][1:-1, 1:-1, :-1]) < 1e-10, n
            g.tstep_integrate(g.dt, g.rk3step); g.halos(); g.boundary()
        else:
            g.substep(dt)
        assert g.rk3step == int(d[f"s{s + 1}_rk3step"])
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert rel(g.pull(n), d[f"s{s + 1}_{n}"]) < 1e-11, (s, n)
        for n4 in range(nsv):
            a, b = g.pull("sv0", n4), d[f"s{s + 1}_sv0"][..., n4]
            assert rel(a[:, :, hc:-hc], b[:, :, hc:-hc]) < 1e-11, (s, n4)
    divmax, divtot, _ = g.divergence()
    assert abs(divmax - float(d["divmax"])) < 1e-12


def test_cuda_ibm_matches_reference_source():
    """masks, diffu/v/w/c_corr and ibmnorm against vectors produced by executing src/modibm.f90 (no oracle in between)."""
    import udales_b200 as U
    d = np.load(GOLD_IBM)
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    g = U.UdalesGPU(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv)
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh"):
        g.push(n, d["in_" + n])
    for n4 in range(nsv):
        for n in ("sv0", "svm", "svp"):
            g.push(n, d["in_" + n][..., n4], n4)
    g.ibm_set({k[4:]: d[k] for k in d.files if k.startswith("pts_")})
    for m, nm in enumerate("uvwc"):
        assert np.array_equal(g.ibm_mask(m), d["mask_" + nm]), nm
    g.ibm_diffcorr()
    for n in ("up", "vp", "wp"):
        assert rel(g.pull(n), d["corr_" + n]) < 1e-13, n
    for n4 in range(nsv):
        assert rel(g.pull("svp", n4), d["corr_svp"][..., n4]) < 1e-13
    g.ibmnorm()
    for n in ("um", "vm", "wm", "up", "vp", "wp"):
        assert rel(g.pull(n), d["norm_" + n]) < 1e-13, n
    for n4 in range(nsv):
        assert rel(g.pull("svm", n4), d["norm_svm"][..., n4]) < 1e-13
        assert rel(g.pull("svp", n4), d["norm_svp"][..., n4]) < 1e-13


def test_cuda_forces_matches_reference_source():
    """udgpu_forces against the vectors produced by executing src/modforces.f90 (eager application through a pull)."""
    import udales_b200 as U
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_forces.npz"))
    I, J, K = (int(x) for x in d["shape"])
    g = U.UdalesGPU(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"])
    for n in ("up", "vp", "wp"):
        g.push(n, d["in_" + n])
    g.set_forcing(d["dpdxl"], d["dpdyl"])
    g.forces()
    for n in ("up", "vp", "wp"):
        assert np.array_equal(g.pull(n), d["out_" + n]), n      # one subtraction per cell: identical bits


def test_cuda_bottom_and_masscorr_match_reference_source():
    """udgpu_bottom (wfmneutral case 91 + scalar bottom correction) and udgpu_masscorr (volume-flow branches) against the vectors
    produced by executing src/modibm.f90 / src/modwallfunctions.f90 / src/modforces.f90 (no oracle in between).  IIu / IIv of
    the golden case become solid_u / solid_v lists, so the masks are the ones udgpu_ibm_commit builds."""
    import udales_b200 as U
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_channelglue.npz"))
    I, J, K = (int(x) for x in d["shape"])
    nsv = int(d["nsv"])
    g = U.UdalesGPU(I, J, K, xlen=float(d["xlen"]), ylen=float(d["ylen"]), zf=d["zf"], nsv=nsv)
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "up", "vp", "wp", "ekm", "ekh"):
        g.push(n, d["in_" + n])
    for n4 in range(nsv):
        g.push("sv0", d["in_sv0"][..., n4], n4)
        g.push("svp", d["in_svp"][..., n4], n4)
    g.set_bottom(float(d["z0"]), float(d["fkar"]))
    g.bottom()
    for n in ("up", "vp", "wp"):
        assert rel(g.pull(n), d["bottom_" + n]) < 1e-13, n
    for n4 in range(nsv):
        assert rel(g.pull("svp", n4), d["bottom_svp"][..., n4]) < 1e-13
    assert rel(g.pull("momfluxb")[1:-1, 1:-1, 1], d["bottom_momfluxb"][1:-1, 1:-1, 1]) < 1e-13
    pts = lambda II: (np.argwhere(II[:, :, :K] == 0) + 1).astype(np.int32)
    g.ibm_set({"solid_u": pts(d["IIu"]), "solid_v": pts(d["IIv"])})
    g.set_masscorr(float(d["uflowrate"]), float(d["vflowrate"]))
    for rk in (1, 2, 3):
        udef, vdef = g.masscorr(float(d["dt"]), rk)
        assert udef == pytest.approx(float(d[f"mc{rk}_udef"]), rel=1e-12)
        assert vdef == pytest.approx(float(d[f"mc{rk}_vdef"]), rel=1e-12)
        assert rel(g.pull("up")[1:-1, 1:-1, :-1], d[f"mc{rk}_up"][1:-1, 1:-1, :-1]) < 1e-12
        assert rel(g.pull("vp")[1:-1, 1:-1, :-1], d[f"mc{rk}_vp"][1:-1, 1:-1, :-1]) < 1e-12
