"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Tolerances (fp64, stated by BASELINE.md §4): stencil tendencies rel-L_inf <= 1e-12, pressure
rel-L_inf <= 1e-10, post-projection divergence RMS far below north_star's 1e-6.
"""
import numpy as np
import pytest

from helpers import ibm_lists, interior, make_pair, push_state, relerr, sv_interior, svp_interior, tend_interior

pytestmark = pytest.mark.gpu

TOL_STENCIL = 1e-12
TOL_PRES = 1e-10

F_NO_LAZY, F_V1, F_NO_HALO = 1, 8, 16

SIZES = [(16, 16, 16), (32, 24, 20), (64, 64, 64), (48, 40, 33), (20, 36, 7)]


@pytest.mark.parametrize("shape", SIZES)
@pytest.mark.parametrize("model", ["vreman", "smag", "dns"])
def test_closure(shape, model):
    kw = dict(vreman=dict(), smag=dict(lvreman=False, lsmagorinsky=True), dns=dict(lvreman=False, lsmagorinsky=False))[model]
    o, g = make_pair(*shape, **kw)
    o.closure(); g.closure()
    for name in ("ekm", "ekh"):
        a, b = g.pull(name), getattr(o, name)
        assert relerr(a, b) < TOL_STENCIL, name          # whole array incl. all ghost cells
    # reassure_fluxtop touched the top ghosts of u0/v0
    assert relerr(g.pull("u0"), o.u0) == 0.0


@pytest.mark.parametrize("shape", SIZES + [(96, 40, 12), (36, 18, 5)])
@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=False), dict(BCtopm=2, Uinf=1.0, Vinf=0.2)])
@pytest.mark.parametrize("flags", [0, F_V1])
def test_fused_advection_subgrid(shape, kw, flags):
    """advection(); subgrid() back to back = ONE fused TMA kernel (flags 0) vs the direct kernels (V1)."""
    o, g = make_pair(*shape, gpu_flags=flags, **kw)
    o.advection(); o.subgrid()
    g.advection(); g.subgrid()
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n
    # top ghost level of the tendencies is never touched (stays 0)
    assert np.abs(g.pull("wp")[:, :, -1]).max() == 0.0


@pytest.mark.parametrize("shape", SIZES)
@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=False), dict(BCtopm=2, Uinf=1.0, Vinf=0.2)])
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_V1])
def test_advection_and_subgrid(shape, kw, flags):
    """operator-by-operator (a pull between the calls forces the un-fused, accumulating kernels)."""
    o, g = make_pair(*shape, gpu_flags=flags, **kw)
    o.advection(); g.advection()
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n
    o.subgrid(); g.subgrid()
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n


@pytest.mark.parametrize("shape", SIZES + [(12, 10, 8), (30, 18, 9), (128, 64, 32), (256, 64, 8), (64, 256, 9), (512, 128, 4),
                                           (128, 512, 5), (1024, 64, 3), (64, 1024, 4), (96, 64, 16), (256, 100, 6)])
@pytest.mark.parametrize("flags", [0, F_V1])
def test_poisson_solve(shape, flags):
    o, g = make_pair(*shape, gpu_flags=flags)
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal(shape)
    p_ref = o.poisson_solve(rhs)
    p = g.poisson_solve(rhs)
    assert relerr(p, p_ref) < TOL_PRES


@pytest.mark.parametrize("shape", SIZES)
@pytest.mark.parametrize("rk3step", [1, 2, 3])
def test_poisson_fillps_tderive(shape, rk3step):
    o, g = make_pair(*shape)
    o.advection(); o.subgrid()
    g.advection(); g.subgrid()
    dt = 0.03
    o.fillps(dt, rk3step); g.fillps(dt, rk3step)
    rhs_ref = interior(o.p)
    assert relerr(g.pull("rhs"), rhs_ref) < 1e-11
    o2, g2 = o, g
    # full poisson from the same tendencies
    o2.poisson(dt, rk3step); g2.poisson(dt, rk3step)
    assert relerr(interior(g2.pull("p")), interior(o2.p)) < TOL_PRES
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g2.pull(n)), tend_interior(getattr(o2, n))) < 1e-10, n
    pr, pr_ref = g2.pull("pres0"), o2.pres0
    assert relerr(pr[:, 1:-1, 1:-1], pr_ref[:, 1:-1, 1:-1]) < TOL_PRES     # interior + x-face halos
    assert relerr(pr[1:-1, :, 1:-1], pr_ref[1:-1, :, 1:-1]) < TOL_PRES     # interior + y-face halos


@pytest.mark.parametrize("shape", [(32, 24, 20), (64, 64, 64), (20, 36, 7)])
@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=True), dict(BCtopm=2, Uinf=1.0)])
@pytest.mark.parametrize("flags", [0, F_NO_HALO])
def test_substeps_track_oracle(shape, kw, flags):
    """six RK3 substeps (two full time steps) through the reference call surface."""
    o, g = make_pair(*shape, gpu_flags=flags, **kw)
    dt = 0.02
    o.dt = g.dt = dt
    for s in range(6):
        o.substep(dt); g.substep(dt)
        assert o.rk3step == g.rk3step
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert relerr(g.pull(n), getattr(o, n)) < 1e-11, (s, n)       # whole arrays incl. halos/ghosts
        dmax, dtot, drms = g.divergence()
        omax, otot, orms = o.chkdiv()
        assert drms < 1e-12 and dmax < 1e-11
        assert abs(drms - orms) < 1e-13


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 24, 20), (64, 64, 64), (20, 36, 7), (4, 4, 3), (6, 4, 5)])
@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=True), dict(BCtopm=2, Uinf=1.0, Vinf=0.3),
                                dict(lvreman=False, lsmagorinsky=False)])
def test_halo_ownership_is_bitwise_neutral(shape, kw):
    """closure and the fused tderive+integrate kernel write their own periodic images and ghost levels
    (closurebc, bcp, halos, boundary: src/modboundary.f90:434-505, 1344-1408, 67-109, 163-204); with
    UDGPU_F_NO_HALO_FUSION the separate wrap / ghost kernels run instead.  Both must give identical bits
    in every cell of every array, halos and ghosts included."""
    import udales_b200 as U
    o, ga = make_pair(*shape, **kw)
    _, gb = make_pair(*shape, gpu_flags=F_NO_HALO, **kw)
    dt = 0.02
    ga.dt = gb.dt = dt
    for s in range(5):
        ga.substep(dt); gb.substep(dt)
        for n in ("u0", "v0", "w0", "um", "vm", "wm", "pres0", "ekm", "ekh", "p"):
            a, b = ga.pull(n), gb.pull(n)
            assert np.array_equal(a, b), (s, n, np.abs(a - b).max())
    # the same after the host touched a field mid-run (falls back to the separate kernels until halos+boundary ran)
    u = ga.pull("u0"); ga.push("u0", u); gb.push("u0", u)
    for s in range(3):
        ga.substep(dt); gb.substep(dt)
        for n in ("u0", "w0", "um", "pres0", "ekm"):
            assert np.array_equal(ga.pull(n), gb.pull(n)), (s, n)


def test_tstep_update_adaptive():
    o, g = make_pair(32, 24, 20)
    o.closure(); g.closure()
    d_ref, r_ref, ct_ref, dn_ref = o.tstep_update(0.05, 0, courant=1.1, diffnr=0.25, dtmax=2.0, ladaptive=True)
    d, r, ct, dn = g.tstep_update(0.05, 0, courant=1.1, diffnr=0.25, dtmax=2.0, ladaptive=True)
    assert r == r_ref == 1
    assert ct == pytest.approx(ct_ref, rel=1e-14) and dn == pytest.approx(dn_ref, rel=1e-14)
    assert d == pytest.approx(d_ref, rel=1e-14)
    # rk3step 2, 3: dt untouched
    d2, r2, _, _ = g.tstep_update(d, r, dtmax=2.0)
    assert (d2, r2) == (d, 2)


def test_pull_after_integrate_gives_zero_tendencies():
    o, g = make_pair(16, 16, 16)
    g.dt = 0.02
    g.substep(0.02)
    for n in ("up", "vp", "wp"):
        assert np.abs(g.pull(n)).max() == 0.0      # src/modtstep.f90:322-324


def test_full_size_properties_256():
    """BASELINE config 2 (256^3): size-independent properties instead of the (slow) oracle."""
    import udales_b200 as U
    n = 256
    g = U.UdalesGPU(n, n, n)
    rng = np.random.default_rng(0)
    # linearity + operator inversion of the Poisson solve
    a = rng.standard_normal((n, n, n)); b = rng.standard_normal((n, n, n))
    pa, pb, pab = g.poisson_solve(a), g.poisson_solve(b), g.poisson_solve(a + 2 * b)
    assert relerr(pab, pa + 2 * pb) < 1e-11
    dx = 0.5
    lap = sum((np.roll(pa, -1, ax) - 2 * pa + np.roll(pa, 1, ax)) for ax in (0, 1)) / dx ** 2
    pk = np.concatenate([pa[:, :, :1], pa, pa[:, :, -1:]], axis=2)
    lap += (pk[:, :, 2:] - 2 * pk[:, :, 1:-1] + pk[:, :, :-2]) / dx ** 2
    res = (lap - a)[:, :, :-1]
    assert np.abs(res).max() < 1e-9 * np.abs(a).max()
    # projection property after substeps from an LCG-perturbed channel
    F = lambda: np.zeros((n + 2, n + 2, n + 2), order="F")
    u = F(); v = F(); w = F()
    u[1:-1, 1:-1, 1:-1] = 1.0 + 0.05 * rng.standard_normal((n, n, n))
    v[1:-1, 1:-1, 1:-1] = 0.05 * rng.standard_normal((n, n, n))
    w[1:-1, 1:-1, 2:-1] = 0.05 * rng.standard_normal((n, n, n - 1))
    for nm, f in (("u0", u), ("v0", v), ("w0", w)):
        g.push(nm, f)
    g.halos(); g.boundary()
    for nm in ("u0", "v0", "w0"):
        g.push(nm.replace("0", "m"), g.pull(nm))
    g.dt = 0.05
    for s in range(3):
        g.substep(0.05)
        dmax, dtot, drms = g.divergence()
        assert drms < 1e-12, (s, drms)
    assert np.isfinite(g.pull("u0")).all()


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 24, 20), (48, 40, 33)])
@pytest.mark.parametrize("kw", [dict(nsv=2), dict(nsv=1, iadv_sv=2), dict(nsv=3, lvreman=False, lsmagorinsky=False),
                                dict(nsv=1, lvreman=False, lsmagorinsky=True), dict(nsv=4), dict(nsv=5, iadv_sv=2)])
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_V1])
def test_scalars(shape, kw, flags):
    """advecc_kappa / advecc_2nd + diffc, scalar integrate / halos / top BC."""
    o, g = make_pair(*shape, gpu_flags=flags, **kw)
    hc = o.ihc
    o.advection(); g.advection()
    if flags:   # eager: advection alone is observable
        for n4 in range(o.nsv):
            assert relerr(svp_interior(g.pull("svp", n4), hc), svp_interior(o.svp[..., n4], hc)) < TOL_STENCIL
    o.subgrid(); g.subgrid()
    for n4 in range(o.nsv):
        assert relerr(svp_interior(g.pull("svp", n4), hc), svp_interior(o.svp[..., n4], hc)) < TOL_STENCIL
    dt = 0.02
    o.dt = g.dt = dt
    o2, g2 = make_pair(*shape, gpu_flags=flags, **kw)
    o2.dt = g2.dt = dt
    for s in range(4):
        o2.substep(dt); g2.substep(dt)
        for n4 in range(o2.nsv):
            for nm in ("sv0", "svm"):
                a, b = g2.pull(nm, n4), getattr(o2, nm)[..., n4]
                # interior levels: whole lateral extent incl. the width-hc halos
                assert relerr(a[:, :, hc:-hc], b[:, :, hc:-hc]) < 1e-11, (s, nm, n4)
                # top ghost levels: defined on the momentum-halo footprint only (fluxtopscal, src/modboundary.f90:1530-1531);
                # the bottom ghosts and the outer halo ring of the ghost levels are never written nor read by the path
                e = hc - 1
                fa = a[e:a.shape[0] - e, e:a.shape[1] - e, -hc:]
                fb = b[e:b.shape[0] - e, e:b.shape[1] - e, -hc:]
                assert relerr(fa, fb) < 1e-11, (s, nm, n4, "top ghosts")
        assert relerr(g2.pull("u0"), o2.u0) < 1e-11


IBM_BOXES = {(16, 16, 16): [(4, 7, 3, 6, 5), (12, 16, 10, 12, 3), (1, 2, 15, 16, 7)],
             (32, 24, 20): [(5, 12, 4, 9, 8), (20, 32, 15, 20, 12), (1, 3, 1, 2, 4)],
             (20, 36, 7): [(3, 8, 10, 20, 3), (15, 20, 30, 36, 7)]}


@pytest.mark.parametrize("shape", list(IBM_BOXES))
@pytest.mark.parametrize("nsv", [0, 2])
def test_ibm_masks_diffcorr_ibmnorm(shape, nsv):
    """initibm masks, diffu/v/w/c_corr and ibmnorm (src/modibm.f90:153-192, 990-1164, 697-826) against the oracle."""
    o, g = make_pair(*shape, nsv=nsv)
    lists = ibm_lists(*shape, IBM_BOXES[shape])
    o.ibm_set(lists); g.ibm_set(lists)
    for m in range(4):
        assert np.array_equal(g.ibm_mask(m), o.ibm_mask(m)), m
    o.advection(); o.subgrid(); g.advection(); g.subgrid()
    o.ibm_diffcorr(); g.ibm_diffcorr()
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n
    hc = o.ihc
    for n4 in range(nsv):
        assert relerr(svp_interior(g.pull("svp", n4), hc), svp_interior(o.svp[..., n4], hc)) < TOL_STENCIL
    o.ibmnorm(); g.ibmnorm()
    for n in ("um", "vm", "wm"):
        assert relerr(interior(g.pull(n)), interior(getattr(o, n))) < TOL_STENCIL, n
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n
    for n4 in range(nsv):
        assert relerr(sv_interior(g.pull("svm", n4), hc), sv_interior(o.svm[..., n4], hc)) < TOL_STENCIL
        assert relerr(svp_interior(g.pull("svp", n4), hc), svp_interior(o.svp[..., n4], hc)) < TOL_STENCIL
    # solid points really are masked
    su = lists["solid_u"]
    assert np.abs(g.pull("um")[su[:, 0], su[:, 1], su[:, 2]]).max() == 0.0


@pytest.mark.parametrize("shape", list(IBM_BOXES))
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_NO_HALO])
def test_ibm_substeps_track_oracle(shape, flags):
    """six substeps of the masked channel (BASELINE config 3 in miniature): diff*_corr + ibmnorm between subgrid and
    poisson (src/program.f90:166,171), whole arrays incl. halos and ghost levels."""
    o, g = make_pair(*shape, gpu_flags=flags, nsv=1)
    lists = ibm_lists(*shape, IBM_BOXES[shape])
    o.ibm_set(lists); g.ibm_set(lists)
    dt = 0.02
    o.dt = g.dt = dt
    hc = o.ihc
    for s in range(6):
        o.substep(dt); g.substep(dt)
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert relerr(g.pull(n), getattr(o, n)) < 1e-11, (s, n)
        assert relerr(g.pull("sv0", 0)[:, :, hc:-hc], o.sv0[:, :, hc:-hc, 0]) < 1e-11, s
        # the projected field is divergence free also next to the blocks (the reference does not mask the pressure)
        assert g.divergence()[2] < 1e-12


@pytest.mark.parametrize("shape", [(16, 16, 16), (30, 10, 6), (36, 24, 20), (66, 8, 5), (64, 12, 7)])
@pytest.mark.parametrize("flags", [0, F_V1])
def test_scalars_mixed_sign_flow(shape, flags):
    """kappa scheme with u changing sign cell by cell (upwind direction flips on every face, also on the faces the
    31-cell warps of the marching kernel share with their helper lane and on the periodic boundary)."""
    o, g = make_pair(*shape, gpu_flags=flags, ubase=0.0, nsv=2)
    hc = o.ihc
    o.advection(); g.advection(); o.subgrid(); g.subgrid()
    for n4 in range(2):
        assert relerr(svp_interior(g.pull("svp", n4), hc), svp_interior(o.svp[..., n4], hc)) < TOL_STENCIL
    dt = 0.02
    o.dt = g.dt = dt
    o2, g2 = make_pair(*shape, gpu_flags=flags, ubase=0.0, nsv=2)
    o2.dt = g2.dt = dt
    for s in range(3):
        o2.substep(dt); g2.substep(dt)
        for n4 in range(2):
            assert relerr(g2.pull("sv0", n4)[:, :, hc:-hc], o2.sv0[:, :, hc:-hc, n4]) < 1e-11, (s, n4)


def test_reference_restart_block_divergence_on_gpu():
    """udgpu_divergence on a block of the reference binary's own restart output (examples/102): round-off, as the
    reference's projection left it."""
    import udales_b200 as U
    from test_oracle import load_restart_block
    n, f, dglob = load_restart_block(None)
    g = U.UdalesGPU(n, n, n, xlen=float(n), ylen=float(n), zf=np.arange(n) + 0.5)
    for nm, a in f.items():
        g.push(nm, a)
    dmax, dtot, drms = g.divergence()
    assert dmax < 5e-15 and drms < 1e-15


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 24, 20)])
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_NO_HALO])
@pytest.mark.parametrize("ibm", [False, True])
def test_forces_in_the_substep(shape, flags, ibm):
    """forces (src/modforces.f90:46, src/program.f90:158) inside the resident substep: applied lazily inside the fused
    tderive+integrate kernel (flags 0), eagerly (NO_LAZY / with IBM masking, where ibmnorm must see the forced tendencies)."""
    o, g = make_pair(*shape, gpu_flags=flags)
    K = shape[2]
    rng = np.random.default_rng(2)
    fx, fy = -2e-3 * (1 + rng.random(K + 1)), 5e-4 * rng.standard_normal(K + 1)
    o.set_forcing(fx, fy); g.set_forcing(fx, fy)
    if ibm:
        lists = ibm_lists(*shape, IBM_BOXES[shape])
        o.ibm_set(lists); g.ibm_set(lists)
    dt = 0.02
    o.dt = g.dt = dt
    for s in range(4):
        o.substep(dt); g.substep(dt)
        for n in ("u0", "v0", "w0", "um"):
            assert relerr(g.pull(n), getattr(o, n)) < 1e-11, (s, n)
        assert g.divergence()[2] < 1e-12
    # observable in between: pull after forces() shows the forced tendencies
    o.advection(); o.subgrid(); o.forces()
    g.advection(); g.subgrid(); g.forces()
    for n in ("up", "vp", "wp"):
        assert relerr(tend_interior(g.pull(n)), tend_interior(getattr(o, n))) < TOL_STENCIL, n


# ---- round 2: entry points the bench's end-to-end number goes through, and lazy-state bookkeeping ------------------
def test_push_of_a_non_tendency_keeps_the_pending_zero_fill():
    """tstep_integrate leaves up = vp = wp = 0 (src/modtstep.f90:322-324); the library writes those zeros lazily.
    Pushing a field that is NOT a tendency must not cancel the pending fill (round-1 bug at udgpu_push)."""
    o, g = make_pair(16, 16, 16, nsv=1)
    g.dt = 0.02
    g.substep(0.02)
    g.push("u0", g.pull("u0"))                  # non-tendency push while the zero-fill is pending
    g.push("ekm", g.pull("ekm"))
    for n in ("up", "vp", "wp"):
        assert np.abs(g.pull(n)).max() == 0.0, n
    assert np.abs(g.pull("svp", 0)).max() == 0.0
    # and a pushed tendency is kept (not zeroed) and accumulated into by the next operator, like the reference would
    g.substep(0.02)
    t = np.asfortranarray(np.full(g.shape("svp"), 0.25))
    g.push("svp", t, 0)
    assert np.array_equal(g.pull("svp", 0), t)
    u = np.asfortranarray(np.full(g.shape("up"), 0.5))
    g.push("up", u)
    assert np.array_equal(g.pull("up"), u)


@pytest.mark.parametrize("shape", [(32, 24, 20), (64, 64, 32)])
@pytest.mark.parametrize("ladaptive", [False, True])
def test_rk3_step_host_equals_three_substeps_and_the_oracle(shape, ladaptive):
    """udgpu_rk3_step_host (the entry point of bench.py's end-to-end number): u0,v0,w0,pres0 on HOST arrays in, one RK3
    time step, the same four out == three resident substeps == the oracle."""
    o, ga = make_pair(*shape)
    _, gb = make_pair(*shape)
    dt0, dtmax = 0.02, (0.05 if ladaptive else 0.02)
    o.dt = ga.dt = gb.dt = dt0
    host = {n: np.array(getattr(o, n), order="F", copy=True) for n in ("u0", "v0", "w0", "pres0")}
    for step in range(2):
        for s in range(3):
            o.substep(dtmax, ladaptive=ladaptive, courant=1.1, diffnr=0.25)
            gb.substep(dtmax, ladaptive=ladaptive, courant=1.1, diffnr=0.25)
        ga.rk3_step_host(host["u0"], host["v0"], host["w0"], host["pres0"], dtmax=dtmax, ladaptive=ladaptive, courant=1.1, diffnr=0.25)
        assert ga.dt == pytest.approx(o.dt, rel=1e-12) and gb.dt == pytest.approx(o.dt, rel=1e-12)
        for n in ("u0", "v0", "w0"):
            assert np.array_equal(host[n], gb.pull(n)), (step, n)          # same kernels, same bits
            assert relerr(host[n], getattr(o, n)) < 1e-11, (step, n)       # whole array incl. halos and ghost levels
        pr = host["pres0"]
        assert np.array_equal(pr, gb.pull("pres0"))
        assert relerr(pr[:, 1:-1, 1:-1], o.pres0[:, 1:-1, 1:-1]) < TOL_PRES
        assert relerr(pr[1:-1, :, 1:-1], o.pres0[1:-1, :, 1:-1]) < TOL_PRES


@pytest.mark.parametrize("shape", [(32, 24, 20), (128, 64, 24), (96, 40, 12)])
def test_poisson_solve_resident_and_device_ptr(shape):
    """udgpu_poisson_solve_resident on the resident rhs buffer == udgpu_poisson_solve with host buffers == the oracle;
    udgpu_device_ptr hands out the very buffer (written through a raw device pointer with torch)."""
    import ctypes as C
    import torch
    o, g = make_pair(*shape)
    rng = np.random.default_rng(9)
    rhs = np.asfortranarray(rng.standard_normal(shape))
    p_ref = o.poisson_solve(rhs)
    p_host = g.poisson_solve(rhs)
    g.push("rhs", rhs)
    g.poisson_solve_resident()
    p_res = g.pull("rhs")
    assert np.array_equal(p_res, p_host)
    assert relerr(p_res, p_ref) < TOL_PRES
    # device pointer: fill the rhs buffer from the device side, solve, read it back through the pointer
    dptr = C.c_void_p()
    g._chk(g.L.udgpu_device_ptr(g.h, 13, 0, C.byref(dptr)))     # UDGPU_RHS
    n = rhs.size
    src = torch.from_numpy(np.ascontiguousarray(rhs.ravel(order="F"))).cuda()
    g.sync()
    rt = C.CDLL("libcudart.so.12")          # the runtime the library itself is linked against (already loaded)
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rc = rt.cudaMemcpy(dptr.value, src.data_ptr(), n * 8, 3)   # cudaMemcpyDeviceToDevice
    assert int(rc) == 0
    torch.cuda.synchronize()
    g.poisson_solve_resident()
    g.sync()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    rc = rt.cudaMemcpy(out.data_ptr(), dptr.value, n * 8, 3)
    assert int(rc) == 0
    assert np.array_equal(out.cpu().numpy().reshape(shape, order="F"), p_host)
    # a tendency handed out by pointer counts as written by the host: no lazy zero-fill over it
    g.dt = 0.02
    g.substep(0.02)
    g._chk(g.L.udgpu_device_ptr(g.h, 6, 0, C.byref(dptr)))      # UDGPU_UP (materialises the zeros first)
    assert np.abs(g.pull("up")).max() == 0.0


def test_full_size_256_against_the_oracle_directly():
    """BASELINE config 2 at its full size, compared with the oracle cell by cell: closure, the fused tendencies, the
    Poisson solve inside a substep and the integrated state after one RK3 substep of 256^3."""
    n = 256
    o, g = make_pair(n, n, n, stretched=False)
    dt = 0.05
    o.dt = g.dt = dt
    o.substep(dt); g.substep(dt)
    for nm in ("u0", "v0", "w0"):
        assert relerr(g.pull(nm), getattr(o, nm)) < 1e-11, nm
    assert relerr(g.pull("ekm"), o.ekm) < TOL_STENCIL
    assert relerr(interior(g.pull("p")), interior(o.p)) < TOL_PRES
    pr = g.pull("pres0")
    assert relerr(pr[:, 1:-1, 1:-1], o.pres0[:, 1:-1, 1:-1]) < TOL_PRES
    dmax, dtot, drms = g.divergence()
    omax, otot, orms = o.chkdiv()
    assert drms < 1e-12 and abs(dmax - omax) < 1e-11


@pytest.mark.parametrize("shape", [(16, 16, 16), (32, 24, 20)])
@pytest.mark.parametrize("flags", [0, F_NO_LAZY, F_NO_HALO])
@pytest.mark.parametrize("case", ["channel999", "u_only", "ibm"])
def test_resident_channel_glue_in_the_substep(shape, flags, case):
    """examples/999 physics resident on the device: bottom -> wfmneutral, forces, masscorr (volume flow) inside the substep
    (src/program.f90:152,158,169), lazily folded into the fused tderive+integrate kernel (flags 0) or eager (NO_LAZY / with IBM
    masking, where ibmnorm follows masscorr).  Six substeps against the oracle, whole arrays."""
    nsv = 1 if case != "u_only" else 0
    o, g = make_pair(*shape, gpu_flags=flags, nsv=nsv)
    K = shape[2]
    rng = np.random.default_rng(4)
    fx, fy = -1e-3 * (1 + rng.random(K + 1)), 2e-4 * rng.standard_normal(K + 1)
    for x in (o, g):
        x.set_bottom(0.01, 0.41)
        if case == "u_only":
            x.set_masscorr(uflowrate=1.1)
        else:
            x.set_masscorr(uflowrate=1.1, vflowrate=0.05)
            x.set_forcing(fx, fy)
    if case == "ibm":
        lists = ibm_lists(*shape, IBM_BOXES[shape])
        g.ibm_set(lists); o.ibm_set(lists)
        I, J = shape[:2]
        o.set_masscorr(1.1, 0.05, o.ibm_mask(0)[1:-1, 1:-1, 1:].astype(np.int32), o.ibm_mask(1)[1:-1, 1:-1, 1:].astype(np.int32))
    dt = 0.02
    o.dt = g.dt = dt
    hc = o.ihc
    for s in range(6):
        o.substep(dt); g.substep(dt)
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            assert relerr(g.pull(n), getattr(o, n)) < 1e-11, (s, n)
        for n4 in range(nsv):
            assert relerr(g.pull("sv0", n4)[:, :, hc:-hc], o.sv0[:, :, hc:-hc, n4]) < 1e-11, s
        assert g.divergence()[2] < 1e-12
    # the bulk velocity is what masscorr was told to hold (fluid volume mean of u after a full RK3 step)
    assert relerr(interior(g.pull("momfluxb"))[:, :, 0], interior(o.momfluxb())[:, :, 0]) < 1e-11


@pytest.mark.parametrize("shape", [(64, 64, 16), (128, 64, 9), (256, 64, 8), (64, 128, 5)])
@pytest.mark.parametrize("rk3step", [1, 3])
def test_fillps_fused_into_the_first_transform_is_bitwise_neutral(shape, rk3step, monkeypatch):
    """poisson() with fillps + bcpup evaluated inside the first forward transform (no rhs array traffic) against the
    separate k_fillps pass: the same expression in the same order, so p, the projected tendencies and pres0 are the
    same bits; and against the oracle."""
    monkeypatch.setenv("UDGPU_FILL_FUSED", "1")
    o, ga = make_pair(*shape)
    monkeypatch.setenv("UDGPU_FILL_FUSED", "0")
    _, gb = make_pair(*shape)
    dt = 0.03
    for x in (o, ga, gb):
        x.advection(); x.subgrid(); x.poisson(dt, rk3step)
    for n in ("p", "up", "vp", "wp", "pres0"):
        assert np.array_equal(ga.pull(n), gb.pull(n)), n
    assert relerr(interior(ga.pull("p")), interior(o.p)) < TOL_PRES


@pytest.mark.parametrize("shape", [(64, 64, 64), (32, 24, 16), (128, 64, 128), (64, 64, 256), (30, 18, 48), (64, 32, 512)])
@pytest.mark.parametrize("variant", ["stream", "L8", "L16", "L32"])
def test_one_pass_segmented_z_solve(shape, variant, monkeypatch):
    """k_zsolve_seg (one pass over HBM, segments of the z recurrence combined through shared memory) against the oracle's
    solmpj restatement (src/modpois.f90:1107-1166) and against the streaming two-sweep kernel, for every instantiated
    segment length (levels per thread); shapes include tiles that run past the end of the plane (30 x 18)."""
    if variant == "stream":
        monkeypatch.setenv("UDGPU_ZSEG", "0")
    else:
        monkeypatch.setenv("UDGPU_ZSEG", "1")
        monkeypatch.setenv("UDGPU_ZSEG_L", variant[1:])
    st = shape[2] <= 256          # 1.04 ** 512: the stretched grid itself is ill-conditioned beyond 1e-10
    o, g = make_pair(*shape, stretched=st)
    rng = np.random.default_rng(11)
    rhs = rng.standard_normal(shape)
    rhs -= rhs.mean()
    p_ref = o.poisson_solve(rhs)
    p = g.poisson_solve(rhs)
    assert relerr(p, p_ref) < TOL_PRES
    monkeypatch.setenv("UDGPU_ZSEG", "0")
    o2, g2 = make_pair(*shape, stretched=st)
    assert relerr(p, g2.poisson_solve(rhs)) < 1e-11


@pytest.mark.parametrize("shape", [(64, 64, 16), (128, 40, 9), (256, 64, 8), (512, 36, 5), (1024, 64, 3), (256, 100, 6), (64, 30, 7)])
def test_x_transform_with_line_local_threads_is_bitwise_neutral(shape, monkeypatch):
    """k_rfft_xline (a line's threads in neighbouring lanes, no staging tile, warp-level exchanges) performs the same
    operations on every value as the staged k_rfft_fast x pass: the solved pressure is identical bit for bit, for every
    fast length, partial batches (jmax not a multiple of the lines per CTA) and the halo'd (8-byte misaligned) last pass."""
    rng = np.random.default_rng(21)
    rhs = rng.standard_normal(shape)
    monkeypatch.setenv("UDGPU_XLINE", "1")
    o, g = make_pair(*shape)
    p1 = g.poisson_solve(rhs)
    g.advection(); g.subgrid(); g.poisson(0.02, 1)
    q1 = g.pull("p")
    assert relerr(p1, o.poisson_solve(rhs)) < TOL_PRES
    monkeypatch.setenv("UDGPU_XLINE", "0")
    o2, g2 = make_pair(*shape)
    p0 = g2.poisson_solve(rhs)
    g2.advection(); g2.subgrid(); g2.poisson(0.02, 1)
    q0 = g2.pull("p")
    assert np.array_equal(p1, p0)
    assert np.array_equal(interior(q1), interior(q0))


@pytest.mark.parametrize("kind", ["channel", "scalars", "ibm", "thermo"])
def test_slab_parity_cases_on_one_gpu(kind):
    """the cases of tests/slab_parity.py (what tests/mgpu_worker.py runs on 2+ GPUs and bench.py runs before its timed
    region) with a single slab: channel incl. Poisson alone and adaptive dt, 2 kappa scalars, IBM blocks across the
    periodic seam, temperature + buoyancy + forcing + volume-flow correction over IBM blocks"""
    import udales_b200 as U
    from oracle.oracle import Oracle, stretched_zf
    from slab_parity import TOL, run_case
    shape = (64, 64, 32) if kind == "channel" else (64, 64, 16)
    e = run_case(U, Oracle, kind, shape, 1, 0, 0, None, nsub=3, stretched_zf=stretched_zf)
    assert e < TOL


def test_sparse_pull_and_add_points():
    """udgpu_pull_points / udgpu_add_points: what a host-side wall function needs instead of whole arrays — values of u0 at a
    point list equal the pulled array, additions to a tendency land where they should (duplicates accumulate), a lazily
    pending forces() survives the addition, and out-of-range offsets are refused"""
    import udales_b200 as U
    o, g = make_pair(32, 24, 20)
    rng = np.random.default_rng(4)
    pts = np.stack([rng.integers(0, 34, 500), rng.integers(0, 26, 500), rng.integers(0, 22, 500)], axis=1)
    u = g.pull("u0")
    assert np.array_equal(g.pull_points("u0", pts), u[pts[:, 0], pts[:, 1], pts[:, 2]])
    g.advection(); g.subgrid()
    prof = -1e-3 * np.ones(21)
    g.set_forcing(prof, 0.0 * prof)
    g.forces()                                           # stays lazily pending
    tp = pts.copy(); tp[:, 2] = np.minimum(tp[:, 2], 20)  # tendency arrays have ktot + 1 levels
    tp = np.concatenate([tp, tp[:50]])                   # duplicates
    vals = rng.standard_normal(tp.shape[0])
    g2 = make_pair(32, 24, 20)[1]
    g2.advection(); g2.subgrid(); g2.set_forcing(prof, 0.0 * prof); g2.forces()
    ref = g2.pull("up")
    np.add.at(ref, (tp[:, 0], tp[:, 1], tp[:, 2]), vals)
    g.add_points("up", tp, vals)
    assert np.abs(g.pull("up") - ref).max() < 1e-13
    with pytest.raises(U.UdalesGPUError):
        g.add_points("u0", tp, vals)                     # not a tendency
    with pytest.raises(U.UdalesGPUError):
        g.pull_points("u0", np.array([[0, 0, 23]]))      # level beyond the array: offset past the end
