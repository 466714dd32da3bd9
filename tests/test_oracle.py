"""CPU checks of the oracle itself: conventions that numpy / analysis can pin independently."""
import numpy as np
import pytest

from oracle.oracle import Oracle, rfft_packed, stretched_zf


@pytest.mark.parametrize("n", [4, 6, 8, 12, 30, 64, 100, 256, 1024])
def test_rfft_packed_matches_numpy(n):
    """FFTW r2c + the reference's packing and 1/sqrt(n) (src/modpois.f90:478-490) vs numpy.fft (pocketfft)."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    F = np.fft.rfft(x) / np.sqrt(n)
    h = n // 2
    ref = np.empty(n)
    ref[0] = F[0].real
    ref[1:n - 1:2] = F[1:h].real
    ref[2:n - 1:2] = F[1:h].imag
    ref[n - 1] = F[h].real
    y = rfft_packed(x)
    assert np.abs(y - ref).max() < 5e-14 * max(1.0, np.abs(ref).max())
    # inverse: c2r unnormalised * 1/sqrt(n)  (src/modpois.f90:669-679) == numpy irfft * n / sqrt(n)
    xb = rfft_packed(y, inverse=True)
    assert np.abs(xb - x).max() < 5e-14 * np.abs(x).max()
    G = np.zeros(h + 1, dtype=complex)
    G[0] = ref[0]; G[1:h] = ref[1:n - 1:2] + 1j * ref[2:n - 1:2]; G[h] = ref[n - 1]
    xb2 = np.fft.irfft(G, n) * n / np.sqrt(n)
    assert np.abs(rfft_packed(ref, inverse=True) - xb2).max() < 5e-14 * np.abs(x).max()


def _apply_operator(o, p):
    """7-point operator the FFT2D solver inverts, Neumann in z, periodic x/y."""
    I, J, K = p.shape
    dzf, lo = o.metric("dzf"); dzh, lo2 = o.metric("dzh")
    lap = (np.roll(p, -1, 0) - 2 * p + np.roll(p, 1, 0)) / o.dx ** 2 \
        + (np.roll(p, -1, 1) - 2 * p + np.roll(p, 1, 1)) / o.dy ** 2
    pk = np.concatenate([p[:, :, :1], p, p[:, :, -1:]], axis=2)
    for k in range(K):
        kk = k + 1
        lap[:, :, k] += ((pk[:, :, kk + 1] - pk[:, :, kk]) / dzh[kk + 1 - lo2]
                         - (pk[:, :, kk] - pk[:, :, kk - 1]) / dzh[kk - lo2]) / dzf[kk - lo]
    return lap


@pytest.mark.parametrize("shape", [(16, 12, 10), (8, 8, 8), (32, 20, 6)])
def test_poisson_inverts_discrete_operator(shape):
    I, J, K = shape
    o = Oracle(I, J, K, xlen=8.0, ylen=6.0, zf=stretched_zf(K, 5.0, 1.1))
    rng = np.random.default_rng(1)
    rhs = rng.standard_normal(shape)
    dzf, lo = o.metric("dzf")
    w = dzf[1 - lo:K + 1 - lo]
    rhs -= (rhs.mean(axis=(0, 1)) * w).sum() / w.sum()       # compatible rhs
    p = o.poisson_solve(rhs)
    res = _apply_operator(o, p) - rhs
    assert np.abs(res).max() < 1e-11 * np.abs(rhs).max()


def test_poisson_eigenfunction():
    """cos modes are eigenfunctions: solution = rhs / (lambda_x + lambda_y) on an xy-only mode."""
    I, J, K = 16, 16, 4
    o = Oracle(I, J, K, xlen=16.0, ylen=16.0)
    x = (np.arange(I) + 0.5) * o.dx
    y = (np.arange(J) + 0.5) * o.dy
    mx, my = 2, 3
    f = np.cos(2 * np.pi * mx * x / 16.0)[:, None, None] * np.cos(2 * np.pi * my * y / 16.0)[None, :, None] * np.ones((1, 1, K))
    lam = -4 / o.dx ** 2 * np.sin(np.pi * mx / I) ** 2 - 4 / o.dy ** 2 * np.sin(np.pi * my / J) ** 2
    p = o.poisson_solve(f)
    assert np.abs(p - f / lam).max() < 1e-12


def test_eigenvalue_tables():
    """xrt/yrt as src/modpois.f90:100-107: slot pairs (2m, 2m+1) share -4 dxi^2 sin^2(m pi / n)."""
    o = Oracle(16, 12, 4, xlen=8.0, ylen=3.0)
    xrt, _ = o.metric("xrt")
    m = np.arange(1, 8)
    lam = -4 / o.dx ** 2 * np.sin(np.pi * m / 16) ** 2
    assert xrt[0] == 0.0
    assert np.allclose(xrt[1:15:2], lam, rtol=1e-14) and np.allclose(xrt[2:15:2], lam, rtol=1e-14)
    assert np.isclose(xrt[15], -4 / o.dx ** 2)


@pytest.mark.parametrize("kw", [dict(), dict(lvreman=False, lsmagorinsky=True), dict(lvreman=False, lsmagorinsky=False),
                                dict(BCtopm=2, Uinf=1.0)])
def test_projection_is_divergence_free(kw):
    """After a full substep the velocity divergence is at round-off, like the reference's own restart
    fields (examples/102 warm start: max|div| 8.6e-16)."""
    o = Oracle(24, 20, 16, zf=stretched_zf(16, 8.0, 1.05), **kw)
    o.init_channel()
    assert o.chkdiv()[0] > 1e-2
    o.dt = 0.02
    for _ in range(6):
        o.substep(0.02)
        assert o.chkdiv()[0] < 5e-14
    assert np.isfinite(o.u0).all() and np.abs(o.u0).max() < 2.0


def test_uniform_flow_has_zero_tendency():
    """advection + diffusion of a uniform horizontal flow vanish away from the bottom wall."""
    o = Oracle(12, 10, 8)
    o.u0[...] = 1.0; o.v0[...] = 0.5; o.w0[...] = 0.0
    o.u0[:, :, 0] = 0.0; o.v0[:, :, 0] = 0.0   # k = kb-1 ghost stays 0
    o.advection(); o.subgrid()
    assert np.abs(o.up[1:-1, 1:-1, 1:-1]).max() < 1e-13
    assert np.abs(o.wp[1:-1, 1:-1, 1:-1]).max() < 1e-13
    assert np.abs(o.up[1:-1, 1:-1, 0]).max() > 0     # the zero ghost below produces a wall stress


def test_kappa_limiter_and_scalar_conservation():
    """kappa scheme is in flux form: with periodic x/y and w = 0 at top/bottom the volume integral of
    the scalar tendency vanishes."""
    o = Oracle(16, 12, 10, nsv=2, zf=stretched_zf(10, 5.0, 1.08))
    o.init_channel()
    o.advection()
    dzf, lo = o.metric("dzf")
    w = dzf[1 - lo:10 + 1 - lo]
    for n in range(2):
        tend = o.svp[2:-2, 2:-2, :10, n]
        assert abs((tend * w[None, None, :]).sum()) < 1e-10 * np.abs(tend).sum()


def test_lcg_is_decomposition_independent_recipe():
    """randomize_field (src/modstartup.f90:2367-2396): state=((ir+lin) mod 134456*8121+28411) mod 134456."""
    o = Oracle(6, 5, 4)
    o.randomize("u0", 1.0, 43)
    i, j, k = 3, 2, 4
    lin = i + 6 * (j - 1) + 6 * 5 * (k - 1)
    st = (43 + lin) % 134456
    st = (st * 8121 + 28411) % 134456
    assert o.u0[i, j, k] == pytest.approx((st / 134456 - 0.5) * 2.0, abs=0)


def load_restart_block(target):
    """fill u0, v0, w0 of an Oracle / UdalesGPU-shaped (n+2)^3 array set from the reference restart block"""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_restart102_block.npz"))
    n = int(d["n"])
    out = {}
    for nm in ("u0", "v0", "w0"):
        a = np.zeros((n + 2, n + 2, n + 2), order="F")
        a[:, :, 1:] = d[nm]            # stored levels kb .. kb+n; level kb-1 stays 0
        out[nm] = a
    return n, out, float(d["divmax_global"])


def test_reference_restart_block_is_divergence_free():
    """The only output of the real reference binary in its repository (examples/102 restart files, written after the
    projection by src/modsave.f90:85-99): chkdiv (src/modchecksim.f90:182) on a 24^3 block of it, through the oracle."""
    from oracle.oracle import Oracle
    n, f, dglob = load_restart_block(None)
    o = Oracle(n, n, n, xlen=float(n), ylen=float(n), zf=np.arange(n) + 0.5)       # dx = dy = dz = 1 m as in examples/102
    for nm, a in f.items():
        getattr(o, nm)[...] = a
    divmax, divtot, divrms = o.chkdiv()
    assert np.abs(f["u0"]).max() > 1.0            # a real turbulent field, not zeros
    assert divmax < 2e-15 and divmax <= dglob * 1.0000001
