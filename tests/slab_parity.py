"""x-slab (multi-GPU) parity cases against the single-pencil CPU oracle on identical global inputs.

Used by tests/mgpu_worker.py (torchrun, every transport variant) and — outside the timed region — by bench.py, which
prints the verdict as ``"parity": {...}`` in its JSON line so that the driver-run SCALE lines carry it for P = 2, 4, 8.
Tolerance 1e-9 = the reference's own decomposition-invariance tolerance
(tests/integration/processor_boundaries/test_processor_boundaries.py:28); Poisson alone 1e-10.
TEST INFRASTRUCTURE: the oracle is the checker here, never the thing measured."""
import numpy as np

TOL = 1e-9
TOL_POISSON = 1e-10

IBM_BOXES = [(5, 12, 4, 9, 6), (30, 37, 20, 31, 9), (62, 64, 1, 3, 4), (1, 2, 60, 64, 5)]   # (i0,i1,j0,j1,ktop), blocks cross slab edges and the periodic seam


def _slab_lists(lists, world, rank, itot):
    """global point lists -> the local ones of this x-slab (local 1-based i)"""
    imax = itot // world
    lo = rank * imax
    out = {}
    for k, pts in lists.items():
        sel = pts[(pts[:, 0] > lo) & (pts[:, 0] <= lo + imax)].copy()
        sel[:, 0] -= lo
        out[k] = sel.astype(np.int32)
    return out


def run_case(U, Oracle, kind, shape, world, rank, dev, uid, flags=0, nsub=6, stretched_zf=None, failures=None):
    """one case; returns the worst absolute error (Poisson: relative) seen.  kind: channel | scalars | ibm.
    failures: None = assert on the first mismatch; a list = record mismatches and keep going (every rank must run the
    same sequence of collective calls even when one of them sees an error)"""
    def check(cond, info):
        if cond:
            return
        if failures is None:
            raise AssertionError(info)
        failures.append(str(info))
    from helpers import ibm_lists
    I, J, K = shape
    nsv = {"channel": 0, "scalars": 2, "ibm": 1, "thermo": 0}[kind]
    thermo = kind == "thermo"      # temperature + buoyancy + surface heat flux over IBM blocks (slab means through the allreduce)
    zf = stretched_zf(K, K * 0.5, 1.03) if stretched_zf else None
    o = Oracle(I, J, K, zf=zf, nsv=nsv)
    o.init_channel()
    g = U.UdalesGPU(I, J, K, zf=o.zf, device=dev, nprocx=world, myidx=rank, nccl_uid=uid, flags=flags, nsv=nsv, ltempeq=thermo)
    hc = o.ihc
    for n in ("u0", "v0", "w0", "um", "vm", "wm", "pres0"):
        g.push(n, U.slab_of(getattr(o, n), world, rank))
    for n4 in range(nsv):
        g.push("sv0", U.slab_of(o.sv0[..., n4], world, rank, halo=hc), n4)
        g.push("svm", U.slab_of(o.svm[..., n4], world, rank, halo=hc), n4)
    worst = 0.0
    imax = I // world
    if kind in ("ibm", "thermo"):
        lists = ibm_lists(I, J, K, IBM_BOXES)
        o.ibm_set(lists)
        g.ibm_set(_slab_lists(lists, world, rank, I))
        for m in range(4):
            a, b = g.ibm_mask(m), U.slab_of(o.ibm_mask(m), world, rank)
            check(np.array_equal(a, b), ("mask", m))
    if kind == "channel":
        # Poisson alone
        rng = np.random.default_rng(3)
        rhs = rng.standard_normal(shape)
        p_ref = o.poisson_solve(rhs)
        p = g.poisson_solve(np.asfortranarray(rhs[rank * imax:(rank + 1) * imax]))
        e = np.abs(p - p_ref[rank * imax:(rank + 1) * imax]).max() / np.abs(p_ref).max()
        worst = max(worst, e)
        check(e < TOL_POISSON, ("poisson", shape, e))
        # ... and on the resident buffer (the "Poisson solves/s" entry point of the bench)
        g.push("rhs", np.asfortranarray(rhs[rank * imax:(rank + 1) * imax]))
        g.poisson_solve_resident()
        check(np.array_equal(g.pull("rhs"), p), ("poisson_solve_resident differs from poisson_solve", shape))
    if thermo:
        from helpers import add_thermo
        add_thermo(o, g, slab=(U.slab_of, world, rank))
        o.set_bottom(0.01); g.set_bottom(0.01)
        # pressure-gradient forcing + volume-flow correction over the IBM masks: both stay lazily pending through ibmnorm
        prof = -1e-3 * (1.0 + 0.1 * np.arange(K + 1))
        o.set_forcing(prof, 0.1 * prof); g.set_forcing(prof, 0.1 * prof)
        IIu = np.asfortranarray((o.ibm_mask(0)[1:-1, 1:-1, 1:K + 2] != 0.0).astype(np.int32))
        o.set_masscorr(1.0, None, IIu, None); g.set_masscorr(1.0, None)
    dt = 0.02
    o.dt = g.dt = dt
    for s in range(nsub):
        o.substep(dt)
        g.substep(dt)
        if thermo:
            for n in ("thl0", "thlm"):
                a, b = g.pull(n), U.slab_of(getattr(o, n), world, rank)
                e = np.abs(a[:, :, 1:] - b[:, :, 1:]).max() / 288.0
                worst = max(worst, e)
                check(e < TOL, (kind, n, s, shape, e))
            e = np.abs(g.thermo_profile("thvh") - o.thermo_profile("thvh")).max() / 288.0
            worst = max(worst, e)
            check(e < 1e-12, (kind, "thvh", s, e))
        for n in ("u0", "v0", "w0", "um", "vm", "wm"):
            a, b = g.pull(n), U.slab_of(getattr(o, n), world, rank)
            e = np.abs(a - b).max()
            worst = max(worst, e)
            check(e < TOL, (kind, n, s, shape, e))
        # pres0: interior + x-face halo columns + y-face halo rows (what the next advection reads)
        a, b = g.pull("pres0"), U.slab_of(o.pres0, world, rank)
        e = max(np.abs(a[:, 1:-1, 1:-1] - b[:, 1:-1, 1:-1]).max(), np.abs(a[1:-1, :, 1:-1] - b[1:-1, :, 1:-1]).max())
        worst = max(worst, e)
        check(e < TOL, (kind, "pres0", s, shape, e))
        for n4 in range(nsv):
            a, b = g.pull("sv0", n4), U.slab_of(o.sv0[..., n4], world, rank, halo=hc)
            e = np.abs(a[:, :, hc:-hc] - b[:, :, hc:-hc]).max()     # all interior levels incl. the width-hc lateral halos
            worst = max(worst, e)
            check(e < TOL, (kind, "sv0", n4, s, shape, e))
        dmax, dtot, drms = g.divergence()
        omax, otot, orms = o.chkdiv()
        check(drms < 1e-12 and abs(dmax - omax) < 1e-12, (kind, dmax, omax, drms))
    if kind == "channel":
        # adaptive time step: global maxima through the allreduce
        d_ref, _, ct_ref, dn_ref = o.tstep_update(0.05, 0, courant=1.1, diffnr=0.25, dtmax=2.0)
        d, _, ct, dn = g.tstep_update(0.05, 0, courant=1.1, diffnr=0.25, dtmax=2.0)
        check(abs(ct - ct_ref) < 1e-12 * ct_ref and abs(dn - dn_ref) < 1e-12 * dn_ref and abs(d - d_ref) < 1e-12 * d_ref, ('adaptive dt', d, d_ref))
    g.close()
    return worst


def bench_parity(U, Oracle, world, rank, dev, fresh_uid):
    """the bounded set bench.py runs before its timed region: channel (Poisson + 6 substeps), 2 kappa scalars, IBM blocks
    (+1 scalar), temperature with buoyancy over IBM blocks, all on 64x64xK grids split into `world` x-slabs.  Returns the JSON-able verdict."""
    from oracle.oracle import stretched_zf
    cases = [("channel", (64, 64, 32), 6), ("scalars", (64, 64, 16), 3), ("ibm", (64, 64, 16), 3), ("thermo", (64, 64, 16), 3)]
    import os
    if world > 1:
        cases.append(("channel-ce", (64, 64, 32), 3))     # the copy-engine pipeline (default only for large blocks) forced on
    out = {"tol": TOL, "tol_poisson": TOL_POISSON, "max_abs_err": 0.0, "ok": True, "cases": [], "slabs": world,
           "what": "x-slabs vs the single-pencil CPU oracle on the same global input: Poisson solve, then substeps of u0 v0 w0 um vm wm "
                   "pres0 (sv0) on whole slabs incl. halo columns, divergence, adaptive dt"}
    for kind, shape, nsub in cases:
        fails = []
        if kind == "channel-ce":
            os.environ["UDGPU_XMODE"] = "ce"; os.environ["UDGPU_XCHUNKS"] = "2"; kind = "channel"
            ce = True
        else:
            ce = False
        e = run_case(U, Oracle, kind, shape, world, rank, dev, fresh_uid() if world > 1 else None, nsub=nsub, stretched_zf=stretched_zf,
                     failures=fails)
        if ce:
            os.environ.pop("UDGPU_XMODE", None); os.environ.pop("UDGPU_XCHUNKS", None)
        rec = {"case": kind + ("-ce" if ce else ""), "grid": list(shape), "substeps": nsub, "max_abs_err": e}
        if fails:
            out["ok"] = False
            rec["failed"] = fails[:3]
        out["cases"].append(rec)
        out["max_abs_err"] = max(out["max_abs_err"], e)
    return out
