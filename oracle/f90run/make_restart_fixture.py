#!/usr/bin/env python
"""tests/golden/ref_restart102_block.npz: a block of the ONLY real output of the reference binary that ships with it —
the restart files examples/102/warmstart_files/initd00000267_{000,001}_{000,001}.102 (64^3, 2x2 ranks, written by
src/modsave.f90:85-99 after the projection).  Run in the build container only (needs /root/reference).

Pinned property: the projected field is divergence free to round-off on the staggered grid,
div = (u(i+1)-u(i))/dx + (v(j+1)-v(j))/dy + (w(k+1)-w(k))/dzf, dx = dy = dz = 1 m (src/modchecksim.f90:182).
The block keeps one halo column / row and level k+1 so that the formula can be evaluated on 24^3 cells, through the
oracle's chkdiv and through udgpu_divergence, without periodic wrap."""
import os
import struct
import sys

import numpy as np

REF = os.environ.get("UDALES_REFERENCE", "/root/reference")
D = os.path.join(REF, "examples", "102", "warmstart_files")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "ref_restart102_block.npz")


def records(path):
    raw = open(path, "rb").read()
    pos, out = 0, []
    while pos < len(raw):
        n = struct.unpack("<i", raw[pos:pos + 4])[0]
        out.append(raw[pos + 4:pos + 4 + n])
        assert struct.unpack("<i", raw[pos + 4 + n:pos + 8 + n])[0] == n
        pos += 8 + n
    return out


def main():
    nx = ny = 2
    I = J = K = 64
    loc = 32
    glob = {nm: np.zeros((I + 2, J + 2, K + 1)) for nm in ("u0", "v0", "w0", "pres0", "thl0")}   # (0:I+1, 0:J+1, 1:K+1)
    meta = None
    for px in range(nx):
        for py in range(ny):
            rec = records(os.path.join(D, f"initd00000267_{px:03d}_{py:03d}.102"))
            assert len(rec[0]) == loc * loc * K * 8 and len(rec[1]) == loc * loc * K * 5 * 4
            for q, nm in enumerate(("u0", "v0", "w0", "pres0", "thl0")):
                a = np.frombuffer(rec[2 + q], dtype="<f8").reshape((loc + 2, loc + 2, K + 1), order="F")
                # interior of the rank; rank halos only where they are the global (periodic) halo
                glob[nm][1 + px * loc:1 + (px + 1) * loc, 1 + py * loc:1 + (py + 1) * loc, :] = a[1:-1, 1:-1, :]
                if px == 0: glob[nm][0, 1 + py * loc:1 + (py + 1) * loc, :] = a[0, 1:-1, :]
                if px == nx - 1: glob[nm][I + 1, 1 + py * loc:1 + (py + 1) * loc, :] = a[-1, 1:-1, :]
                if py == 0: glob[nm][1 + px * loc:1 + (px + 1) * loc, 0, :] = a[1:-1, 0, :]
                if py == ny - 1: glob[nm][1 + px * loc:1 + (px + 1) * loc, J + 1, :] = a[1:-1, -1, :]
            meta = struct.unpack("<2d", rec[-1])
    u, v, w = glob["u0"], glob["v0"], glob["w0"]
    div = (u[2:, 1:-1, :-1] - u[1:-1, 1:-1, :-1]) + (v[1:-1, 2:, :-1] - v[1:-1, 1:-1, :-1]) + (w[1:-1, 1:-1, 1:] - w[1:-1, 1:-1, :-1])
    print("global: max|div| %.3e rms %.3e  |u|max %.3f  timee %.4f dt %.5f" % (np.abs(div).max(), np.sqrt((div ** 2).mean()), np.abs(u).max(), *meta))
    # periodic consistency of the halos written by the reference (halos ran before the dump)
    assert np.array_equal(u[0, 1:-1], u[I, 1:-1]) and np.array_equal(u[I + 1, 1:-1], u[1, 1:-1])
    i0, j0, n = 9, 17, 24
    blk = {nm: np.ascontiguousarray(glob[nm][i0 - 1:i0 + n + 1, j0 - 1:j0 + n + 1, 0:n + 1]) for nm in ("u0", "v0", "w0")}
    np.savez_compressed(OUT, i0=i0, j0=j0, n=n, timee=meta[0], dt=meta[1], divmax_global=np.abs(div).max(), **blk)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")
    # a 32^3 block (with its neighbour columns / rows and level k+1) of the real turbulent state incl. pres0: input of a
    # closure + substep parity run on real LES data instead of synthetic noise (tests/test_gpu_parity.py)
    m = 32
    i1, j1 = 17, 9
    turb = {nm: np.ascontiguousarray(glob[nm][i1 - 1:i1 + m + 1, j1 - 1:j1 + m + 1, 0:m + 1]) for nm in ("u0", "v0", "w0", "pres0", "thl0")}
    out2 = OUT.replace("ref_restart102_block", "ref_restart102_turb32")
    np.savez_compressed(out2, i0=i1, j0=j1, n=m, timee=meta[0], dt=meta[1], **turb)
    print("wrote", out2, os.path.getsize(out2) // 1024, "KiB")


if __name__ == "__main__":
    main()
