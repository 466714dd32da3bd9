"""Host-side pieces of bench.py that can be checked without a GPU: the synthetic urban-block geometry (BASELINE config 3
style) and its partition over x-slabs, the weak-scaling grids, the channel initial condition."""
import numpy as np

import bench


def test_weak_scaling_grids_keep_cells_per_gpu():
    for n in (1, 2, 4, 8):
        I, J, K = bench.grid_for(n, 256)
        assert I * J * K == n * 256 ** 3 and I % n == 0 and J % n == 0


def test_channel_slab_is_decomposition_independent():
    I, J, K = 16, 8, 6
    u, v, w = bench.channel_slab(I, J, K, 0, I)
    for P in (2, 4):
        imax = I // P
        for r in range(P):
            us, vs, ws = bench.channel_slab(I, J, K, r * imax, imax)
            assert np.array_equal(us[1:-1], u[1 + r * imax:1 + (r + 1) * imax])
            assert np.array_equal(ws[1:-1], w[1 + r * imax:1 + (r + 1) * imax])
    assert np.abs(w[:, :, 1]).max() == 0.0          # w(kb) = 0


def test_urban_blocks_lists_are_consistent_and_partition_over_slabs():
    I, J, K = 64, 32, 16
    whole = bench.urban_blocks(I, J, K, 0, I)
    for nm in "uvwc":
        s, b = whole["solid_" + nm], whole["bound_" + nm]
        assert s.shape[1] == 3 and b.shape[1] == 3 and len(s) > 0 and len(b) > 0
        assert s.min() >= 1 and (s.max(axis=0) <= (I, J, K)).all()
        sset = {tuple(x) for x in s}
        assert not (sset & {tuple(x) for x in b})                 # a boundary point is a fluid point
    # every solid cell makes its two u faces, two v faces and two w faces solid
    cs = {tuple(x) for x in whole["solid_c"]}
    us = {tuple(x) for x in whole["solid_u"]}
    for (i, j, k) in list(cs)[:200]:
        assert (i, j, k) in us and ((i % I) + 1, j, k) in us
    # x-slabs: the local lists, shifted back, are a partition of the global ones
    for P in (2, 4, 8):
        imax = I // P
        for kind in whole:
            parts = []
            for r in range(P):
                loc = bench.urban_blocks(I, J, K, r * imax, imax)[kind].copy()
                assert loc.size == 0 or (loc[:, 0].min() >= 1 and loc[:, 0].max() <= imax)
                loc[:, 0] += r * imax
                parts.append(loc)
            allp = np.concatenate(parts)
            assert len(allp) == len(whole[kind])
            assert {tuple(x) for x in allp} == {tuple(x) for x in whole[kind]}
