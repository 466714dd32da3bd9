"""Host-side multi-rank logic on CPU (gloo, world_size 2): unique-id broadcast plumbing and the slab
arithmetic used to shard global arrays (mirror of decomp_2d zstart/zsize for nprocy = 1)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import udales_b200 as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_of_partitions_and_overlaps_by_halo():
    I, J, K = 16, 6, 4
    a = np.asfortranarray(np.arange((I + 2) * (J + 2) * (K + 2), dtype=float).reshape((I + 2, J + 2, K + 2), order="F"))
    for P in (1, 2, 4, 8):
        imax = I // P
        parts = [U.slab_of(a, P, r) for r in range(P)]
        for r, s in enumerate(parts):
            assert s.shape == (imax + 2, J + 2, K + 2) and s.flags.f_contiguous
            assert np.array_equal(s[1:-1], a[1 + r * imax:1 + (r + 1) * imax])
            # halo columns are the neighbours' edge columns (periodic ring closes through the global halo)
            assert np.array_equal(s[0], a[r * imax]) and np.array_equal(s[-1], a[(r + 1) * imax + 1])
        assert np.array_equal(np.concatenate([s[1:-1] for s in parts], axis=0), a[1:-1])


def test_gloo_world2_uid_broadcast_and_gather():
    script = textwrap.dedent("""
        import os, sys, numpy as np, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        import udales_b200 as U
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        obj = [bytes(range(128)) if r == 0 else None]          # stands in for ncclGetUniqueId (needs a GPU)
        dist.broadcast_object_list(obj, src=0)
        assert obj[0] == bytes(range(128)) and len(obj[0]) == 128
        I, J, K = 8, 4, 3
        rng = np.random.default_rng(0)
        a = np.asfortranarray(rng.standard_normal((I + 2, J + 2, K + 2)))
        mine = U.slab_of(a, w, r)
        parts = [None] * w
        dist.all_gather_object(parts, mine[1:-1].copy())
        assert np.array_equal(np.concatenate(parts, axis=0), a[1:-1])
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([1.0 + r]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == float(w)
        dist.destroy_process_group()
        print("ok", r)
    """ % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", "-c", script]
    # torchrun has no -c: write the script to a temp file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(script)
        path = f.name
    cmd = cmd[:-2] + [path]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    os.unlink(path)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
