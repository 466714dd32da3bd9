python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -x -k "adaptive or tstep_update or golden or rk3_step_host" 2>&1 | tail -3
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import udales_b200 as U
from bench import channel_slab
n = 256
g = U.UdalesGPU(n, n, n, xlen=n / 2.0, ylen=n / 2.0, zf=(np.arange(n) + 0.5) * 0.5)
u, v, w = channel_slab(n, n, n, 0, n)
for nm, f in (("u0", u), ("v0", v), ("w0", w)):
    g.push(nm, f)
g.halos(); g.boundary()
for nm in ("u0", "v0", "w0"):
    g.push(nm.replace("0", "m"), g.pull(nm))
dt = 0.25 * 0.5 / 1.1
g.dt = dt
for lad in (False, True, False, True):
    for _ in range(6):
        g.substep(dt, ladaptive=lad, courant=1e9, diffnr=1e9)
    g.sync(); t0 = time.perf_counter()
    for _ in range(60):
        g.substep(dt, ladaptive=lad, courant=1e9, diffnr=1e9)
    g.sync(); ms = 1e3 * (time.perf_counter() - t0) / 60
    print("ladaptive", lad, "ms/substep %.4f" % ms, "G cell-updates/s %.3f" % (n ** 3 / ms / 1e6))
PY
