#!/usr/bin/env python
"""A/B timing of kernel variants on one GPU (run under gpurun): one line per variant with the whole-substep
time and the per-family device times (CUDA events inside the library).  Variants are selected by the
library's UDGPU_* environment switches / init flags, so one call measures all of them back to back.

  python tools/ab_variants.py [--size 256] [--steps 30] name[:ENV=V[,ENV=V]][:flags=N] ...
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

FAM = ["mom_tend", "closure", "poisson_core", "fillps", "tderive_integrate", "halos"]


def run(name, env, flags, size, steps, nsv=0):
    import udales_b200 as U
    from bench import channel_slab
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        n = size
        g = U.UdalesGPU(n, n, n, xlen=n / 2.0, ylen=n / 2.0, zf=(np.arange(n) + 0.5) * 0.5, flags=flags, nsv=nsv)
        u, v, w = channel_slab(n, n, n, 0, n)
        for nm, f in (("u0", u), ("v0", v), ("w0", w)):
            g.push(nm, f)
        g.halos(); g.boundary()
        for nm in ("u0", "v0", "w0"):
            g.push(nm.replace("0", "m"), g.pull(nm))
        if nsv:
            rng = np.random.default_rng(1)
            for n4 in range(nsv):
                s = np.asfortranarray(1.0 + 0.1 * rng.random(g.shape("sv0")))
                g.push("sv0", s, n4); g.push("svm", s, n4)
            g.halos(); g.boundary()
        dt = 0.25 * 0.5 / 1.1
        g.dt = dt
        for _ in range(6):
            g.substep(dt)
        g.sync()
        l0 = g.launch_count()
        t0 = time.perf_counter()
        for _ in range(steps):
            g.substep(dt)
        g.sync()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        nl = (g.launch_count() - l0) / steps
        g.profile_enable(True); g.profile_reset()
        for _ in range(9):
            g.substep(dt)
        fam = {nm: g.profile_get(i)[0] / 9 for i, nm in enumerate(FAM)}
        g.profile_enable(False)
        drms = g.divergence()[2]
        g.close()
        out = {"variant": name, "ms_per_substep": round(ms, 4), "Gcell_s": round(n ** 3 / ms / 1e6, 3), "launches": nl,
               "fam_ms": {k: round(v, 4) for k, v in fam.items()}, "div_rms": drms}
        print(json.dumps(out), flush=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--nsv", type=int, default=0)
    ap.add_argument("variants", nargs="*", default=["base"])
    a = ap.parse_args()
    for spec in a.variants:
        parts = spec.split(":")
        env, flags = {}, 0
        for p in parts[1:]:
            for kv in p.split(","):
                k, v = kv.split("=")
                if k == "flags":
                    flags = int(v)
                else:
                    env[k] = v
        try:
            run(parts[0], env, flags, a.size, a.steps, a.nsv)
        except Exception as e:  # keep going: one broken variant must not hide the others
            print(json.dumps({"variant": parts[0], "error": repr(e)}), flush=True)


if __name__ == "__main__":
    main()
