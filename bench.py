#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native uDALES dynamics core.

Metric (BASELINE.json): cell-updates/s = grid cells advanced through one RK3 substep
(tstep_update, advection, subgrid, poisson, tstep_integrate, halos, boundary — one pass of
src/program.f90:132-207 restricted to the in-scope calls) per second, whole job.
Workload at N=1: BASELINE config 2, neutral periodic channel 256^3, stencil + Poisson only.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA through the C-ABI)
  python bench.py --impl reference --gpus N ...            # CPU arm: the oracle port on host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

B_PER_CELL = {  # algorithmic (compulsory) fp64 bytes per cell, SURVEY.md §8d / DESIGN.md
    "mom_tend": 64.0, "closure": 40.0, "poisson_core": 80.0, "fillps": 56.0, "tderive_integrate": 96.0, "halos": 0.0,
}
PROF_NAMES = ["mom_tend", "closure", "poisson_core", "fillps", "tderive_integrate", "halos", "poisson_inverse+tderive_integrate (pipelined)"]
# DRAM traffic per launch group (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch of the family) of the
# 256^3 substep: regenerated from an `ncu --set full` capture of the current kernels by tools/ncu_traffic.py, which
# writes profiles/ncu_traffic_256.json together with the commit it was captured at.  Absent file -> traffic is null.
def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic_256.json")
    try:
        return json.load(open(p))
    except Exception:
        return None

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, dev=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.dev = dev

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(self.dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def grid_for(n_gpus, base):
    """weak scaling: 256^3 cells per GPU.  1: 256^3 (BASELINE config 2), 2: 512x256x256, 4: 512x512x256, 8: 512^3."""
    b = base
    return {1: (b, b, b), 2: (2 * b, b, b), 4: (2 * b, 2 * b, b), 8: (2 * b, 2 * b, 2 * b)}[n_gpus]


def channel_slab(I, J, K, i_lo, imax, seed=0):
    """synthetic periodic channel of SURVEY.md §8d on the x-slab [i_lo, i_lo+imax): u = 1 + noise, v, w noise.
    Seeded per global i-plane, so the field does not depend on the decomposition."""
    shp = (imax + 2, J + 2, K + 2)
    u = np.zeros(shp, order="F"); v = np.zeros(shp, order="F"); w = np.zeros(shp, order="F")
    for il in range(imax):
        rng = np.random.default_rng([seed, i_lo + il])
        u[il + 1, 1:-1, 1:-1] = 1.0 + 0.05 * (rng.random((J, K)) - 0.5)
        v[il + 1, 1:-1, 1:-1] = 0.05 * (rng.random((J, K)) - 0.5)
        w[il + 1, 1:-1, 2:-1] = 0.05 * (rng.random((J, K - 1)) - 0.5)
    return u, v, w


def urban_blocks(I, J, K, i_lo, imax):
    """BASELINE config 3 in synthetic form: the 64x64 tile of examples/102 (16x16 blocks of height 8 on a regular
    grid) repeated over the domain.  Returns the eight LOCAL point lists of src/modibm.f90 for the
    x-slab [i_lo, i_lo+imax): solid_* = masked points, bound_* = fluid points next to a masked point."""
    sc = np.zeros((I + 2, J + 2, K + 2), dtype=bool)
    hb = min(8, max(1, K // 4))
    for ti in range(0, I, 32):
        for tj in range(0, J, 32):
            sc[1 + ti + 8:1 + ti + 24, 1 + tj + 8:1 + tj + 24, 1:hb + 1] = True
    su = sc | np.roll(sc, 1, axis=0); sv = sc | np.roll(sc, 1, axis=1)
    sw = sc.copy(); sw[:, :, 1:] |= sc[:, :, :-1]
    dirs = {"u": ((0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)), "v": ((1, 0, 0), (-1, 0, 0), (0, 0, 1), (0, 0, -1)),
            "w": ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)),
            "c": ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))}
    lists = {}
    for nm, sol in (("u", su), ("v", sv), ("w", sw), ("c", sc)):
        m = np.ones((I + 2, J + 2, K + 2)); m[:, :, 0] = 0.0
        if nm == "w":
            m[:, :, 1] = 0.0
        m[1:I + 1, 1:J + 1, 1:K + 1][sol[1:I + 1, 1:J + 1, 1:K + 1]] = 0.0
        m[0] = m[I]; m[I + 1] = m[1]; m[:, 0] = m[:, J]; m[:, J + 1] = m[:, 1]
        fluid = m[1:I + 1, 1:J + 1, 1:K + 1] == 1.0
        touch = np.zeros_like(fluid)
        for a, b, c in dirs[nm]:
            touch |= m[1 + a:I + 1 + a, 1 + b:J + 1 + b, 1 + c:K + 1 + c] == 0.0
        for kind, sel in (("solid_", sol[1:I + 1, 1:J + 1, 1:K + 1]), ("bound_", fluid & touch)):
            pts = np.argwhere(sel) + 1
            if kind == "bound_" and nm == "w":
                pts = pts[pts[:, 2] >= 2]
            pts = pts[(pts[:, 0] > i_lo) & (pts[:, 0] <= i_lo + imax)]
            pts[:, 0] -= i_lo
            lists[kind + nm] = pts.astype(np.int32)
    return lists


def workload_config(args, world):
    """the `config` object: identical in both arms (ours / --impl reference) for the same command line"""
    I, J, K = grid_for(world, args.size)
    if args.grid:
        I, J, K = (int(x) for x in args.grid.split(","))
    what = {"channel": f"neutral periodic channel {I}x{J}x{K}, stencil+Poisson only, no IBM/scalars",
            "scalars": f"periodic channel {I}x{J}x{K} + 4 kappa-advected passive scalars",
            "ibm": f"periodic channel {I}x{J}x{K} + urban blocks (16x16x8 per 32x32 tile) masked by the IBM path",
            "thermo": f"periodic channel {I}x{J}x{K} + urban blocks (IBM) + temperature equation with buoyancy, surface heat flux, bottom wall "
                      "function, volume-flow forcing (the physics set of examples/102)",
            "poisson": f"Poisson solve only, {I}x{J}x{K}, rhs resident"}[args.workload]
    what += " (BASELINE config 2)" if (world == 1 and args.workload == "channel" and (I, J, K) == (256, 256, 256)) else \
            f" ({I * J * K // world} cells per GPU, x-slabs nprocx={world})"
    return {"workload": what, "grid": [I, J, K], "substeps_per_step": 1,
            "l2": "per-GPU working set (13 fields x 134 MB at 256^3) >> 126 MB L2, no flush needed",
            "sgs": "vreman", "poisson": "FFT2D x,y + tridiagonal z", "decomposition": f"nprocx={world}, nprocy=1",
            "ladaptive": False}


def native_oracle():
    """the CPU arm's own build of the oracle: -O3 -march=native (FMA contraction allowed), compiled ON THE BOX that
    runs it (the parity build oracle/liboracle.so stays -ffp-contract=off and generic x86-64)."""
    out = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "liboracle_native.so")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    flags = ["-O3", "-march=native", "-fopenmp", "-fPIC", "-std=gnu11", "-Wno-unused-variable"]
    try:
        subprocess.check_call([cc] + flags + ["-shared", "-o", so, os.path.join(ROOT, "oracle", "udales_oracle.c"), "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return so, " ".join(flags)
    except Exception:
        return None, "-O3 -ffp-contract=off (parity build; the native build failed)"


def cpu_arm(n, nsub, dt):
    """time `nsub` RK3 substeps of an n^3 block of the channel with the oracle port on all host cores"""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)     # torchrun sets it to 1
    so, flags = native_oracle()
    import oracle.oracle as OM
    if so:
        OM.use_library(so)
    o = OM.Oracle(n, n, n)
    o.init_channel()
    o.dt = dt
    for _ in range(2):
        o.substep(dt)
    t0 = time.perf_counter()
    for _ in range(nsub):
        o.substep(dt)
    el = time.perf_counter() - t0
    return n ** 3 * nsub / el, el, cores, flags


# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle (C/OpenMP restatement of the reference loops — the Fortran/MPI binary
    cannot be built here: no Fortran compiler, MPI or FFTW).  Rank 0 only."""
    if rank != 0:
        return
    n = min(args.size, 256)
    dt = 0.25 * 0.5 / 1.1
    val, el, cores, flags = cpu_arm(n, args.steps, dt)
    line = {
        "impl": "reference", "metric": "cell-updates/s", "value": val, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": 2, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "cflags": flags,
                         "sample": f"{args.steps} RK3 substeps on a {n}^3 block of the workload (the whole workload at N=1; cell-updates/s "
                                   "is intensive), C/OpenMP restatement of the reference loops on all host cores of ONE box "
                                   "(not the Fortran/MPI binary: no Fortran/MPI/FFTW in the image; in-tree radix-2 real FFT (half-length complex transform + split) instead of FFTW)"},
        "e2e": {"value": val, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def _cudart():
    import ctypes as C
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    return rt


def init_state_on_device(g, torch, imax, J, K, seed):
    """the synthetic channel for grids too large to stage through host memory (1024^3: 8.6 GB per field): u = 1 + noise,
    v, w = noise, generated by torch on the device level block by level block and copied into the library's arrays
    (udgpu_device_ptr); halos()/boundary() then fill the halo cells and um, vm, wm become copies of u0, v0, w0."""
    import ctypes as C
    rt = _cudart()
    ptr = {}
    for nm in ("u0", "v0", "w0", "um", "vm", "wm"):
        d = C.c_void_p()
        g._chk(g.L.udgpu_device_ptr(g.h, U_FIELD_IDS[nm], 0, C.byref(d)))
        ptr[nm] = d.value
    g.sync()
    pi, pj = imax + 2, J + 2
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    blk = max(1, min(K, (256 << 20) // (pi * pj * 8)))
    for nm, base, lo in (("u0", 1.0, 1), ("v0", 0.0, 1), ("w0", 0.0, 2)):
        for k0 in range(lo, K + 1, blk):       # storage level = Fortran k (kh = 1); w(kb) stays 0
            nl = min(blk, K + 1 - k0)
            t = torch.zeros((nl, pj, pi), dtype=torch.float64, device="cuda")
            t[:, 1:-1, 1:-1] = base + 0.05 * (torch.rand((nl, J, imax), dtype=torch.float64, device="cuda", generator=gen) - 0.5)
            torch.cuda.synchronize()
            assert rt.cudaMemcpy(ptr[nm] + k0 * pi * pj * 8, t.data_ptr(), nl * pi * pj * 8, 3) == 0
            del t
    g.halos(); g.boundary(); g.sync()
    n = pi * pj * (K + 2) * 8
    for a, b in (("um", "u0"), ("vm", "v0"), ("wm", "w0")):
        assert rt.cudaMemcpy(ptr[a], ptr[b], n, 3) == 0
    torch.cuda.synchronize()


U_FIELD_IDS = {"u0": 0, "v0": 1, "w0": 2, "um": 3, "vm": 4, "wm": 5}


def run_ours(args, rank, world):
    # NCCL prints its version banner on stdout: keep stdout clean for the ONE JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import udales_b200 as U
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def fresh_uid():
        if world == 1:
            return None
        obj = [U.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def max_over_ranks(x, op="max"):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        return float(t.item())

    # ---- parity of this very build and decomposition against the CPU oracle, OUTSIDE every timed region ------------
    parity = None
    if not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from slab_parity import bench_parity
        from oracle.oracle import Oracle
        parity = bench_parity(U, Oracle, world, rank, dev, fresh_uid)
        parity["max_abs_err"] = max_over_ranks(parity["max_abs_err"])
        parity["ok"] = bool(max_over_ranks(1.0 if parity["ok"] else 0.0, "min") > 0.5)

    cfg = workload_config(args, world)
    I, J, K = cfg["grid"]
    imax = I // world
    nsv = 4 if args.workload == "scalars" else 0
    dt = 0.25 * 0.5 / 1.1
    hbm, how = peaks()

    def make(I, J, K, nsv=0, on_device=False, workload="channel"):
        imax = I // world
        g = U.UdalesGPU(I, J, K, xlen=I / 2.0, ylen=J / 2.0, zf=(np.arange(K) + 0.5) * 0.5, device=dev,
                        nprocx=world, myidx=rank, nccl_uid=fresh_uid(), nsv=nsv, ltempeq=(workload == "thermo"))
        if on_device:
            init_state_on_device(g, torch, imax, J, K, 1234 + rank)
        else:
            u, v, w = channel_slab(I, J, K, rank * imax, imax)
            for nm, f in (("u0", u), ("v0", v), ("w0", w)):
                g.push(nm, f)
            if nsv:
                rng = np.random.default_rng([7, rank])
                for n4 in range(nsv):
                    sf = np.asfortranarray(1.0 + 0.1 * rng.random(g.shape("sv0")))
                    g.push("sv0", sf, n4)
            g.halos(); g.boundary()
            for nm in ("u0", "v0", "w0"):
                g.push(nm.replace("0", "m"), g.pull(nm))
            for n4 in range(nsv):
                g.push("svm", g.pull("sv0", n4), n4)
        if workload in ("ibm", "thermo"):
            g.ibm_set(urban_blocks(I, J, K, rank * imax, imax))
        if workload == "thermo":
            # examples/102: ltempeq, lbuoyancy, wtsurf = 0.01, wttop = -0.01, luvolflowr; stably stratified thl0 + noise
            g.set_thermo(lbuoyancy=True, thls=288.0, BCtopT=1, wttop=-0.01, BCbotT=1, wtsurf=0.01)
            g.set_bottom(0.01); g.set_masscorr(1.0, None)
            rng = np.random.default_rng([11, rank])
            th = np.zeros(g.shape("thl0"), order="F")
            th[...] = 288.0 + 0.01 * np.arange(K + 2)[None, None, :] + 0.05 * rng.random(th.shape)
            ek = np.full(g.shape("ekh"), 1.5e-5 / 0.71, order="F")
            g.push("ekh", ek); g.push("thl0", th); g.halos(); g.boundary()
            g.push("thlm", g.pull("thl0")); g.thermodynamics()
        g.dt = dt
        return g

    def barrier(g):
        g.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(g, fn, steps, warmup):
        """W untimed calls, then exactly `steps` timed ones between barriers, CUDA events on the library stream, max over ranks"""
        st = torch.cuda.ExternalStream(g.stream(), device=dev)
        for _ in range(warmup):
            fn()
        barrier(g)
        l0 = g.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        barrier(g)
        return max_over_ranks(e0.elapsed_time(e1)), g.launch_count() - l0

    big = I * J * K // world > 300 ** 3      # host staging of > 27 M cells per rank is slow: generate on the device
    g = make(I, J, K, nsv=nsv, on_device=big and not nsv, workload=args.workload)
    ncell = I * J * K            # whole job
    ncell_loc = imax * J * K
    W = max(args.warmup, 3)

    if args.workload == "poisson":
        # BASELINE config 4 style: the Poisson solve alone on the resident right-hand side (udgpu_poisson_solve_resident)
        rng = np.random.default_rng([3, rank])
        g.push("rhs", np.asfortranarray(rng.standard_normal(g.shape("rhs"))))
        sampler = ClockSampler(dev); sampler.start()
        ms, launches = timed(g, g.poisson_solve_resident, args.steps, W)
        ach = 80.0 * ncell_loc * args.steps / (ms * 1e-3) / 1e9
        line = {"metric": "poisson-solves/s", "value": args.steps / (ms * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": args.steps,
                "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cells_per_s": ncell * args.steps / (ms * 1e-3),
                "roofline": {"kernel": "poisson_core", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                             "traffic": None, "peak_source": how, "bytes_per_cell": 80.0},
                "gpu_launches": launches, "clocks": sampler.stop(), "parity": parity}
        g.close()
        sys.stdout.flush(); os.dup2(real_stdout, 1)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(dev); sampler.start()
    ms, launches = timed(g, lambda: g.substep(dt), args.steps, W)
    clocks = sampler.stop()
    value = ncell * args.steps / (ms * 1e-3)
    dmax, dtot, drms = g.divergence()

    # ---- per-kernel-family device times (CUDA events on the library stream, live) ----
    g.profile_enable(True); g.profile_reset()
    nprof = min(args.steps, 10)
    for _ in range(nprof):
        g.substep(dt)
    fam = {}
    for i, nm in enumerate(PROF_NAMES):
        t, cnt = g.profile_get(i)
        fam[nm] = t / nprof
    g.profile_enable(False)
    roof_all = {}
    bpc = dict(B_PER_CELL)
    if world > 1 and fam.get(PROF_NAMES[6], 0) > 0:
        # copy-engine pipeline (UDGPU_XMODE=ce): "poisson_core" is the forward half + z solve (fillps fused: 48 B read + 8 B
        # written by the first transform, 32 B for the x transform and the z solve), the inverse half runs chunk by chunk
        # together with tderive+integrate in its own slot
        bpc["poisson_core"] = 56.0 + 32.0
        bpc[PROF_NAMES[6]] = 32.0 + 96.0
    elif world > 1 and fam.get("fillps", 0) < 0.5 * 0.17 * ncell_loc / 256 ** 3:
        # slab solve with fillps fused into the forward y transform: the rhs array is neither written nor read (16 B/cell
        # less than fillps 56 + solve 80); what is left in the fillps slot is the slab exchange of up
        bpc["poisson_core"] = 56.0 + 80.0 - 16.0
        bpc["fillps"] = 0.0
    if nsv:   # K3: (4 + 2 n) * 8 B/cell for n fields in one pass; scalar integrate: svm, svp -> sv0 = 24 B/cell/field
        bpc["mom_tend"] += (4 + 2 * nsv) * 8.0
        bpc["tderive_integrate"] += 24.0 * nsv
    for nm, t in fam.items():
        if t > 0 and bpc.get(nm, 0) > 0:
            ach = bpc[nm] * ncell_loc / (t * 1e-3) / 1e9
            roof_all[nm] = {"ms": t, "achieved_gbs": ach, "frac": ach / hbm, "bytes_per_cell": bpc[nm]}
        elif t > 0:
            roof_all[nm] = {"ms": t}
    cand = {k: v for k, v in roof_all.items() if "frac" in v}
    dom = max(cand, key=lambda k: cand[k]["ms"])
    traffic = ncu_traffic() if (world == 1 and (I, J, K) == (256, 256, 256) and args.workload == "channel") else None
    roofline = {"kernel": dom, "bound": "hbm", "achieved": cand[dom]["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                "frac": cand[dom]["frac"],
                "traffic": (traffic or {}).get("families", {}).get(dom), "traffic_source": (traffic or {}).get("source"),
                "peak_source": how, "bytes_per_cell": bpc[dom], "families": roof_all,
                "substep": {"bytes_per_cell": sum(B_PER_CELL.values()), "achieved_gbs": sum(B_PER_CELL.values()) * ncell_loc / (ms / args.steps * 1e-3) / 1e9,
                            "frac": sum(B_PER_CELL.values()) * ncell_loc / (ms / args.steps * 1e-3) / 1e9 / hbm} if not nsv else None}

    # ---- the same with the adaptive time step of examples/001/999 (ladaptive = .true.): k_cfl + allreduce + one host
    #      synchronisation per RK3 time step inside the timed region (src/modtstep.f90:69-144) ----
    nad = max(6, min(args.steps, 30) // 3 * 3)
    ms_ad, _ = timed(g, lambda: g.substep(dt, ladaptive=True, courant=1e9, diffnr=1e9), nad, 3)
    adaptive = {"ms_per_step": ms_ad / nad, "value": ncell * nad / (ms_ad * 1e-3), "steps": nad,
                "note": "ladaptive=.true.: maxima kernel + allreduce + host sync once per RK3 time step; dtmax keeps dt fixed so the flow is the same"}

    # ---- end to end through the C-ABI with HOST buffers --------------------------------------------------
    # The prognostic state lives in pinned host arrays (a host-resident model).  One call of
    # udgpu_rk3_step_host = one RK3 time step = three passes of the hot path: H2D of u0,v0,w0,pres0
    # (um = u0 at the start of a time step, src/modtstep.f90:330-338), 3 substeps, D2H of the same four.
    names_io = ("u0", "v0", "w0", "pres0")
    nloc = (imax + 2) * (J + 2) * (K + 2)
    e2e = None
    if not big:
        host = {nm: torch.empty(nloc, dtype=torch.float64).pin_memory() for nm in names_io}
        while g.rk3step != 3:   # finish the running time step first so that um == u0
            g.substep(dt)
        for nm in names_io:
            g.pull_raw(nm, host[nm].data_ptr())
        g.sync()
        ne2e = max(3, min(args.steps // 3, 10))
        ptrs = [host[nm].data_ptr() for nm in names_io]
        for it in range(2 + ne2e):
            if it == 2:
                barrier(g); t0 = time.perf_counter()
            g.rk3_step_host(*ptrs, dtmax=dt)
        barrier(g)
        t_e2e = max_over_ranks((time.perf_counter() - t0) / ne2e) / 3.0      # per substep (= per step of this bench)
        bi = len(names_io) * nloc * 8 * world / 3.0
        e2e = {"value": ncell / t_e2e, "unit": "cell-updates/s", "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bi,
               "ms_per_step": 1e3 * t_e2e, "substeps_per_call": 3,
               "note": "udgpu_rk3_step_host: u0,v0,w0,pres0 pushed from / pulled to pinned host arrays once per RK3 time step "
                       "(3 substeps); bytes and ms are per substep"}
        del host
    g.close()
    del g
    torch.cuda.empty_cache()

    # ---- north_star grid: 1024^3 periodic channel, the SAME global grid at every N (strong scaling) -------------------
    ns = None
    if args.workload == "channel" and not args.grid and not args.no_1024:
        n1 = args.size1024
        g1 = make(n1, n1, n1, on_device=True)
        k1 = max(3, min(args.steps // 10, 9) // 3 * 3)
        ms1, l1 = timed(g1, lambda: g1.substep(dt), k1, 3)
        g1.profile_enable(True); g1.profile_reset()
        for _ in range(3):
            g1.substep(dt)
        fam1 = {nm: g1.profile_get(i)[0] / 3 for i, nm in enumerate(PROF_NAMES)}
        d1 = g1.divergence()
        ns = {"grid": [n1, n1, n1], "scaling": "strong", "value": n1 ** 3 * k1 / (ms1 * 1e-3), "unit": "cell-updates/s",
              "ms_per_step": ms1 / k1, "steps": k1, "warmup": 3, "gpu_launches": l1, "divergence_rms": d1[2],
              "families_ms": fam1, "substep_frac_of_hbm": sum(B_PER_CELL.values()) * (n1 ** 3 // world) / (ms1 / k1 * 1e-3) / 1e9 / hbm,
              "note": "north_star asks >= 6x from 1 to 8 GPUs on this grid: compare this object's value across the N = 1, 2, 4, 8 lines"}
        g1.close()
        del g1

    # ---- CPU baseline beside it (oracle port, bounded sample) ----
    cpu = None
    if not args.no_cpu and world == 1:
        n = min(args.size, 256)
        nsub = 30          # ~10 s of host work at 256^3 on 16 cores
        val, el, cores, flags = cpu_arm(n, nsub, dt)
        cpu = {"value": val, "unit": "cell-updates/s", "cores": cores, "kind": "port", "cflags": flags,
               "sample": f"{nsub} RK3 substeps of the same {n}^3 workload; C/OpenMP restatement of the reference loops, "
                         "in-tree radix-2 real FFT (half-length complex transform + split; FFTW absent) — not the Fortran/MPI binary"}

    line = {
        "metric": "cell-updates/s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": launches, "clocks": clocks,
        "poisson_solves_per_s": (1e3 / fam["poisson_core"]) if fam.get("poisson_core") and world == 1 else None,
        "divergence_rms": drms, "parity": parity, "adaptive": adaptive, "grid1024": ns,
    }
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the small-grid oracle comparison that precedes the timed region")
    ap.add_argument("--no-1024", action="store_true", help="skip the north_star 1024^3 strong-scaling measurement")
    ap.add_argument("--size1024", type=int, default=1024, help=argparse.SUPPRESS)
    ap.add_argument("--grid", default="", help="explicit global grid I,J,K (e.g. 1024,1024,512 = BASELINE config 4) instead of the weak-scaling grid")
    ap.add_argument("--workload", default="channel", choices=["channel", "scalars", "ibm", "thermo", "poisson"],
                    help="channel = BASELINE config 2 (the headline); scalars = + 4 kappa scalars (config 5 style); "
                         "ibm = + urban blocks masked on the device (config 3 style); poisson = the solve alone (config 4 style)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
