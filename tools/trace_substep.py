#!/usr/bin/env python
"""Stage timeline of one RK3 substep per rank (CUDA events on the main / barrier / copy streams, udgpu_trace_dump):
the overlap evidence for the pipelined slab transposes.  Run under torchrun like bench.py.
  torchrun --nproc-per-node N tools/trace_substep.py --grid 512,512,512 --out gpurun_out/trace_n8"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="")
    ap.add_argument("--out", default="gpurun_out/trace")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import udales_b200 as U
    from bench import grid_for, init_state_on_device
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        obj = [U.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]
    I, J, K = (int(x) for x in a.grid.split(",")) if a.grid else grid_for(world, 256)
    g = U.UdalesGPU(I, J, K, xlen=I / 2.0, ylen=J / 2.0, zf=(np.arange(K) + 0.5) * 0.5, device=dev, nprocx=world, myidx=rank, nccl_uid=uid)
    init_state_on_device(g, torch, I // world, J, K, 99 + rank)
    dt = 0.25 * 0.5 / 1.1
    g.dt = dt
    for _ in range(6):
        g.substep(dt)
    g.sync()
    if world > 1:
        dist.barrier()
    g.trace(True)
    for _ in range(2):
        g.substep(dt)
    g.trace_dump(f"{a.out}_{I}x{J}x{K}_r{rank}.txt")
    g.trace(False)
    g.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
