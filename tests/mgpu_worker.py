"""Multi-GPU parity worker (run under torchrun, one rank per GPU): x-slab decomposition vs the
single-pencil CPU oracle on identical global inputs, for every transport variant of the library.
Tolerance = the reference's own decomposition-invariance tolerance 1e-9
(tests/integration/processor_boundaries).  Cases live in tests/slab_parity.py."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ENV_KEYS = ("UDGPU_XMODE", "UDGPU_XCHUNKS", "UDGPU_XSTREAMS", "UDGPU_HALO_NB_BARRIER")
# name -> (cfg.flags, environment)
VARIANTS = {
    "default": (0, {}),                                        # transport chosen by block size (peer stores at these sizes)
    "ce": (0, {"UDGPU_XMODE": "ce", "UDGPU_XCHUNKS": "2"}),    # copy-engine pipeline in k-chunks, p halo carried by the transposes
    "chunks3": (0, {"UDGPU_XMODE": "ce", "UDGPU_XCHUNKS": "3"}),   # uneven k-chunks
    "chunks1": (0, {"UDGPU_XMODE": "ce", "UDGPU_XCHUNKS": "1", "UDGPU_XSTREAMS": "1"}),
    "store": (0, {"UDGPU_XMODE": "store"}),                    # FFT kernels store straight into the peers' windows
    "nccl": (4, {}),                                           # UDGPU_F_NCCL_TRANSPOSE: ncclSend/Recv for transposes and halos
    "allbarrier": (0, {"UDGPU_HALO_NB_BARRIER": "0"}),         # halo exchanges rendezvous with all ranks
    "nolazy": (1, {}),                                         # UDGPU_F_NO_LAZY_FUSION: every call eager (no pipelining with integrate)
}


def main():
    import udales_b200 as U
    from oracle.oracle import Oracle, stretched_zf
    from slab_parity import run_case
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)

    def fresh_uid():
        # one ncclUniqueId per communicator: its bootstrap listener is single-use
        obj = [U.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    shapes = [(64, 64, 32), (128, 64, 24)]
    worst = 0.0
    for name, (flags, env) in VARIANTS.items():
        if quick and name not in ("default", "ce", "nccl"):
            continue
        for k in ENV_KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        cases = [("channel", s) for s in shapes] + [("scalars", (64, 64, 16)), ("ibm", (64, 64, 16)), ("thermo", (64, 64, 16))]
        if name not in ("default", "ce", "nccl"):
            cases = cases[:1] + cases[2:3]
        for kind, shape in cases:
            e = run_case(U, Oracle, kind, shape, world, rank, dev, fresh_uid(), flags=flags, nsub=6 if kind == "channel" else 3,
                         stretched_zf=stretched_zf)
            worst = max(worst, e)
            print(f"rank {rank}: {name} {kind} {shape} ok, worst so far {worst:.2e}", flush=True)
    dist.barrier()
    print(f"MGPU OK rank {rank}/{world} worst abs err {worst:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
