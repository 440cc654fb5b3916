"""Host-side copy ceiling of the box: every rank copies pinned host <-> device buffers on its own GPU at the same time
(D2H alone, H2D alone, both directions), max-over-ranks time, aggregate GB/s.  Run under torchrun at N = 1, 2, 4, 8:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py

The bench's end-to-end figure moves `d2h_bytes_per_step` per step and GPU; this probe says how much the host side of
the box can take when all GPUs copy at once (profiles/r02_d2h_probe.txt)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sp_orb_slam_b200 import sharding  # noqa: E402


def timed(fn, reps):
    # the copies run on side streams: wall clock between two device-wide synchronisations (seconds of copying per call)
    import time
    torch.cuda.synchronize()
    sharding.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    rank, local_rank, world = sharding.init_distributed()
    torch.cuda.set_device(local_rank)
    sharding.bind_to_gpu_numa(local_rank)
    out = {"n_gpus": world}
    for mb in (1, 32):                       # chunk sizes: ~ one frame's outputs, one batch's outputs
        n = mb << 20
        chunks = max(1, (256 << 20) // n)
        host_out = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(chunks)]
        host_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(chunks)]
        dev_out = torch.empty(n, dtype=torch.uint8, device="cuda")
        dev_in = torch.empty(n, dtype=torch.uint8, device="cuda")
        s_out, s_in = torch.cuda.Stream(), torch.cuda.Stream()

        def d2h():
            with torch.cuda.stream(s_out):
                for h in host_out:
                    h.copy_(dev_out, non_blocking=True)

        def h2d():
            with torch.cuda.stream(s_in):
                for h in host_in:
                    dev_in.copy_(h, non_blocking=True)

        def both():
            d2h(); h2d()
        for name, fn in (("d2h", d2h), ("h2d", h2d), ("both", both)):
            fn()
            reps = 24
            ms = timed(fn, reps)
            gb = chunks * n * reps * world / 1e9 * (2 if name == "both" else 1)
            out[f"{name}_{mb}MB_chunks_GBps_aggregate"] = round(gb / (ms * 1e-3), 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
