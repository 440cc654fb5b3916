"""Multi-GPU plumbing: independent camera streams sharded one-per-GPU.

The reference has no distributed code at all (single process, single GPU,
batch 1 -- SURVEY.md §2a).  Frames of different camera streams never interact
inside the extractor and the matcher only pairs frames of one stream, so the
path shards with NO collective on the data path: stream ``s`` lives on rank
``s % world``.  ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in CPU
tests) is used only for (i) the barrier + max-over-ranks timing of the bench and
(ii) the optional *global keypoint budget*: one 64-bin histogram all-reduce
(256 B per rank) that turns a fleet-wide keypoint budget into a common score cut.
"""
from __future__ import annotations

import os
import sys

import numpy as np

N_BINS = 64


def assign_streams(n_streams: int, world_size: int) -> list[list[int]]:
    """Stream ids owned by each rank (round-robin); every stream has exactly one owner."""
    assert n_streams >= 0 and world_size >= 1
    return [list(range(r, n_streams, world_size)) for r in range(world_size)]


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str | None = None):
    """Join the process group described by the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world == 1:
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        # NCCL prints its version banner on stdout when the first communicator is created; bench.py must print exactly
        # one JSON line there, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
            if backend == "nccl":
                t = torch.zeros(1, device=kw["device_id"])
                dist.all_reduce(t)
                torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return rank, local_rank, world


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus += list(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this process to the CPUs that are local to its GPU (sysfs ``local_cpulist`` of the GPU's PCI function), so
    that the page-locked frame / result buffers allocated afterwards are first-touched on the GPU's NUMA node and the
    H2D / D2H DMA of one rank does not cross sockets.  Best effort: returns a description, never raises."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return f"{bdf}: no narrower local CPU list"
        os.sched_setaffinity(0, cpus)
        return f"{bdf}: bound to {len(cpus)} local CPUs"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})"


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def aggregate_throughput(frames_local: int, elapsed_ms_local: float):
    """-> (frames over all ranks, max elapsed ms over ranks).  Whole-job fps = frames / max time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return frames_local, elapsed_ms_local
    dev = _device()
    f = torch.tensor([float(frames_local)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(elapsed_ms_local)], dtype=torch.float64, device=dev)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(f.item())), float(t.item())


def score_histogram(scores: np.ndarray, lo: float = 0.007, hi: float = 1.0) -> np.ndarray:
    """64 log-spaced bins over [lo, hi]; bin 63 holds the highest scores."""
    s = np.clip(np.asarray(scores, np.float64), lo, hi)
    idx = np.minimum((np.log(s / lo) / np.log(hi / lo) * N_BINS).astype(np.int64), N_BINS - 1)
    return np.bincount(idx, minlength=N_BINS).astype(np.int64)


def bin_lower_edge(b: int, lo: float = 0.007, hi: float = 1.0) -> float:
    return float(lo * (hi / lo) ** (b / N_BINS))


def global_keypoint_budget(local_scores: np.ndarray, budget: int, lo: float = 0.007, hi: float = 1.0) -> float:
    """Common score cut such that at most ~``budget`` keypoints survive over ALL ranks.

    One all-reduce (SUM) of a 64-bin int64 histogram.  Returns ``lo`` when the fleet is under budget."""
    import torch
    import torch.distributed as dist
    h = score_histogram(local_scores, lo, hi)
    if dist.is_available() and dist.is_initialized():
        t = torch.from_numpy(h).to(_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h = t.cpu().numpy()
    if h.sum() <= budget:
        return lo
    above = np.cumsum(h[::-1])[::-1]          # above[b] = keypoints in bins >= b
    ok = np.flatnonzero(above <= budget)
    b = int(ok[0]) if len(ok) else N_BINS - 1
    return bin_lower_edge(b, lo, hi)
