"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from sp_orb_slam_b200 import sharding


def test_assign_streams_partition():
    for n, w in [(8, 8), (8, 2), (5, 4), (0, 3), (3, 8)]:
        parts = sharding.assign_streams(n, w)
        assert len(parts) == w
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_histogram_bins():
    h = sharding.score_histogram(np.array([0.007, 0.0071, 0.5, 1.0, 2.0]))
    assert h.sum() == 5 and h[0] == 2 and h[-1] == 2
    assert abs(sharding.bin_lower_edge(0) - 0.007) < 1e-12 and abs(sharding.bin_lower_edge(64) - 1.0) < 1e-9


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, _, w = sharding.init_distributed("gloo")
    mine = sharding.assign_streams(5, w)[r]
    frames, ms = sharding.aggregate_throughput(100 * len(mine), 10.0 + 5.0 * r)
    rng = np.random.RandomState(rank)
    scores = np.exp(rng.uniform(np.log(0.007), 0.0, 1000))
    cut = sharding.global_keypoint_budget(scores, 600)
    under = sharding.global_keypoint_budget(scores[:10], 600)
    sharding.barrier()
    q.put((rank, mine, frames, ms, cut, int((scores >= cut).sum()), under))


def test_gloo_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, f0, t0, c0, k0, u0), (r1, m1, f1, t1, c1, k1, u1) = res
    assert m0 == [0, 2, 4] and m1 == [1, 3]
    assert f0 == f1 == 500 and t0 == t1 == 15.0          # SUM of frames, MAX of time
    assert c0 == c1 and c0 > 0.007                        # one common cut on every rank
    assert k0 + k1 <= 600 and k0 + k1 > 450               # budget honoured, not grossly undershot (bin granularity)
    assert u0 == u1 == 0.007                              # under budget -> no cut
