"""Latency of spfe_search_guided (host pointers in, results out) next to the CPU oracle loop on the same inputs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sp_oracle as O  # noqa: E402
from sp_orb_slam_b200 import SPExtractor, capi, synth  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")
H, W = 480, 752
ex = SPExtractor(800, H, W, WEIGHTS, emit_heat=False, emit_cov=False)
fr = ex.extract(synth.make_frame(H, W, seed=91, n_shapes=400))
rng = np.random.RandomState(0)
for m, r in [(300, 4.0), (1000, 7.0), (4000, 15.0)]:
    src = rng.randint(0, fr["n"], m)
    qdesc = (fr["desc"][src] + 0.02 * rng.randn(m, 256)).astype(np.float32)
    qxy = (fr["kp_xy"][src] + rng.uniform(-3, 3, (m, 2))).astype(np.float32)
    kw = dict(mode=capi.GUIDED_AREA, best_init=256.0, th_le=0.7, th_lt=0.7)
    for _ in range(3):
        got = ex.search_guided(qdesc, qxy, r, fr["occ_grid"], fr["kp_xy"], fr["desc"], **kw)[0]
    t0 = time.perf_counter()
    for _ in range(20):
        ex.search_guided(qdesc, qxy, r, fr["occ_grid"], fr["kp_xy"], fr["desc"], **kw)
    t_gpu = (time.perf_counter() - t0) / 20
    kset = ex.desc_set(1024).from_frame(0, 0)       # the frame's descriptors where the extractor left them
    qset = ex.desc_set(4096).upload(qdesc)          # the local map's descriptors, uploaded once
    for _ in range(3):
        got_s = ex.search_guided(qset, qxy, r, fr["occ_grid"], fr["kp_xy"], kset, **kw)[0]
    t0 = time.perf_counter()
    for _ in range(50):
        ex.search_guided(qset, qxy, r, fr["occ_grid"], fr["kp_xy"], kset, **kw)
    t_set = (time.perf_counter() - t0) / 50
    kset.close(); qset.close()
    t0 = time.perf_counter()
    for _ in range(5):
        ref = O.search_guided(qdesc, qxy, r, fr["occ_grid"], fr["kp_xy"], fr["desc"], mode=0, best_init=256.0, th_le=0.7, th_lt=0.7)[0]
    t_cpu = (time.perf_counter() - t0) / 5
    print(f"m={m:5d} n={fr['n']} r={r:4.1f}: host-pointer call {t_gpu * 1e3:7.3f} ms (incl. H2D of {(m + fr['n']) * 1024 / 1e6:.1f} MB), "
          f"device-resident sets {t_set * 1e3:7.3f} ms, CPU oracle loop {t_cpu * 1e3:7.3f} ms, identical {np.array_equal(got, ref) and np.array_equal(got_s, ref)}, "
          f"matches {(ref >= 0).sum()}", flush=True)
ex.close()
