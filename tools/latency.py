"""Blocking single-frame latency of the reference-style call (spfe_extract == SPExtractor::operator()) and throughput at
the other BASELINE geometries.  python tools/latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sp_orb_slam_b200 import SPExtractor, synth  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")

for H, W, nf in [(480, 752, 800), (480, 640, 800), (1080, 1920, 2000)]:
    ex = SPExtractor(nf, H, W, WEIGHTS)                      # default flags: everything operator() fills
    frames = synth.make_stream(H, W, 8, seed=3, n_shapes=int(400 * H * W / (752 * 480)))
    for f in frames[:3]:
        ex.extract(f)
    ts = []
    for i in range(40):
        t0 = time.perf_counter()
        o = ex(frames[i % 8])
        ts.append(time.perf_counter() - t0)
    ts = np.sort(np.array(ts)) * 1e3
    print(f"{W}x{H} nf={nf}: blocking operator() incl. H2D + all D2H (heat_, heat_inv_, cov): median {np.median(ts):.3f} ms, "
          f"p90 {ts[int(0.9 * len(ts))]:.3f} ms, n={len(o[0])}", flush=True)
    ex.close()
