"""Blocking single-frame latency of the reference-style call (spfe_extract == SPExtractor::operator(), reference
sp_extractor.cpp:361-514): the C call alone (ctypes, no result unpacking) with the launch plan replayed as a CUDA graph
(default) and enqueued call by call (SPFE_GRAPH=0), and through the Python mirror.   python tools/latency.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sp_orb_slam_b200 import SPExtractor, capi, synth  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")


def c_call_ms(ex, frames, reps=60):
    o = capi.FrameOut()
    ts = []
    for i in range(reps + 5):
        f = frames[i % len(frames)]
        t0 = time.perf_counter()
        rc = ex._lib.spfe_extract(ex._ctx, f.ctypes.data_as(C.c_void_p), f.strides[0], C.byref(o))
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    return np.sort(np.array(ts[5:])) * 1e3, o.n


for H, W, nf in [(480, 752, 800), (480, 640, 800), (1080, 1920, 2000)]:
    frames = synth.make_stream(H, W, 8, seed=3, n_shapes=int(400 * H * W / (752 * 480)))
    row = {}
    for mode, env in (("graph", "1"), ("call-by-call", "0")):
        os.environ["SPFE_GRAPH"] = env
        for name, kw in (("all of operator()'s outputs", {}), ("lazy heat", dict(emit_heat=False, emit_heat_inv=False, lazy_heat=True))):
            ex = SPExtractor(nf, H, W, WEIGHTS, **kw)
            ts, n = c_call_ms(ex, frames)
            row[(mode, name)] = (float(np.median(ts)), float(ts[int(0.9 * len(ts))]), n)
            if mode == "graph" and not kw:
                t0 = time.perf_counter()
                for i in range(20):
                    ex(frames[i % 8])
                row["python"] = (time.perf_counter() - t0) / 20 * 1e3
            ex.close()
    os.environ.pop("SPFE_GRAPH")
    print(f"{W}x{H} nf={nf} (n={row[('graph', chr(97) + 'll of operator()' + chr(39) + 's outputs')][2]}): blocking spfe_extract incl. staging, H2D, kernels, all D2H:")
    for k, v in row.items():
        if k != "python":
            print(f"    {k[0]:13s} {k[1]:30s} median {v[0]:.3f} ms   p90 {v[1]:.3f} ms")
    print(f"    through the Python mirror (numpy copies of every output): {row['python']:.3f} ms", flush=True)
