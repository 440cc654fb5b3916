"""Latency of spfe_dust_pose_optimize (one launch = the whole 40-iteration Levenberg solve) as a caller sees it
(blocking call: uploads, launch, copies back), beside the oracle's C restatement of the g2o loop on one host core.
usage: python tools/dust_pose_time.py [reps]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import sp_oracle as O  # noqa: E402  (checker / CPU baseline only)
from sp_orb_slam_b200 import SPExtractor, synth  # noqa: E402
from test_pose_dust import CAM, make_scene  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ex = SPExtractor(800, 480, 752, os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw"), emit_heat=False, emit_cov=False, max_batch=2)
ex.extract_batch(list(synth.make_stream(480, 752, 2, seed=5)))
out = {}
for n in (100, 300, 1000):
    s = make_scene(70 + n, n=n)
    ref = O.dust_pose_optimize(s["dust"], s["start"], s["Xw"], *CAM)
    for name, kw in (("host_map", dict(dust=s["dust"])), ("device_map", dict(slot=0, frame=0))):
        for _ in range(10):
            r = ex.dust_pose_optimize(s["start"], s["Xw"], *CAM, **kw)
        t0 = time.perf_counter()
        for _ in range(reps):
            r = ex.dust_pose_optimize(s["start"], s["Xw"], *CAM, **kw)
        out[f"gpu_ms_n{n}_{name}"] = round((time.perf_counter() - t0) / reps * 1e3, 4)
        out[f"iters_n{n}_{name}"] = [int(r["n_iter"]), int(r["stats"][2])]
    t0 = time.perf_counter()
    for _ in range(max(reps // 4, 5)):
        O.dust_pose_optimize(s["dust"], s["start"], s["Xw"], *CAM)
    out[f"oracle_c_ms_n{n}"] = round((time.perf_counter() - t0) / max(reps // 4, 5) * 1e3, 4)
    out[f"oracle_iters_n{n}"] = [int(ref["n_iter"]), int(ref["stats"][2])]
# throughput mode: one solve per frame of a batch, one launch (one CTA per problem).  The ctypes arguments are built once
# so that the loop times the C call (staging memcpy + H2D + kernel + D2H), not the Python packing.
import ctypes as C  # noqa: E402
from sp_orb_slam_b200 import capi  # noqa: E402
s = make_scene(7, n=300)
one = ex.dust_pose_optimize(s["start"], s["Xw"], *CAM, dust=s["dust"])
out["batch_problem"] = {"n": 300, "iters": int(one["n_iter"]), "trials": int(one["stats"][2])}
for cnt in (1, 16, 64, 148, 296, 592):
    arr = (capi.DustPose * cnt)()
    d, Xw, dust = ex._dust_struct(s["Xw"], s["dust"], *CAM, 0.9, 0.9, 40, 0, 0)
    for i in range(cnt):
        arr[i] = d
    start = np.tile(s["start"], (cnt, 1))
    poses = start.copy()
    ninl, nit = np.zeros(cnt, np.int32), np.zeros(cnt, np.int32)
    vp = C.c_void_p

    def call():
        poses[:] = start
        rc = ex._lib.spfe_dust_pose_optimize_batch(ex._ctx, arr, cnt, vp(poses.ctypes.data), None, None, vp(ninl.ctypes.data), vp(nit.ctypes.data))
        assert rc == 0
    for _ in range(3):
        call()
    assert np.array_equal(poses[-1], one["pose"]) and int(nit[-1]) == one["n_iter"]
    k = max(reps // 5, 5)
    t0 = time.perf_counter()
    for _ in range(k):
        call()
    dt = (time.perf_counter() - t0) / k
    out[f"batch{cnt}_ms"] = round(dt * 1e3, 4)
    out[f"batch{cnt}_solves_per_s"] = round(cnt / dt)
ex.close()
print(json.dumps(out))
