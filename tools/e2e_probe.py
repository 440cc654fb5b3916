"""Where does the end-to-end (host-pointer) path lose time against the device-resident one?
Sweeps slots / batch and reports host time inside submit vs wait.  python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sp_orb_slam_b200 import SPExtractor, synth  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")
H, W = 480, 752


def run(B, S, steps=60):
    import torch
    ex = SPExtractor(800, H, W, WEIGHTS, max_batch=B, num_slots=S, emit_heat=False, emit_cov=False, match_prev=True)
    n_pool = max(4, -(-(140 << 20) // (B * H * W)))
    frames = synth.make_stream(H, W, 16, seed=1234, n_shapes=400)
    pinned = ex.pinned_frames(n_pool * B).reshape(n_pool, B, H, W)
    pinned[:] = frames[np.arange(n_pool * B) % 16].reshape(n_pool, B, H, W)
    d_pool = torch.from_numpy(pinned.copy()).cuda()
    for i in range(5):
        ex.submit_device(0, d_pool.data_ptr() + (i % n_pool) * B * H * W, B)
    ex.sync(0)
    ex.timer_start(0)
    for i in range(steps):
        ex.submit_device(0, d_pool.data_ptr() + (i % n_pool) * B * H * W, B)
    ms = ex.timer_stop(0)
    dev_fps = steps * B / ms * 1e3
    for i in range(S):
        ex.submit_pinned(i, pinned[i % n_pool])
        ex.wait(i, B, unpack=False)
    t_sub = t_wait = 0.0
    t0 = time.perf_counter()
    for i in range(steps):
        s = i % S
        if i >= S:
            a = time.perf_counter()
            ex.wait(s, B, unpack=False)
            t_wait += time.perf_counter() - a
        a = time.perf_counter()
        ex.submit_pinned(s, pinned[i % n_pool])
        t_sub += time.perf_counter() - a
    for i in range(steps, steps + S):
        ex.wait(i % S, B, unpack=False)
    dt = time.perf_counter() - t0
    print(f"B={B:3d} S={S}: device {dev_fps:8.0f} fps | e2e {steps * B / dt:8.0f} fps  {dt / steps * 1e3:6.3f} ms/step  "
          f"host in submit {t_sub / steps * 1e3:6.3f} ms/step, in wait {t_wait / steps * 1e3:6.3f} ms/step", flush=True)
    ex.close()


if __name__ == "__main__":
    for B, S in [(32, 1), (32, 2), (32, 3), (32, 4), (16, 4), (64, 3), (8, 6)]:
        run(B, S)
