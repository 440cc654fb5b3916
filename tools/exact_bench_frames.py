"""The exact mode on the 201 frames of bench.py's parity block: which frames differ from the oracle, and the oracle margin
(tests/parity_util.py) that explains each differing key point.   python tools/exact_bench_frames.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import sp_oracle as O, weights as OW  # noqa: E402
from parity_util import explain_differences  # noqa: E402
from sp_orb_slam_b200 import SPExtractor, synth  # noqa: E402

W8 = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")
H, W, nf, shapes, total = 480, 752, 800, 900, 200
scenes = [synth.make_stream(H, W, min(48, total + 1 - o), seed=4321 + o, n_shapes=shapes) for o in range(0, total + 1, 48)]
frames = np.concatenate(scenes)[:total + 1]                      # exactly bench.py's cpu_baseline frames
w = OW.read_spw(W8)
ex = SPExtractor(nf, H, W, W8, max_batch=8, emit_heat=False, emit_cov=False, exact=True)
ndiff = 0
for i0 in range(0, len(frames), 8):
    outs = ex.extract_batch(list(frames[i0:i0 + 8]))
    for j, o in enumerate(outs):
        ref = O.extract(w, frames[i0 + j], nf, keep_forward=True)
        d = explain_differences(ref["forward"], ref["kp_xy"], o["kp_xy"], nf)
        if d:
            ndiff += 1
            print(f"frame {i0 + j}: {len(d)} differing key points, (x, y, oracle log-ratio margin): {[(x, y, round(e, 6)) for x, y, e in d]}", flush=True)
print(f"{len(frames)} frames, {ndiff} with a difference")
