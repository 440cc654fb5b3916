"""Smallest possible run of the dust-pose kernel (host maps only, no frame is extracted first), meant for
`compute-sanitizer --tool racecheck python tools/dust_pose_min.py`: under racecheck the tensor-core kernels of an
extraction take minutes, this takes seconds.  Prints the parity of each call against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import sp_oracle as O  # noqa: E402  (checker only)
from sp_orb_slam_b200 import SPExtractor  # noqa: E402
from test_pose_dust import CAM, make_scene  # noqa: E402

ex = SPExtractor(100, 64, 64, os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw"), emit_heat=False, emit_cov=False, max_batch=1)
for n, shape in ((300, (60, 94)), (40, (60, 80)), (700, (135, 240))):
    s = make_scene(300 + n, n=n, rows=shape[0], cols=shape[1])
    ref = O.dust_pose_optimize(s["dust"], s["start"], s["Xw"], *CAM)
    got = ex.dust_pose_optimize(s["start"], s["Xw"], *CAM, dust=s["dust"])
    lin = ex.dust_linearize(s["start"], s["Xw"], *CAM, dust=s["dust"])
    ok = got["n_iter"] == ref["n_iter"] and np.array_equal(got["visible"], ref["visible"]) and np.abs(got["pose"] - ref["pose"]).max() < 1e-9
    print(f"n={n} map={shape}: iterations {got['n_iter']} inliers {got['n_inlier']} parity {'ok' if ok else 'FAILED'} chi2 {lin['chi2']:.6f}")
res = ex.dust_pose_optimize_batch([dict(pose=s["start"], Xw=s["Xw"], cam=CAM, dust=s["dust"]) for _ in range(8)])
print("batch of 8:", all(np.array_equal(r["pose"], got["pose"]) for r in res))
ex.close()
