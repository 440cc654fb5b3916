"""One launch of every kernel on the product path, for `ncu --set full` (round 2): the default 64-frame launch plan with
computeCovariance and the in-pipeline matcher, the exact-mode plan (16 frames), the descriptor-set matcher (mutual NN and
2-NN, 2001 x 1777 rows) and a guided search on device-resident sets.   python tools/r02_kernels.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sp_orb_slam_b200 import SPExtractor, capi, synth  # noqa: E402

W8 = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")
H, W = 480, 752
frames = synth.make_stream(H, W, 48, seed=1234, n_shapes=900)
batch = [frames[i % 48] for i in range(64)]
ex = SPExtractor(800, H, W, W8, max_batch=64, emit_heat=False, emit_heat_inv=False, emit_cov=True, match_prev=True, lazy_heat=True, desc_f16=True)
outs = ex.extract_batch(batch)
print("default plan:", ex.launch_count(), "launches,", outs[1]["n"], "key points")
rng = np.random.RandomState(0)
q = rng.randn(2001, 256).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
t = np.concatenate([q[:1200] + 0.05 * rng.randn(1200, 256).astype(np.float32), rng.randn(577, 256).astype(np.float32)])
t /= np.linalg.norm(t, axis=1, keepdims=True)
ex.match(q, t)
ex.knn2(q, t)
kset = ex.desc_set(1024).from_frame(0, 1)
o = outs[1]
m = 1000
src = rng.randint(0, o["n"], m)
qdesc = (o["desc"][src] + 0.02 * rng.randn(m, 256)).astype(np.float32)
qset = ex.desc_set(1024).upload(qdesc)
qxy = (o["kp_xy"][src] + rng.uniform(-3, 3, (m, 2))).astype(np.float32)
ex.search_guided(qset, qxy, 7.0, o["occ_grid"], o["kp_xy"], kset, mode=capi.GUIDED_AREA, best_init=256.0, th_le=0.7, th_lt=0.7)
ex.close()
exx = SPExtractor(800, H, W, W8, max_batch=16, emit_heat=False, emit_cov=False, exact=True)
exx.extract_batch(batch[:16])
print("exact plan:", exx.launch_count(), "launches")
exx.close()
