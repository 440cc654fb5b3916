import os, sys
import numpy as np
sys.path.insert(0, '.')
from sp_orb_slam_b200 import SPExtractor, synth
H, W = 480, 752
ex = SPExtractor(800, H, W, 'tests/golden/superpoint_v1.spw', max_batch=16)
fr = synth.make_stream(H, W, 16, seed=1234, n_shapes=900)
o = ex.extract_batch(list(fr))
q = ex.debug_read(0, "cov_qlen", 16)
n = np.array([x["n"] for x in o])
ql = np.concatenate([q[b, :n[b]] for b in range(16)])
print("floods", ql.size, "mean pops", ql.mean(), "pct", {p: int(np.percentile(ql, p)) for p in (50, 75, 90, 95, 99, 99.9)}, "max", ql.max())
for cap in (64, 128, 256, 512, 1024):
    print("  pops <=", cap, ":", round(float((ql <= cap).mean()) * 100, 2), "%")
print("counters", ex.debug_read(0, "cov_counters", 1), "replayed", ex.debug_read(0, "cov_replayed", 16).sum(0))
