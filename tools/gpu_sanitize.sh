#!/bin/bash
# compute-sanitizer over a small extract + match + covariance + guided search (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from sp_orb_slam_b200 import SPExtractor, SPMatcher, synth, capi
H, W = 120, 160
ex = SPExtractor(100, H, W, 'tests/golden/superpoint_v1.spw', max_batch=3, num_slots=2, match_prev=True)
fr = synth.make_stream(H, W, 3, seed=5, n_shapes=40)
o = ex.extract_batch(list(fr))
ex.submit(1, list(fr[:2])); o2 = ex.wait(1, 2)
q2t, d = ex.match(o[0]['desc'], o[1]['desc'])
idx, dd = ex.knn2(o[0]['desc'], o[1]['desc'])
g = ex.search_guided(o[0]['desc'], o[0]['kp_xy'], 7.0, o[1]['occ_grid'], o[1]['kp_xy'], o[1]['desc'], mode=capi.GUIDED_AREA, best_init=256.0, th_le=0.7, th_lt=0.7)
print('ok', [x['n'] for x in o], int((q2t >= 0).sum()), int((g[0] >= 0).sum()))
ex.close()
PY
for tool in memcheck racecheck; do
  echo "=== $tool" | tee gpurun_out/sanitizer_$tool.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py 2>&1 | tail -25 | tee -a gpurun_out/sanitizer_$tool.log
done
