import sys, numpy as np
sys.path.insert(0, '.')
from sp_orb_slam_b200 import SPExtractor, synth
import torch
H, W, B = 480, 752, int(sys.argv[1]) if len(sys.argv) > 1 else 8
ex = SPExtractor(800, H, W, 'tests/golden/superpoint_v1.spw', max_batch=B, emit_heat=True, emit_cov=True)
frames = synth.make_stream(H, W, B, seed=1234, n_shapes=400)
outs = ex.extract_batch(list(frames))
q = ex.debug_read(0, 'cov_qlen', B)
print('replayed (keypoints, pixels) per frame:', ex.debug_read(0, 'cov_replayed', B).tolist()[:8])
print('counters [grabbed, big, pending]:', ex.debug_read(0, 'cov_counters', 1)[0].tolist())
for t, o in enumerate(outs[:3]):
    n = o['n']; ql = q[t, :n]
    print('frame', t, 'n', n, 'flood len: mean %.1f median %d max %d' % (ql.mean(), np.median(ql), ql.max()), 'heat_inv>0 frac %.3f' % (o['heat_inv'] > 0).mean())
d = torch.from_numpy(frames).cuda()
for rep in range(2):
    st = ex.profile_device(0, d.data_ptr(), B)
print({s['name']: round(s['ms'], 4) for s in st if s['name'].startswith(('cov', 'heat', 'nms'))})
