"""Timeline of the fused conv1 kernel's conv1a chain (build with -DSPFE_C1M_TRACE; reads g_c1m_trace through cuda-python)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sp_orb_slam_b200 import SPExtractor, synth, capi
import torch
H, W, B = 480, 752, 64
ex = SPExtractor(800, H, W, os.path.join(ROOT, "tests/golden/superpoint_v1.spw"), max_batch=B, emit_heat=False, emit_cov=False)
fr = synth.make_stream(H, W, 8, seed=3)
d = torch.from_numpy(np.concatenate([fr] * 8)[:B]).cuda()
for _ in range(3):
    ex.submit_device(0, d.data_ptr(), B)
ex.sync(0)
rt = C.CDLL("libcudart.so")
lib = capi.load()
buf = np.zeros((16, 10), np.int64)
sym = C.c_void_p.in_dll(lib, "spfe_c1m_trace_ptr")
rt.cudaMemcpy(buf.ctypes.data_as(C.c_void_p), sym, buf.nbytes, 2)
t = buf.astype(np.float64)
base = t[:, 0:1]
print("per item (cycles after c1a(n) issue): c_full seen, s_empty seen, work done, arrived, e1_done seen by MMA, c1b(n) issued | peer: c_full wait, work, arrive")
for n in range(16):
    print(n + 200, [int(x) for x in (t[n, 1:7] - base[n])], "| peer", [int(t[n, 8]), int(t[n, 7]), int(t[n, 9])], " c1a issue period", int(t[n, 0] - t[n - 1, 0]) if n else 0)
ex.close()
