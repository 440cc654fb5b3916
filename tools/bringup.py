"""GPU bring-up / diagnostics: compare every intermediate tensor of the CUDA
path with the CPU oracle.  Test infrastructure (imports oracle/).

    python tools/bringup.py layers H W [batch]
    python tools/bringup.py match N
    python tools/bringup.py profile H W batch
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sp_oracle as O, weights as OW  # noqa: E402
from sp_orb_slam_b200 import SPExtractor, synth  # noqa: E402

WEIGHTS = os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw")


def stat(name, got, ref):
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    err = np.abs(got - ref)
    scale = np.abs(ref).max() + 1e-12
    print(f"  {name:12s} shape {str(got.shape):18s} max|ref| {scale:9.4f}  max err {err.max():.3e}  rel {err.max() / scale:.3e}  "
          f"mean err {err.mean():.3e}  nan {int(np.isnan(got).sum())}", flush=True)
    return err.max() / scale


def layers(H, W, batch):
    w = OW.read_spw(WEIGHTS)
    frames = synth.make_stream(H, W, batch, seed=7, n_shapes=max(12, int(400 * H * W / (752 * 480))))
    ex = SPExtractor(800, H, W, WEIGHTS, max_batch=batch, num_slots=1, emit_heat=True, emit_cov=True)
    t0 = time.time()
    outs = ex.extract_batch(list(frames))
    print(f"extract_batch({batch} x {H}x{W}) ok in {time.time() - t0:.3f}s, launches so far {ex.launch_count()}", flush=True)
    for b in range(batch):
        print(f"frame {b}:", flush=True)
        fwd = O.frontend_forward(w, frames[b], keep_layers=True)
        for name in ["conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"]:
            got = ex.debug_read(0, name, batch)[b]
            stat(name, got, fwd["layers"][name].transpose(1, 2, 0))
        heads = ex.debug_read(0, "heads", batch)[b]
        stat("convPa", heads[..., :256], fwd["layers"]["convPa"].transpose(1, 2, 0))
        stat("convDa", heads[..., 256:], fwd["layers"]["convDa"].transpose(1, 2, 0))
        stat("coarse", ex.debug_read(0, "coarse", batch)[b], fwd["coarse"].transpose(1, 2, 0))
        stat("score", ex.debug_read(0, "score", batch)[b], fwd["score_map"])
        am = ex.debug_read(0, "argmax", batch)[b]
        print(f"  argmax mismatches {int((am != fwd['argmax']).sum())} of {am.size}", flush=True)
        stat("semi_dust", ex.debug_read(0, "semi_dust", batch)[b], fwd["semi_dust"])
        stat("dense_dust", ex.debug_read(0, "dense_dust", batch)[b], fwd["dense_dust"])
        stat("heat_log", ex.debug_read(0, "heat_log", batch)[b], fwd["heat_log"])
        ref = O.postprocess(fwd, H, W, 800)
        o = outs[b]
        stat("heat", o["heat"], ref["heat"])
        stat("heat_inv", o["heat_inv"], ref["heat_inv"])
        print(f"  keypoints gpu {o['n']} oracle {ref['n']}", flush=True)
        gs = {(int(x), int(y)) for x, y in o["kp_xy"]}
        rs = {(int(x), int(y)) for x, y in ref["kp_xy"]}
        print(f"  keypoint sets: common {len(gs & rs)} only-gpu {len(gs - rs)} only-oracle {len(rs - gs)}", flush=True)
        # NMS exactness given the GPU's own score map: run the oracle's NMS on the GPU scores
        sc = ex.debug_read(0, "score", batch)[b]
        mask = sc >= np.float32(0.007)
        cy, cx = np.nonzero(mask)
        px = cx * 8 + am[mask] % 8
        py = cy * 8 + am[mask] // 8
        pts = np.stack([px, py], 1).astype(np.float32)
        order = O.sort_desc(sc[mask])
        sel, occ = O.nms(pts[order], 800, W, H)
        kp_ref = pts[order][sel]
        same = o["n"] == len(kp_ref) and np.array_equal(o["kp_xy"], kp_ref) and np.array_equal(o["occ_grid"], occ)
        print(f"  NMS on GPU scores == oracle NMS on the same scores: {same} (n={len(kp_ref)})", flush=True)
        # descriptors of common keypoints
        rmap = {(int(x), int(y)): i for i, (x, y) in enumerate(ref["kp_xy"])}
        cos = [float(np.dot(o["desc"][i], ref["desc"][rmap[(int(x), int(y))]])) for i, (x, y) in enumerate(o["kp_xy"]) if (int(x), int(y)) in rmap]
        if cos:
            print(f"  descriptor cosine on common keypoints: min {min(cos):.6f} mean {np.mean(cos):.6f}", flush=True)
        if "cov2" in o and o["n"] == ref["n"] and np.array_equal(o["kp_xy"], ref["kp_xy"]):
            stat("cov2", o["cov2"], ref["cov2"])
            stat("response", o["kp_response"], ref["kp_response"])
    ex.close()


def match(n):
    rng = np.random.RandomState(0)
    ex = SPExtractor(800, 64, 64, WEIGHTS, emit_heat=False, emit_cov=False)
    for nq, nt in [(n, n), (n, n // 2 + 3), (5, n), (1, 1), (n, 0)]:
        q = rng.randn(nq, 256).astype(np.float32)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        t = (q[rng.permutation(nq)[:nt]] if nt <= nq else np.concatenate([q, rng.randn(nt - nq, 256).astype(np.float32)]))
        t = t + 0.05 * rng.randn(*t.shape).astype(np.float32)
        t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-9)
        got, gd = ex.match(q, t)
        ref, rd, _ = O.match_mutual_nn(q, t)
        print(f"match nq={nq} nt={nt}: identical {np.array_equal(got, ref)} matched {int((ref >= 0).sum())} "
              f"max dist err {np.abs(gd - rd).max() if nq and nt else 0:.2e}", flush=True)
    ex.close()


def profile(H, W, batch):
    import torch
    frames = synth.make_stream(H, W, min(batch, 8), seed=3)
    frames = np.concatenate([frames] * ((batch + len(frames) - 1) // len(frames)))[:batch]
    ex = SPExtractor(800, H, W, WEIGHTS, max_batch=batch, num_slots=1, emit_heat=False, emit_cov=False)
    d = torch.from_numpy(frames).cuda()
    for _ in range(3):
        ex.submit_device(0, d.data_ptr(), batch)
    ex.sync(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = ex.profile_device(0, d.data_ptr(), batch)
    st = ex.profile_device(0, d.data_ptr(), batch)
    tot = sum(s["ms"] for s in st)
    for s in st:
        tf = s["flop"] / (s["ms"] * 1e-3) / 1e12 if s["ms"] > 0 else 0
        gb = s["bytes"] / (s["ms"] * 1e-3) / 1e9 if s["ms"] > 0 else 0
        print(f"  {s['name']:14s} {s['ms']:8.3f} ms  {100 * s['ms'] / tot:5.1f}%  {tf:8.1f} TFLOP/s  {gb:8.1f} GB/s", flush=True)
    t0 = time.time()
    n_it = 10
    for _ in range(n_it):
        ex.submit_device(0, d.data_ptr(), batch)
    ex.sync(0)
    dt = time.time() - t0
    print(f"profile {H}x{W} batch {batch}: staged total {tot:.3f} ms -> {batch / tot * 1e3:.0f} fps; back-to-back {batch * n_it / dt:.0f} fps", flush=True)
    ex.close()


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "layers":
        layers(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    elif cmd == "match":
        match(int(sys.argv[2]))
    elif cmd == "profile":
        profile(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
