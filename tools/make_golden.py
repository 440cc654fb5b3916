"""Mint the golden fixtures under tests/golden/ (run in the build container,
where /root/reference exists).  The network half comes from the REFERENCE's own
SPFrontend compiled by oracle/ref_build.sh (oracle/_ref/libspref.so); the
post-processing half from oracle/sp_post.c; matches additionally from
cv2.BFMatcher(NORM_L2, crossCheck=True), the routine the reference calls.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_frontend as R, sp_oracle as O, weights as OW  # noqa: E402
from sp_orb_slam_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = [  # name, H, W, seed, nfeatures, n_shapes
    ("g120x160", 120, 160, 11, 800, None),
    ("g480x640", 480, 640, 5, 800, None),       # BASELINE.json configs[0]: single 640x480 frame
    ("g480x752", 480, 752, 7, 800, None),       # configs[1]: EuRoC geometry
    ("g480x752_cap", 480, 752, 21, 300, 700),   # dense scene, exercises the nf+1 cap
    ("g240x320_ragged", 240, 328, 9, 800, None),
]


def main():
    assert R.available(), "run oracle/ref_build.sh first"
    import cv2
    w = OW.read_spw(os.path.join(GOLD, "superpoint_v1.spw"))
    for name, H, W, seed, nf, shapes in CASES:
        frames = synth.make_stream(H, W, 2, seed=seed, n_shapes=shapes)
        rec = {"frames": frames, "nfeatures": np.int32(nf)}
        outs = []
        for t in range(2):
            fwd = R.forward(w, frames[t])                       # the reference's own network code
            o = O.postprocess(fwd, H, W, nf)
            orc = O.frontend_forward(w, frames[t])              # restatement: margins + dense taps
            outs.append(o)
            top2 = np.sort(orc["nodust"], axis=0)[-2:]
            rec.update({
                f"f{t}_n": np.int32(o["n"]), f"f{t}_kp_xy": o["kp_xy"].astype(np.int16), f"f{t}_score": o["score"],
                f"f{t}_desc": o["desc"].astype(np.float16), f"f{t}_occ_grid": o["occ_grid"],
                f"f{t}_dense_dust": o["dense_dust"].astype(np.float16), f"f{t}_semi_dust": o["semi_dust"].astype(np.float16),
                f"f{t}_cov2": o["cov2"], f"f{t}_response": o["kp_response"],
                f"f{t}_heat_minmax": np.array([o["heat_min"], o["heat_max"]]),
                f"f{t}_heat_q": np.round(o["heat"] * 255).astype(np.uint8),          # 8-bit copy, tolerance 1/255
                f"f{t}_score_map": orc["score_map"], f"f{t}_argmax": orc["argmax"].astype(np.uint8),
                f"f{t}_argmax_margin": (top2[1] - top2[0]).astype(np.float32),       # top-1 minus top-2 probability
                f"f{t}_cand_pixels": fwd["pixels_in"].astype(np.int16), f"f{t}_cand_score": fwd["score"],
            })
        q2t, dist, sec = O.match_mutual_nn(outs[1]["desc"], outs[0]["desc"])
        ref = -np.ones(len(q2t), np.int32)
        for m in cv2.BFMatcher(cv2.NORM_L2, True).match(outs[1]["desc"], outs[0]["desc"]):
            ref[m.queryIdx] = m.trainIdx
        assert np.array_equal(ref, q2t), "oracle matcher disagrees with cv2.BFMatcher"
        rec.update(match_q2t=q2t, match_dist=dist, match_second=sec)
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{name}: n = {outs[0]['n']}, {outs[1]['n']}  matches {int((q2t >= 0).sum())}  -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
