"""Default-mode (fp16) key-point differences against the oracle and the oracle margin that explains each of them
(tests/parity_util.py); the numbers behind the gates in tests/test_gpu_parity.py.   python tools/parity_margins.py"""
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


def main():
    w = OW.read_spw(W8)
    exact = "--exact" in sys.argv
    cases = [("g480x752", None), ("g480x640", None), ("g480x752_cap", None), ("g240x320_ragged", None), ("g120x160", None),
             ("dense752", (480, 752, 800, 900, 8)), ("sparse752", (480, 752, 800, 250, 8)), ("s320", (240, 320, 800, 260, 8)),
             ("1080p", (1080, 1920, 2000, 3600, 2))]
    worst = dict(abs=0.0, rel=0.0, eps=0.0, n=0)
    for name, spec in cases:
        if spec is None:
            z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
            frames, nf = z["frames"], int(z["nfeatures"])
        else:
            H, W, nf, shapes, cnt = spec
            frames = synth.make_stream(H, W, cnt, seed=77, n_shapes=shapes)
        H, W = frames.shape[1:]
        ex = SPExtractor(nf, H, W, W8, max_batch=len(frames), emit_heat=False, emit_cov=False, exact=exact)
        outs = ex.extract_batch(list(frames))
        score = ex.debug_read(0, "score", len(frames))
        argmax = ex.debug_read(0, "argmax", len(frames))
        ex.close()
        for t, (f, o) in enumerate(zip(frames, outs)):
            ref = O.extract(w, f, nf, keep_forward=True)
            fwd = ref["forward"]
            sm = fwd["score_map"]
            d = np.abs(score[t] - sm)
            sel = sm >= 0.0035
            rel = float((d[sel] / sm[sel]).max()) if sel.any() else 0.0
            if not hasattr(main, "arg"):
                main.arg = 0.0
            top2 = np.sort(fwd["nodust"].astype(np.float64), axis=0)[-2:]
            flip = (argmax[t] != fwd["argmax"]) & (sm >= 0.0035)
            if flip.any():
                main.arg = max(main.arg, float(np.log(top2[1] / top2[0])[flip].max()))
            ex_ = explain_differences(fwd, ref["kp_xy"], o["kp_xy"], nf)
            eps = max([e for _, _, e in ex_], default=0.0)
            print(f"{name}[{t}] n_ref {ref['n']:4d} n_gpu {o['n']:4d} diff {len(ex_):2d} eps_needed {eps:.4f} max|ds| {d.max():.2e} max rel ds (s>=0.0035) {rel:.3e}", flush=True)
            worst["abs"] = max(worst["abs"], float(d.max())); worst["rel"] = max(worst["rel"], rel)
            worst["eps"] = max(worst["eps"], eps); worst["n"] = max(worst["n"], len(ex_))
    print("worst:", worst, "largest arg-max log-ratio margin among flipped cells with score >= 0.0035:", getattr(main, "arg", 0.0))


if __name__ == "__main__":
    main()
