"""CPU simulation of operand precision choices for the conv stack (decides what the "exact" mode has to split).

Each 3x3 / 1x1 layer of SPFrontend::forward (sp_extractor.cpp:81-103) is evaluated as fp32-accumulated products of
operands rounded the way a tensor-core path would hold them:

    'h'  : fp16(x)                         (one MMA)
    's'  : fp16(x) + fp16(x - fp16(x))     (hi + lo: 22 significant bits)
    'f'  : x                               (fp32, the reference)

per layer for the activation operand and the weight operand.  A 's','s' layer is evaluated as the three products a
tensor-core kernel would issue, (Ah + Al) * Wh + Ah * Wl (the Al * Wl term, 2^-24 relative, is dropped).
Reports, against the fp32 oracle on the golden frames: max |score - oracle|, the candidate / key-point set difference
and the worst relative logit error.

    python tools/precision_sim.py [golden ...]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sp_oracle as O, weights as OW  # noqa: E402

LAYERS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "convPa", "convPb", "convDa", "convDb"]


def q16(x):
    return x.to(torch.float16).to(torch.float32)


def parts(x, mode):
    if mode == "f":
        return x, None
    hi = q16(x)
    if mode == "h":
        return hi, None
    return hi, q16(x - hi)


def conv(x, w, b, pad, amode, wmode):
    ah, al = parts(x, amode)
    wh, wl = parts(w, wmode)
    y = F.conv2d(ah, wh, None, padding=pad)
    if al is not None:
        y = y + F.conv2d(al, wh, None, padding=pad)
    if wl is not None:
        y = y + F.conv2d(ah, wl, None, padding=pad)
    return y + b.view(1, -1, 1, 1)


def forward(w, img, plan):
    H, W = img.shape
    x = (torch.from_numpy(img.astype(np.float32)) * np.float32(1.0 / 255.0)).view(1, 1, H, W)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in w.items()}

    def L(x, name, pad):
        a, ww = plan[name]
        return conv(x, t[name + ".weight"], t[name + ".bias"], pad, a, ww)

    with torch.no_grad():
        for blk, pool in (("1", True), ("2", True), ("3", True), ("4", False)):
            x = torch.relu(L(x, f"conv{blk}a", 1))
            x = torch.relu(L(x, f"conv{blk}b", 1))
            if pool:
                x = F.max_pool2d(x, 2, 2)
        semi = L(torch.relu(L(x, "convPa", 1)), "convPb", 0).squeeze()
        dense = torch.softmax(semi, 0)
        score, arg = dense[:-1].max(0)
    return semi.numpy(), score.numpy(), arg.numpy()


def keypoints(score, arg, H, W, nf):
    hc, wc = score.shape
    cy, cx = np.mgrid[0:hc, 0:wc]
    m = score >= np.float32(O.SCORE_THRESH)
    pts = np.stack([(cx * 8 + arg % 8)[m], (cy * 8 + arg // 8)[m]], 1).astype(np.float32)
    s = score[m]
    order = O.sort_desc(s)
    sel, _ = O.nms(pts[order], nf, W, H)
    return {(int(x), int(y)) for x, y in pts[order][sel]}, {(int(x), int(y)) for x, y in pts}


def plan_of(spec):
    """spec: dict layer -> 'ah' pairs, default given by '*'."""
    d = spec.get("*", "ff")
    return {n: tuple(spec.get(n, d)) for n in LAYERS}


PLANS = {
    "fp32 (torch, other sum order)": {"*": "ff"},
    "current: conv1a exact, rest fp16 x fp16": {"*": "hh", "conv1a": "ff"},
    "A split on conv1b only": {"*": "hh", "conv1a": "ff", "conv1b": "sh"},
    "A split everywhere, W fp16": {"*": "sh", "conv1a": "ff"},
    "W split everywhere, A fp16": {"*": "hs", "conv1a": "ff"},
    "A+W split on conv1b..conv2b, rest fp16": {"*": "hh", "conv1a": "ff", "conv1b": "ss", "conv2a": "ss", "conv2b": "ss"},
    "A+W split on the encoder, heads fp16": {"*": "ss", "conv1a": "ff", "convPa": "hh", "convPb": "hh", "convDa": "hh", "convDb": "hh"},
    "A+W split everywhere (3 MMAs)": {"*": "ss", "conv1a": "ff"},
}


def main():
    names = sys.argv[1:] or ["g480x752", "g480x640", "g480x752_cap", "g240x320_ragged", "g120x160"]
    w = OW.read_spw(os.path.join(ROOT, "tests", "golden", "superpoint_v1.spw"))
    torch.set_num_threads(os.cpu_count())
    for pname, spec in PLANS.items():
        plan = plan_of(spec)
        tot_kp = tot_cand = nkp = 0
        worst_s = worst_l = 0.0
        for g in names:
            z = np.load(os.path.join(ROOT, "tests", "golden", g + ".npz"))
            nf = int(z["nfeatures"])
            for f in range(2):
                img = z["frames"][f]
                H, W = img.shape
                ref = O.frontend_forward(w, img)
                rk, rc = keypoints(ref["score_map"], ref["argmax"], H, W, nf)
                semi, score, arg = forward(w, img, plan)
                k, c = keypoints(score, arg, H, W, nf)
                tot_kp += len(k ^ rk)
                tot_cand += len(c ^ rc)
                nkp += len(rk)
                worst_s = max(worst_s, float(np.abs(score - ref["score_map"]).max()))
                worst_l = max(worst_l, float(np.abs(semi - ref["semi"]).max()))
        print(f"{pname:48s} kp diff {tot_kp:4d} / {nkp}  cand diff {tot_cand:4d}  max|dscore| {worst_s:.2e}  max|dlogit| {worst_l:.2e}", flush=True)


if __name__ == "__main__":
    main()
