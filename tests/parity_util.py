"""Margin analysis for key-point set comparisons (test infrastructure).

Two evaluations of the same network that differ by rounding can only disagree on a key point where one of the
reference's own decisions was a near-tie.  For every 8x8 cell the *fragility* of the oracle's decisions is the
smallest log-ratio margin among

  thr    |ln(score / 0.007)|                         the >= 0.007 test                       (sp_extractor.cpp:122)
  arg    ln(p1 / p2) of the best two of 64 positions the arg-max that picks the pixel        (:112-119)
  order  |ln(s_c / s_d)| over NMS-competing candidates d of the 8 neighbour cells           (greedy order, :194-214)
  cap    |ln(s_c / s_cut)| when more than nf + 1 candidates survive the suppression          (the cap, :211)

(log ratios, because the score error of a rounded evaluation is proportional to the score: d s ~ s (1 - s) d logit.)
A differing key point is *explained at eps* if a cell within two cells of it (its own decisions, or one cascade step
through the 9x9 suppression window) has fragility < eps.  explain_differences returns, per differing pixel, the
smallest eps that explains it, so that a test can state its bar and a bench can report the measured one.
"""
from __future__ import annotations

import numpy as np

from oracle import sp_oracle as O


def cell_fragility(fwd: dict, nf: int, radius: int = O.NMS_RADIUS) -> np.ndarray:
    """fwd: oracle forward: score_map, argmax and either nodust or argmax_margin (best minus second-best probability,
    what the golden fixtures store).  -> [hc, wc] float32 log-margin fragility."""
    s = fwd["score_map"].astype(np.float64)
    hc, wc = s.shape
    thr = np.abs(np.log(np.maximum(s, 1e-30) / O.SCORE_THRESH))
    if "nodust" in fwd:
        top2 = np.sort(fwd["nodust"].astype(np.float64), axis=0)[-2:]
    else:
        top2 = np.stack([s - fwd["argmax_margin"].astype(np.float64), s])
    arg = np.log(np.maximum(top2[1], 1e-30) / np.maximum(top2[0], 1e-30))
    cand = s >= O.SCORE_THRESH * 0.8                       # anything that could become a candidate under a 20 % score error
    am = fwd["argmax"].astype(np.int64)
    cy, cx = np.mgrid[0:hc, 0:wc]
    px, py = cx * 8 + am % 8, cy * 8 + am // 8
    order = np.full((hc, wc), np.inf)
    big = 1 << 20
    P = lambda a, fill: np.pad(a, 1, constant_values=fill)
    ps, pc, ppx, ppy = P(s, 1.0), P(cand, False), P(px, big), P(py, big)
    for dy in range(3):
        for dx in range(3):
            if dy == 1 and dx == 1:
                continue
            ns, nc = ps[dy:dy + hc, dx:dx + wc], pc[dy:dy + hc, dx:dx + wc]
            near = (np.abs(ppx[dy:dy + hc, dx:dx + wc] - px) <= radius) & (np.abs(ppy[dy:dy + hc, dx:dx + wc] - py) <= radius)
            m = np.where(cand & nc & near, np.abs(np.log(np.maximum(s, 1e-30) / np.maximum(ns, 1e-30))), np.inf)
            order = np.minimum(order, m)
    frag = np.minimum(np.minimum(thr, arg), order)
    # cap: rank of the cut among the oracle's suppression survivors (before the border filter)
    mask = fwd["score_map"] >= np.float32(O.SCORE_THRESH)
    pts = np.stack([px[mask], py[mask]], 1).astype(np.float32)
    sc = fwd["score_map"][mask]
    ordr = O.sort_desc(sc)
    H, W = hc * 8, wc * 8
    sel, _ = O.nms(pts[ordr], 10 ** 6, W, H, border=0)
    kept_scores = np.sort(sc[ordr][sel].astype(np.float64))[::-1]
    if len(kept_scores) > nf + 1:
        cut = 0.5 * (kept_scores[nf] + kept_scores[nf + 1])
        frag = np.minimum(frag, np.where(cand, np.abs(np.log(np.maximum(s, 1e-30) / cut)), np.inf))
    return frag.astype(np.float32)


def explain_differences(fwd: dict, ref_kp, got_kp, nf: int):
    """-> list of (x, y, eps_needed) for every pixel in the symmetric difference of the two key-point sets."""
    rs = {(int(x), int(y)) for x, y in ref_kp}
    gs = {(int(x), int(y)) for x, y in got_kp}
    diff = sorted(rs ^ gs)
    if not diff:
        return []
    frag = cell_fragility(fwd, nf)
    hc, wc = frag.shape
    fp = np.pad(frag, 2, constant_values=np.inf)
    out = []
    for x, y in diff:
        cy, cx = y // 8, x // 8
        out.append((x, y, float(fp[cy:cy + 5, cx:cx + 5].min())))
    return out
