"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and
the committed golden vectors.  Integer / index work is checked bit-exactly;
floating-point work within the tolerances stated here (north star: identical
keypoint sets after NMS, descriptors within 1e-3 cosine).

This file tests the DEFAULT mode: fp16 operands / fp32 accumulation, against the fp32 reference.  (The "exact" mode,
tests/test_gpu_exact.py, reaches IDENTICAL key-point sets on every fixture and fresh frame.)  Measured on a B200 over
34 frames of five geometries (tools/parity_margins.py, profiles/r02_parity_margins.txt): score error at most 4.3 % of
the score (4.0e-3 absolute, at scores ~0.3), 0 - 10 differing key points per frame of 250 - 801, and every one of them
sits at an oracle decision whose own log-ratio margin is below 0.011 (tests/parity_util.py: threshold, arg-max, greedy
order between suppression neighbours, the nf + 1 cut).  The gates:
  (a) given the SAME score / arg-max maps, threshold + NMS + cap + border + raster order + occ_grid are bit-exact;
  (b) every difference of the key-point sets is explained by an oracle margin below EPS_MARGIN, at most
      MAX_DIFF_PER_FRAME of them (no Jaccard-style allowance);
  (c) the score map is within SCORE_RTOL * score + SCORE_ATOL of the oracle: 6.2e-4 at the 0.007 threshold.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN_CASES, ROOT, WEIGHTS
from oracle import sp_oracle as O
from parity_util import explain_differences
from sp_orb_slam_b200 import SPExtractor, SPMatcher, SpfeError, synth

pytestmark = pytest.mark.gpu

SCORE_RTOL = 6e-2        # score error relative to the score (measured <= 4.3e-2; the error is ~ s (1 - s) d logit)
SCORE_ATOL = 2e-4        # floor for tiny scores (threshold: 7e-3)
EPS_MARGIN = 0.02        # log-ratio margin that must explain every key-point difference (measured <= 0.011)
ARGMAX_EPS = 0.05        # log-ratio of the best two positions below which the arg-max of a cell may flip (measured <= 0.02)
MAX_DIFF_PER_FRAME = 12  # differing key points per frame, 752x480 at the 801 cap (measured <= 10)
COS_TOL = 1e-3           # north star: descriptors within 1e-3 cosine
LAYER_RTOL = 6e-3        # per-layer activations, relative to the layer's max |activation|


@pytest.fixture(scope="module")
def ex_cache():
    cache = {}

    def get(H, W, nf=800, **kw):
        key = (H, W, nf, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = SPExtractor(nf, H, W, WEIGHTS, **kw)
        return cache[key]
    yield get
    for e in cache.values():
        e.close()


def oracle_nms_on(score, argmax, nf, H, W):
    mask = score >= np.float32(O.SCORE_THRESH)
    cy, cx = np.nonzero(mask)
    pts = np.stack([cx * 8 + argmax[mask] % 8, cy * 8 + argmax[mask] // 8], 1).astype(np.float32)
    sc = score[mask]
    order = O.sort_desc(sc)
    sel, occ = O.nms(pts[order], nf, W, H)
    return pts[order][sel], sc[order][sel], occ


@pytest.mark.parametrize("H,W", [(64, 96), (120, 136)])
def test_layers_match_oracle(H, W, weights, ex_cache):
    ex = ex_cache(H, W, max_batch=2)
    frames = synth.make_stream(H, W, 2, seed=7, n_shapes=24)
    ex.extract_batch(list(frames))
    for b in range(2):
        fwd = O.frontend_forward(weights, frames[b], keep_layers=True)
        for name in ["conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"]:   # conv1a is fused into conv1b
            got = ex.debug_read(0, name, 2)[b].astype(np.float32)
            ref = fwd["layers"][name].transpose(1, 2, 0)
            assert np.abs(got - ref).max() <= LAYER_RTOL * np.abs(ref).max(), name
        heads = ex.debug_read(0, "heads", 2)[b].astype(np.float32)
        for name, sl in [("convPa", slice(0, 256)), ("convDa", slice(256, 512))]:
            ref = fwd["layers"][name].transpose(1, 2, 0)
            assert np.abs(heads[..., sl] - ref).max() <= LAYER_RTOL * np.abs(ref).max(), name
        coarse = ex.debug_read(0, "coarse", 2)[b].astype(np.float32)
        assert np.abs(coarse - fwd["coarse"].transpose(1, 2, 0)).max() < 3e-3
        np.testing.assert_allclose(ex.debug_read(0, "score", 2)[b], fwd["score_map"], atol=SCORE_ATOL, rtol=SCORE_RTOL)
        np.testing.assert_allclose(ex.debug_read(0, "dense_dust", 2)[b], fwd["dense_dust"], atol=1e-2)
        np.testing.assert_allclose(ex.debug_read(0, "semi_dust", 2)[b], fwd["semi_dust"], atol=8e-2, rtol=1e-2)
        np.testing.assert_allclose(ex.debug_read(0, "heat_log", 2)[b], fwd["heat_log"], atol=8e-2)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_extract_vs_golden(name, golden, weights, ex_cache):
    g = golden(name)
    nf = int(g["nfeatures"])
    _, H, W = g["frames"].shape
    ex = ex_cache(H, W, nf, max_batch=2)
    outs = ex.extract_batch(list(g["frames"]))
    for t, o in enumerate(outs):
        score = ex.debug_read(0, "score", 2)[t]
        argmax = ex.debug_read(0, "argmax", 2)[t]
        # (a) integer path bit-exact given the same score / argmax maps
        kp_ref, sc_ref, occ_ref = oracle_nms_on(score, argmax, nf, H, W)
        assert o["n"] == len(kp_ref)
        assert np.array_equal(o["kp_xy"], kp_ref)
        assert np.array_equal(o["kp_score"], sc_ref)
        assert np.array_equal(o["occ_grid"], occ_ref)
        # (c) score map within tolerance of the golden (reference-compiled) score map
        np.testing.assert_allclose(score, g[f"f{t}_score_map"], atol=SCORE_ATOL, rtol=SCORE_RTOL)
        agree = argmax == g[f"f{t}_argmax"]
        gs = g[f"f{t}_score_map"].astype(np.float64)
        arg_margin = np.log(gs / np.maximum(gs - g[f"f{t}_argmax_margin"], 1e-30))
        assert np.all(agree | (arg_margin < ARGMAX_EPS) | (gs < 0.5 * O.SCORE_THRESH))   # arg-max flips only at near-ties
        # (b) keypoint sets: every difference explained by an oracle near-tie; descriptors of common keypoints within 1e-3 cosine
        gold = {(int(x), int(y)): i for i, (x, y) in enumerate(g[f"f{t}_kp_xy"])}
        mine = {(int(x), int(y)): i for i, (x, y) in enumerate(o["kp_xy"])}
        common = set(gold) & set(mine)
        fwd = dict(score_map=g[f"f{t}_score_map"], argmax=g[f"f{t}_argmax"], argmax_margin=g[f"f{t}_argmax_margin"])
        diffs = explain_differences(fwd, g[f"f{t}_kp_xy"], o["kp_xy"], nf)
        assert len(diffs) <= MAX_DIFF_PER_FRAME, diffs
        assert all(e < EPS_MARGIN for _, _, e in diffs), diffs
        gd = g[f"f{t}_desc"].astype(np.float32)
        cos = np.array([np.dot(o["desc"][mine[k]], gd[gold[k]]) / np.linalg.norm(gd[gold[k]]) for k in common])
        assert cos.min() > 1 - COS_TOL
        np.testing.assert_allclose(np.linalg.norm(o["desc"], axis=1), 1.0, atol=1e-5)
        np.testing.assert_allclose(o["dense_dust"], g[f"f{t}_dense_dust"].astype(np.float32), atol=1.5e-2)
        np.testing.assert_allclose(o["heat"], g[f"f{t}_heat_q"] / 255.0, atol=0.5 / 255 + 2e-2)
        np.testing.assert_allclose(o["heat"] + o["heat_inv"], 1.0, atol=1e-5)


def test_every_keypoint_difference_is_an_oracle_near_tie(weights, ex_cache):
    """Fresh frames (sparse and at the cap): every key point the default mode adds or misses sits at an oracle decision
    (threshold, arg-max, greedy order, cap cut) whose own margin is below EPS_MARGIN; all others are identical."""
    for (H, W, nf, shapes, seed) in [(240, 320, 800, 260, 31), (480, 752, 800, 900, 33)]:
        ex = ex_cache(H, W, nf, max_batch=4)
        frames = synth.make_stream(H, W, 4, seed=seed, n_shapes=shapes)
        outs = ex.extract_batch(list(frames))
        score = ex.debug_read(0, "score", 4)
        for t, o in enumerate(outs):
            ref = O.extract(weights, frames[t], nf, keep_forward=True)
            np.testing.assert_allclose(score[t], ref["forward"]["score_map"], atol=SCORE_ATOL, rtol=SCORE_RTOL)
            diffs = explain_differences(ref["forward"], ref["kp_xy"], o["kp_xy"], nf)
            assert len(diffs) <= MAX_DIFF_PER_FRAME, diffs
            assert all(e < EPS_MARGIN for _, _, e in diffs), diffs


@pytest.mark.parametrize("mode", ["mma", "ffma"])
def test_fused_conv1_equals_unfused(mode, weights, monkeypatch):
    """conv1a never touches HBM on the default path: it is computed inside the conv1b kernel, straight into the swizzled
    shared-memory slab.  SPFE_CONV1=ffma (conv1a by fp32 FFMA on the CUDA cores) must be bit-identical to the two-kernel
    path (SPFE_CONV1=unfused), whose materialised conv1a activation is checked against the oracle.  The default,
    SPFE_CONV1=mma (conv1a on the tensor core: exact u8 pixels x hi/lo-split fp16 weights, fp32 accumulation), sums in a
    different order, so single fp16 roundings of conv1a may flip: conv1b must agree within 2e-3 of the layer's max with
    fewer than 2 % of its elements differing at all, and the keypoint sets must still coincide almost everywhere."""
    H, W = 120, 136
    frames = synth.make_stream(H, W, 2, seed=7, n_shapes=24)
    monkeypatch.setenv("SPFE_CONV1", mode)
    fused = SPExtractor(800, H, W, WEIGHTS, max_batch=2)
    a = fused.extract_batch(list(frames))
    a1b = fused.debug_read(0, "conv1b", 2)
    with pytest.raises(SpfeError):
        fused.debug_read(0, "conv1a", 2)                                       # never materialised in fused mode
    fused.close()
    monkeypatch.setenv("SPFE_CONV1", "unfused")
    plain = SPExtractor(800, H, W, WEIGHTS, max_batch=2)
    b = plain.extract_batch(list(frames))
    p1b = plain.debug_read(0, "conv1b", 2)
    if mode == "ffma":
        assert np.array_equal(a1b, p1b)
        for x, y in zip(a, b):
            for k in ["kp_xy", "desc", "occ_grid", "dense_dust", "heat"]:
                assert np.array_equal(x[k], y[k]), k
    else:
        d = np.abs(a1b.astype(np.float32) - p1b.astype(np.float32))
        assert d.max() <= 2e-3 * np.abs(p1b.astype(np.float32)).max()
        assert (d > 0).mean() < 0.02
        for x, y in zip(a, b):
            xs = {(int(u), int(v)) for u, v in x["kp_xy"]}
            ys = {(int(u), int(v)) for u, v in y["kp_xy"]}
            assert len(xs & ys) >= 0.98 * len(xs | ys)
            np.testing.assert_allclose(x["dense_dust"], y["dense_dust"], atol=2e-3)
    got = plain.debug_read(0, "conv1a", 2).astype(np.float32)
    for t in range(2):
        ref = O.frontend_forward(weights, frames[t], keep_layers=True)["layers"]["conv1a"].transpose(1, 2, 0)
        assert np.abs(got[t] - ref).max() <= 1e-3 * np.abs(ref).max()          # fp32 compute, fp16 storage
    plain.close()


def test_1080p_2000_keypoints_vs_live_oracle(weights, ex_cache):
    """BASELINE configs[4] geometry (1920x1080, 2000-keypoint budget) against the oracle run live on the same frame:
    integer path bit-exact on the GPU's own maps (cap of 2001 reached), score map / keypoint set / descriptors within
    the stated tolerances, covariance bit-exact on the GPU's own heat map."""
    H, W, nf = 1080, 1920, 2000
    ex = ex_cache(H, W, nf, max_batch=2, emit_cov=True, emit_heat=True)
    frame = synth.make_frame(H, W, seed=17, n_shapes=3600)
    o = ex.extract_batch([frame])[0]
    score, argmax = ex.debug_read(0, "score", 1)[0], ex.debug_read(0, "argmax", 1)[0]
    kp_ref, sc_ref, occ_ref = oracle_nms_on(score, argmax, nf, H, W)
    assert o["n"] == len(kp_ref) == nf + 1
    assert np.array_equal(o["kp_xy"], kp_ref) and np.array_equal(o["kp_score"], sc_ref) and np.array_equal(o["occ_grid"], occ_ref)
    ref = O.extract(weights, frame, nf, keep_forward=True)
    np.testing.assert_allclose(score, ref["forward"]["score_map"], atol=SCORE_ATOL, rtol=SCORE_RTOL)
    gold = {(int(x), int(y)): i for i, (x, y) in enumerate(ref["kp_xy"])}
    mine = {(int(x), int(y)): i for i, (x, y) in enumerate(o["kp_xy"])}
    common = set(gold) & set(mine)
    diffs = explain_differences(ref["forward"], ref["kp_xy"], o["kp_xy"], nf)
    assert len(diffs) <= 3 * MAX_DIFF_PER_FRAME, diffs               # 2001 key points, 5.7x the cells of 752x480
    assert all(e < EPS_MARGIN for _, _, e in diffs), diffs
    cos = np.array([np.dot(o["desc"][mine[k]], ref["desc"][gold[k]]) for k in common])
    assert cos.min() > 1 - COS_TOL
    resp, cov2, cov2_inv = O.covariance(o["heat_inv"], o["kp_xy"])
    assert np.array_equal(o["kp_response"], resp) and np.array_equal(o["cov2"], cov2) and np.array_equal(o["cov2_inv"], cov2_inv)


def test_batch_invariance_and_determinism(ex_cache):
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=4)
    frames = synth.make_stream(H, W, 4, seed=5)
    a = ex.extract_batch(list(frames))
    b = ex.extract_batch(list(frames[::-1]))[::-1]
    c = [ex.extract_batch([f])[0] for f in frames]
    for x, y, z in zip(a, b, c):
        for k in ["kp_xy", "desc", "occ_grid", "dense_dust", "semi_dust", "heat", "heat_inv", "cov2", "kp_response"]:
            assert np.array_equal(x[k], y[k]), k      # position in the batch does not matter (bit-exact)
            assert np.array_equal(x[k], z[k]), k      # neither does the batch size


@pytest.mark.parametrize("H,W,nf", [(480, 752, 800), (480, 752, 100), (1080, 1920, 2000)])
def test_full_size_properties(H, W, nf, ex_cache):
    """Size-independent invariants at BASELINE.json's full sizes (no oracle run needed)."""
    ex = ex_cache(H, W, nf, max_batch=2, emit_cov=False, emit_heat=True)
    frames = synth.make_stream(H, W, 2, seed=13, n_shapes=int(900 * H * W / (752 * 480)))
    outs = ex.extract_batch(list(frames))
    for t, o in enumerate(outs):
        kp = o["kp_xy"].astype(int)
        assert 0 < o["n"] <= nf + 1
        key = kp[:, 1] * W + kp[:, 0]
        assert np.all(np.diff(key) > 0)                                               # raster order, unique
        assert kp[:, 0].min() >= 8 and kp[:, 0].max() < W - 8 and kp[:, 1].min() >= 8 and kp[:, 1].max() < H - 8
        occ = o["occ_grid"]
        assert np.array_equal(occ[kp[:, 1] // 8, kp[:, 0] // 8], np.arange(o["n"]))    # occ_grid <-> keypoint index
        assert (occ >= 0).sum() == o["n"]
        for i in range(0, o["n"], 97):                                                # NMS radius (sampled rows)
            d = np.abs(kp - kp[i]).max(1)
            d[i] = 99
            assert d.min() > 4
        np.testing.assert_allclose(np.linalg.norm(o["desc"], axis=1), 1.0, atol=1e-5)
        assert np.all(o["kp_score"] >= np.float32(0.007))
        # exactness of the integer path at full size, against the oracle NMS on the GPU's own maps
        score, argmax = ex.debug_read(0, "score", 2)[t], ex.debug_read(0, "argmax", 2)[t]
        kp_ref, sc_ref, occ_ref = oracle_nms_on(score, argmax, nf, H, W)
        assert np.array_equal(o["kp_xy"], kp_ref) and np.array_equal(occ, occ_ref)
        assert abs(float(o["heat"].min())) < 1e-6 and abs(float(o["heat"].max()) - 1.0) < 1e-6


@pytest.mark.parametrize("name", ["g480x640", "g480x752", "g480x752_cap"])
def test_matcher_vs_golden(name, golden, ex_cache):
    """Mutual-NN on the GOLDEN descriptors must reproduce the golden (cv2.BFMatcher-verified) match list."""
    g = golden(name)
    ex = ex_cache(64, 64, 800, emit_heat=False, emit_cov=False)
    q, t = g["f1_desc"].astype(np.float32), g["f0_desc"].astype(np.float32)
    ref, rdist, rsec = O.match_mutual_nn(q, t)
    got, dist = ex.match(q, t)
    assert np.array_equal(got, ref)
    np.testing.assert_allclose(dist, rdist, atol=2e-6)
    assert (got >= 0).sum() > 0.6 * len(q)


def test_matcher_random_and_edges(ex_cache):
    ex = ex_cache(64, 64, 800, emit_heat=False, emit_cov=False)
    rng = np.random.RandomState(0)
    for nq, nt in [(801, 801), (2001, 1777), (5, 900), (1, 1), (130, 64), (64, 0), (0, 10)]:
        q = rng.randn(nq, 256).astype(np.float32)
        q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-9)
        t = np.concatenate([q[rng.permutation(nq)[: min(nq, nt)]] + 0.05 * rng.randn(min(nq, nt), 256).astype(np.float32),
                            rng.randn(max(nt - nq, 0), 256).astype(np.float32)])[:nt] if nt else np.zeros((0, 256), np.float32)
        if nt:
            t /= np.linalg.norm(t, axis=1, keepdims=True)
        got, dist = ex.match(q, t)
        ref, rdist, _ = O.match_mutual_nn(q, t)
        assert np.array_equal(got, ref), (nq, nt)
        if nq and nt:
            np.testing.assert_allclose(dist, rdist, atol=2e-6)
    q = rng.randn(300, 256).astype(np.float32)
    got, dist = ex.match(q, q)                              # self-match = identity, distance 0
    assert np.array_equal(got, np.arange(300)) and np.all(dist == 0)
    dup = np.concatenate([q[:1], q[:1], q[1:5]])            # duplicate train rows: the first index wins
    got, _ = ex.match(q[:1], dup)
    assert got.tolist() == [0]


def test_search_by_brute_force_overloads(golden, ex_cache):
    """Both SearchByBruteForce overloads (sp_matcher.cpp:1642-1674, sp_matcher_loop.cpp:334-376) on plain arrays."""
    g = golden("g480x640")
    ex = ex_cache(64, 64, 800, emit_heat=False, emit_cov=False)
    m = SPMatcher(ex)
    d1, d2 = g["f0_desc"].astype(np.float32), g["f1_desc"].astype(np.float32)
    rng = np.random.RandomState(2)
    v1, v2 = rng.rand(len(d1)) < 0.7, rng.rand(len(d2)) < 0.8
    out = m.SearchByBruteForce(d1, v1, d2)
    idx_t = np.flatnonzero(v1)
    ref, _, _ = O.match_mutual_nn(d2, d1[idx_t])
    exp = np.where(ref >= 0, idx_t[np.maximum(ref, 0)], -1)
    assert np.array_equal(out, exp) and not np.any(~v1[out[out >= 0]])
    out2, n2 = m.SearchByBruteForce(d1, v1, d2, v2)
    idx_q = np.flatnonzero(v2)
    ref2, _, _ = O.match_mutual_nn(d2[idx_q], d1[idx_t])
    exp2 = np.full(len(d1), -1, np.int64)
    exp2[idx_t[ref2[ref2 >= 0]]] = idx_q[ref2 >= 0]
    assert np.array_equal(out2, exp2) and n2 == int((ref2 >= 0).sum())


def test_stream_matching_match_prev(ex_cache):
    """SPFE_MATCH_PREV: frame t vs t-1 inside the pipeline == oracle matcher on the extractor's own descriptors,
    including across batch boundaries (carry) and after reset_stream."""
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=3, match_prev=True, emit_heat=False, emit_cov=False)
    frames = synth.make_stream(H, W, 6, seed=17)
    ex.reset_stream(0)
    outs = ex.extract_batch(list(frames[:3])) + ex.extract_batch(list(frames[3:]))
    assert outs[0]["n_prev"] == 0 and np.all(outs[0]["match_prev"] == -1)
    for t in range(1, 6):
        assert outs[t]["n_prev"] == outs[t - 1]["n"]
        ref, rdist, _ = O.match_mutual_nn(outs[t]["desc"], outs[t - 1]["desc"])
        assert np.array_equal(outs[t]["match_prev"], ref), t
        np.testing.assert_allclose(outs[t]["match_dist"], rdist, atol=2e-6)
        assert (ref >= 0).mean() > 0.5
    ex.reset_stream(0)
    again = ex.extract_batch(list(frames[3:]))
    assert again[0]["n_prev"] == 0 and np.all(again[0]["match_prev"] == -1)
    assert np.array_equal(again[1]["match_prev"], outs[4]["match_prev"])


def test_covariance_vs_oracle_on_same_heat(ex_cache):
    """computeCovariance is order- and comparison-dependent (shared visited map, FIFO floods).  The device version
    (parallel lone floods, ordered conflict-resolution rounds, sequential remainder; csrc/cov.cuh) must be bit-exact against the oracle's
    sequential C restatement run on the GPU's own heat_inv and keypoints."""
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=4)
    o = ex.extract(synth.make_frame(H, W, seed=23))
    resp, cov2, cov2_inv = O.covariance(o["heat_inv"], o["kp_xy"])
    assert np.array_equal(o["kp_response"], resp)
    assert np.array_equal(o["cov2"], cov2) and np.array_equal(o["cov2_inv"], cov2_inv)
    assert np.all(o["cov2"] >= 1.0)
    heat, heat_inv, _, _ = O.to_heat(ex.debug_read(0, "heat_log", 1)[0])
    assert np.array_equal(o["heat"], heat) and np.array_equal(o["heat_inv"], heat_inv)     # to_heat: bit-exact


@pytest.mark.parametrize("force", [1, 3])
def test_covariance_fallback_paths(force, monkeypatch):
    """The paths real frames rarely take: SPFE_COV_FORCE=1 sends every lone flood through the big-limit kernel (and the
    big-limit blocks of the rounds), =3 additionally fails those, so the whole frame is replayed sequentially."""
    H, W = 240, 320
    monkeypatch.setenv("SPFE_COV_FORCE", str(force))
    ex = SPExtractor(800, H, W, WEIGHTS, max_batch=2)
    frames = synth.make_stream(H, W, 2, seed=23, n_shapes=300)
    for o in ex.extract_batch(list(frames)):
        resp, cov2, cov2_inv = O.covariance(o["heat_inv"], o["kp_xy"])
        assert o["n"] > 100 and np.array_equal(o["kp_response"], resp)
        assert np.array_equal(o["cov2"], cov2) and np.array_equal(o["cov2_inv"], cov2_inv)
    ctr = ex.debug_read(0, "cov_counters", 1)[0]
    replayed = ex.debug_read(0, "cov_replayed", 2)
    assert ctr[1] > 100                                        # every flood went to the big list
    assert (replayed[:, 0].min() > 100) == (force == 3)        # ... and, with force = 3, to the sequential replay
    ex.close()


def test_covariance_dense_full_size_and_device_only(ex_cache):
    """Dense 752x480 scenes (hundreds of keypoints 5 px apart -> overlapping floods that must be replayed in order),
    a batch of them, and the mode where the heat maps never leave the device."""
    H, W = 480, 752
    ex = ex_cache(H, W, 800, max_batch=3)
    frames = synth.make_stream(H, W, 3, seed=77, n_shapes=1500)
    outs = ex.extract_batch(list(frames))
    qlen = ex.debug_read(0, "cov_qlen", 3)
    done = ex.debug_read(0, "cov_done", 3)
    ctr = ex.debug_read(0, "cov_counters", 1)[0]
    assert ctr[2] > 0                                     # some keypoints conflicted and went through the ordered rounds
    for t, o in enumerate(outs):
        resp, cov2, cov2_inv = O.covariance(o["heat_inv"], o["kp_xy"])
        assert np.array_equal(o["kp_response"], resp), t
        assert np.array_equal(o["cov2"], cov2) and np.array_equal(o["cov2_inv"], cov2_inv), t
        assert np.all(qlen[t, :o["n"]] > 0) and np.all(done[t, :o["n"]] == 1)             # every flood completed
    for _ in range(3):                                    # consecutive batches (odd and even claim epochs): still exact, and the
        again = ex.extract_batch(list(frames))            # parallel rounds resolve every conflict (nothing left for the sequential replay)
        assert ex.debug_read(0, "cov_replayed", 3)[:, 0].sum() == 0
        for o, o2 in zip(outs, again):
            assert np.array_equal(o["cov2"], o2["cov2"]) and np.array_equal(o["kp_response"], o2["kp_response"])
    dev = ex_cache(H, W, 800, max_batch=3, emit_heat=False, emit_cov=True)
    outs2 = dev.extract_batch(list(frames))
    for o, o2 in zip(outs, outs2):
        assert "heat" not in o2
        for k in ["kp_xy", "kp_response", "cov2", "cov2_inv", "desc"]:
            assert np.array_equal(o[k], o2[k]), k


def test_operator_call_mirrors_reference(ex_cache):
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=4)
    img = synth.make_frame(H, W, seed=3)
    kps, desc = ex(img, None)
    assert kps.shape[1] == 3 and desc.shape == (len(kps), 256) and desc.dtype == np.float32
    assert ex.occ_grid_.shape == (H // 8, W // 8) and ex.occ_grid_.dtype == np.int16
    assert ex.dense_dust_.shape == (H // 8, W // 8) and ex.heat_.shape == (H, W) and ex.heat_inv_.shape == (H, W)
    assert len(ex.getCov2Inv()) == len(kps) == len(ex.getCov())
    assert ex.GetLevels() == 1 and ex.GetScaleFactor() == 1.0 and ex.GetScaleFactors() == [1.0]
    with pytest.raises(RuntimeError, match="input image is empty"):
        ex(np.zeros((0, 0), np.uint8))
    with pytest.raises(SpfeError):
        ex(np.zeros((H, W + 8), np.uint8))
    sub = np.zeros((H, 2 * W), np.uint8)
    sub[:, :W] = img
    k2, d2 = ex(sub[:, :W])                                      # non-contiguous rows (row_stride > width)
    assert np.array_equal(k2, kps) and np.array_equal(d2, desc)


def test_blank_and_saturated_frames(ex_cache):
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=4)
    for val in (0, 128, 255):
        o = ex.extract(np.full((H, W), val, np.uint8))
        assert o["n"] == 0 and o["desc"].shape == (0, 256) and np.all(o["occ_grid"] == -1)


@pytest.mark.parametrize("H,W", [(8, 8), (16, 24), (24, 8), (40, 72)])
def test_tiny_frames(H, W, weights):
    """Smallest legal geometries (one cell and up): single work items, clamped CTA pairs, TMA boxes larger than the
    tensors.  The integer path must stay bit-exact on the GPU's own maps and the score map within tolerance."""
    ex = SPExtractor(50, H, W, WEIGHTS, max_batch=3, match_prev=True)
    frames = synth.make_stream(H, W, 3, seed=3, n_shapes=6)
    outs = ex.extract_batch(list(frames))
    for t, o in enumerate(outs):
        score, argmax = ex.debug_read(0, "score", 3)[t], ex.debug_read(0, "argmax", 3)[t]
        kp_ref, sc_ref, occ_ref = oracle_nms_on(score, argmax, 50, H, W)
        assert o["n"] == len(kp_ref) and np.array_equal(o["kp_xy"], kp_ref) and np.array_equal(o["occ_grid"], occ_ref)
        if min(H, W) >= 16:                                   # (the reference's .squeeze() calls need hc, wc >= 2; so does its restatement)
            fwd = O.frontend_forward(weights, frames[t])
            np.testing.assert_allclose(score, fwd["score_map"], atol=SCORE_ATOL, rtol=SCORE_RTOL)
            np.testing.assert_allclose(o["dense_dust"], fwd["dense_dust"], atol=1e-2)
        assert o["heat"].shape == (H, W)                      # (a constant heat map normalises to 0/0 = NaN, as in the reference)
    ex.close()


def test_global_keypoint_budget_cut():
    """The optional multi-GPU keypoint budget end to end on one rank: histogram -> common score cut (sharding.py, one
    all-reduce when a process group exists) -> spfe_set_score_threshold -> fewer keypoints, NMS still exact at the new cut."""
    from sp_orb_slam_b200 import sharding
    H, W = 240, 320
    ex = SPExtractor(800, H, W, WEIGHTS, emit_heat=False, emit_cov=False)
    frame = synth.make_frame(H, W, seed=5, n_shapes=200)
    base = ex.extract(frame)
    score, argmax = ex.debug_read(0, "score", 1)[0], ex.debug_read(0, "argmax", 1)[0]
    cut = sharding.global_keypoint_budget(score[score >= O.SCORE_THRESH], budget=base["n"] // 3)
    assert cut > O.SCORE_THRESH
    ex.set_score_threshold(cut)
    o = ex.extract(frame)
    mask = score >= np.float32(cut)
    cy, cx = np.nonzero(mask)
    pts = np.stack([cx * 8 + argmax[mask] % 8, cy * 8 + argmax[mask] // 8], 1).astype(np.float32)
    order = O.sort_desc(score[mask])
    sel, occ = O.nms(pts[order], 800, W, H)
    assert 0 < o["n"] < base["n"] and np.array_equal(o["kp_xy"], pts[order][sel]) and np.array_equal(o["occ_grid"], occ)
    with pytest.raises(SpfeError):
        ex.set_score_threshold(2.0)
    ex.close()


def test_slots_pipeline(ex_cache):
    H, W = 240, 320
    ex = ex_cache(H, W, 800, max_batch=2, num_slots=3, emit_heat=False, emit_cov=False)
    frames = synth.make_stream(H, W, 6, seed=41)
    for s in range(3):
        ex.submit(s, list(frames[2 * s: 2 * s + 2]))
    with pytest.raises(SpfeError):
        ex.submit(0, list(frames[:2]))                           # slot busy until waited
    res = [o for s in range(3) for o in ex.wait(s, 2)]
    single = SPExtractor(800, H, W, WEIGHTS, emit_heat=False, emit_cov=False)
    for f, o in zip(frames, res):
        r = single.extract(f)
        assert np.array_equal(r["kp_xy"], o["kp_xy"]) and np.array_equal(r["desc"], o["desc"])
    single.close()
    assert ex.launch_count() > 0
    # zero-staging entry: frames DMA'd straight from the caller's page-locked buffer give the same results
    pinned = ex.pinned_frames(6)
    pinned[:] = frames
    for s in range(3):
        ex.submit_pinned(s, pinned[2 * s: 2 * s + 2])
    res2 = [o for s in range(3) for o in ex.wait(s, 2)]
    for o, o2 in zip(res, res2):
        assert np.array_equal(o["kp_xy"], o2["kp_xy"]) and np.array_equal(o["desc"], o2["desc"])


def test_native_stream_bench():
    """The C++ throughput harness over the C ABI (no Python in the loop): pinned frames in, pipelined slots, full outputs."""
    import json
    from sp_orb_slam_b200 import build
    exe = build.build_stream_bench()
    r = subprocess.run([exe, WEIGHTS, "240", "320", "4", "3", "12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["frames"] == 48 and d["frames_per_s"] > 100 and d["keypoints_per_frame"] > 50 and d["matches_per_frame"] > 20
    assert d["launches"] > 0


def test_cpp_shim_selftest(tmp_path, ex_cache):
    """The C++ drop-in classes, driven like Frame::ExtractORB / trackReferenceKeyFrameANN drive the reference."""
    from sp_orb_slam_b200 import build
    exe = build.build_shim()
    H, W = 240, 320
    frames = synth.make_stream(H, W, 2, seed=19)
    for i, f in enumerate(frames):
        f.tofile(tmp_path / f"f{i}.raw")
    pre = str(tmp_path / "out")
    r = subprocess.run([exe, WEIGHTS, str(H), str(W), str(tmp_path / "f0.raw"), str(tmp_path / "f1.raw"), pre],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'throws runtime_error("input image is empty"): yes' in r.stdout
    assert "throughput mode (lazy heat, fp16 descriptors): identical" in r.stdout
    ex = ex_cache(H, W, 800, max_batch=4)
    a, b = ex.extract(frames[0]), ex.extract(frames[1])
    kps = np.loadtxt(pre + "_kps_a.txt").reshape(-1, 5)
    assert np.array_equal(kps[:, :2], a["kp_xy"])
    np.testing.assert_allclose(kps[:, 2], a["kp_response"], rtol=1e-6)
    np.testing.assert_allclose(kps[:, 3:], a["cov2_inv"], rtol=1e-6)
    valid = np.array([(i % 3 != 2) and (i % 10 != 9) for i in range(a["n"])])
    exp = SPMatcher(ex).SearchByBruteForce(a["desc"], valid, b["desc"])
    got = np.loadtxt(pre + "_kf_frame.txt", dtype=np.int64).reshape(-1)
    assert np.array_equal(got, exp)
    # guided searches through the C++ templates == the Python mirror (both over spfe_search_guided; the device result is
    # checked against the oracle in tests/test_guided.py)
    i = np.arange(a["n"])
    frame_b = dict(desc=b["desc"], kp_un=b["kp_xy"], occ_grid=b["occ_grid"])
    cos = np.where(i % 2 == 1, 0.9995, 0.99).astype(np.float32)
    in_view, observed = (i % 7 != 6).astype(np.uint8), (i % 5 != 4).astype(np.uint8)
    M = SPMatcher(ex)
    mp2kp, n3 = M.SearchByProjectionMapPoints(frame_b, a["desc"], a["kp_xy"], cos, th=3.0, th_dist=0.7, in_view=in_view, observed=observed)
    dust2kp, n4 = M.DustAssociate(frame_b, a["desc"], ((a["kp_xy"] - 3.5) / 8.0).astype(np.float32), in_view=in_view)
    assert f"SearchByProjection(F,MPs) {n3} matches; dust association {n4} matches" in r.stdout and n3 > 50 and n4 > 50
    exp_g = -np.ones((b["n"], 2), np.int64)
    for col, q2kp in enumerate((mp2kp, dust2kp)):
        for mp, kp in enumerate(q2kp):                       # applied in map-point order: later assignments overwrite
            if kp >= 0:
                exp_g[kp, col] = mp
    assert np.array_equal(np.loadtxt(pre + "_guided.txt", dtype=np.int64).reshape(-1, 2), exp_g)
