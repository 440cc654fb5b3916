"""CPU tests of the oracle itself: against the committed golden vectors (minted
from the reference's own compiled SPFrontend), against OpenCV for the
third-party arithmetic, and against brute-force restatements."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, WEIGHTS
from oracle import sp_oracle as O, weights as OW
from sp_orb_slam_b200 import synth

REF_PT = "/root/reference/orb_ros/data/models/superpoint.pt"


def test_weights_fixture_matches_reference_archive(weights):
    if not os.path.exists(REF_PT):
        pytest.skip("reference checkout not present (GPU box)")
    ref = OW.read_legacy_pt(REF_PT)
    assert list(ref) == list(weights)
    for k in ref:
        assert np.array_equal(ref[k], weights[k]), k
    assert sum(v.size for v in ref.values()) == 1300865


def test_synth_is_deterministic():
    a = synth.make_stream(64, 96, 2, seed=3)
    b = synth.make_stream(64, 96, 2, seed=3)
    assert a.dtype == np.uint8 and a.shape == (2, 64, 96) and np.array_equal(a, b)
    assert not np.array_equal(a[0], a[1])


@pytest.mark.parametrize("name", ["g120x160", "g240x320_ragged", "g480x640"])
def test_oracle_reproduces_golden(name, weights, golden):
    """Golden = reference's own SPFrontend (libtorch C++) + post-processing; oracle = torch-python restatement."""
    g = golden(name)
    nf = int(g["nfeatures"])
    for t in range(2):
        o = O.extract(weights, g["frames"][t], nf, keep_forward=True)
        f = o["forward"]
        assert np.array_equal(f["pixels_in"].astype(np.int16), g[f"f{t}_cand_pixels"])      # candidate set: exact
        np.testing.assert_allclose(f["score"], g[f"f{t}_cand_score"], rtol=0, atol=2e-5)
        assert o["n"] == int(g[f"f{t}_n"])
        assert np.array_equal(o["kp_xy"].astype(np.int16), g[f"f{t}_kp_xy"])                # keypoints: exact
        assert np.array_equal(o["occ_grid"], g[f"f{t}_occ_grid"])
        cos = np.sum(o["desc"] * g[f"f{t}_desc"].astype(np.float32), 1)
        assert cos.min() > 1 - 1e-4
        np.testing.assert_allclose(o["heat"], g[f"f{t}_heat_q"] / 255.0, atol=0.5 / 255 + 1e-4)
        np.testing.assert_allclose(o["cov2"], g[f"f{t}_cov2"], rtol=2e-2, atol=2e-2)


def _greedy_nms_python(pts, nf, W, H, border=8, r=4):
    """Independent brute-force statement of sp_extractor.cpp:161-250."""
    alive = np.ones(len(pts), bool)
    kept = []
    for i in range(len(pts)):
        if not alive[i]:
            continue
        d = np.abs(pts - pts[i]).max(1)
        alive &= ~(d <= r)
        kept.append(i)
        if len(kept) > nf:
            break
    kept = [i for i in kept if border <= pts[i, 0] < W - border and border <= pts[i, 1] < H - border]
    kept.sort(key=lambda i: (pts[i, 1], pts[i, 0]))
    return np.array(kept, np.int32)


@pytest.mark.parametrize("seed", range(6))
def test_oracle_nms_vs_bruteforce(seed):
    rng = np.random.RandomState(seed)
    H, W = 96, 128
    hc, wc = H // 8, W // 8
    cells = np.flatnonzero(rng.rand(hc * wc) < 0.8)
    pos = rng.randint(0, 64, len(cells))
    pts = np.stack([(cells % wc) * 8 + pos % 8, (cells // wc) * 8 + pos // 8], 1).astype(np.float32)
    score = rng.permutation(len(cells)).astype(np.float32)
    order = O.sort_desc(score)
    assert np.array_equal(score[order], np.sort(score)[::-1])
    nf = [1000, 20, 5][seed % 3]
    sel, occ = O.nms(pts[order], nf, W, H)
    ref = _greedy_nms_python(pts[order], nf, W, H)
    assert np.array_equal(sel, ref)
    assert (occ >= 0).sum() == len(sel)
    for k, i in enumerate(sel):
        x, y = pts[order][i]
        assert occ[int(y) // 8, int(x) // 8] == k


def test_sort_desc_ties_keep_raster_order():
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], np.float32)
    assert O.sort_desc(s).tolist() == [1, 3, 0, 2, 4]


def test_to_heat_properties():
    rng = np.random.RandomState(0)
    x = np.log(np.clip(rng.rand(40, 56).astype(np.float32), 1e-3, None))
    heat, heat_inv, mn, mx = O.to_heat(x)
    assert mn == float((-x).min()) and mx == float((-x).max())
    np.testing.assert_allclose(heat + heat_inv, 1.0, atol=1e-6)
    assert heat.min() >= -1e-6 and heat.max() <= 1 + 1e-6
    np.testing.assert_allclose(heat, (-x - mn) / (mx - mn), atol=1e-6)


def test_to_heat_vs_opencv_convert_scale():
    """to_heat (sp_extractor.cpp:461-474) is OpenCV arithmetic: a MatExpr folded into one convertTo(alpha, beta) with
    alpha = 1 / (max - min), beta = -min * alpha computed in double and narrowed to float.  cv2.normalize(NORM_MINMAX)
    runs the same convertTo with the same alpha / beta, so it pins that derivation.  This cv2 build (4.x, AVX2) fuses
    x * alpha + beta into one FMA; OpenCV 3.2 (the reference's, SSE2) rounds the product first, which is what the oracle
    and the device do -- the two differ by at most one ulp, and the oracle's alpha / beta under an emulated FMA reproduce
    cv2 exactly (up to the emulation's own double rounding)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for _ in range(3):
        x = np.log(np.clip(rng.rand(120, 160).astype(np.float32) ** 3, 1e-3, None)).astype(np.float32)
        heat, _, mn, mx = O.to_heat(x)
        img = (-x).astype(np.float32)
        ref = cv2.normalize(img, None, alpha=0, beta=1, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
        assert (mn, mx) == cv2.minMaxLoc(img)[:2]
        assert np.abs(heat - ref).max() <= 2.0 ** -23                       # one ulp below 1.0
        a, b = np.float32(1.0 / (mx - mn)), np.float32(-mn * (1.0 / (mx - mn)))
        assert np.array_equal(heat, img * a + b)                            # the oracle = separate multiply and add
        fma = (img.astype(np.float64) * np.float64(a) + np.float64(b)).astype(np.float32)
        assert (fma != ref).mean() < 1e-3


def test_covariance_simple_peak():
    h = np.zeros((16, 16), np.float32)
    h[8, 8], h[8, 7], h[8, 9], h[7, 8], h[9, 8] = 1.0, 0.5, 0.5, 0.25, 0.25
    resp, cov2, cov2_inv = O.covariance(h, np.array([[8, 8]], np.float32))
    assert resp[0] == 1.0
    np.testing.assert_allclose(cov2[0], [max(1.0, 1.0 / 2.5), 1.0])     # x: (0.5+0.5)/2.5 = 0.4 -> floored to 1
    np.testing.assert_allclose(cov2_inv[0], 1.0 / cov2[0])


def test_matcher_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(1)
    for nq, nt in [(300, 280), (50, 400), (1, 1), (64, 3)]:
        q = rng.randn(nq, 256).astype(np.float32)
        t = np.concatenate([q[: min(nq, nt)] + 0.3 * rng.randn(min(nq, nt), 256).astype(np.float32),
                            rng.randn(max(nt - nq, 0), 256).astype(np.float32)])[:nt]
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        t /= np.linalg.norm(t, axis=1, keepdims=True)
        q2t, dist, _ = O.match_mutual_nn(q, t)
        ref = -np.ones(nq, np.int32)
        for m in cv2.BFMatcher(cv2.NORM_L2, True).match(q, t):
            ref[m.queryIdx] = m.trainIdx
            assert abs(m.distance - dist[m.queryIdx]) < 1e-5
        assert np.array_equal(ref, q2t)
    assert abs(O.l2(q[0], t[0]) - cv2.norm(q[0], t[0], cv2.NORM_L2)) < 1e-6


def test_matcher_edge_cases():
    q = np.eye(4, 256, dtype=np.float32)
    q2t, _, _ = O.match_mutual_nn(q, np.zeros((0, 256), np.float32))
    assert q2t.tolist() == [-1] * 4
    q2t, d, _ = O.match_mutual_nn(q, q[::-1].copy())
    assert q2t.tolist() == [3, 2, 1, 0] and np.all(d == 0)
    dup = np.concatenate([q[:1], q[:1]])          # duplicate train rows: first index wins
    q2t, _, _ = O.match_mutual_nn(q[:1], dup)
    assert q2t.tolist() == [0]


def test_reference_frontend_pins_oracle(weights):
    """oracle/_ref = the reference's own SPFrontend compiled here; the restatement must agree with it."""
    from oracle import ref_frontend as R
    if not R.available():
        pytest.skip("oracle/_ref not built (run oracle/ref_build.sh where /root/reference exists)")
    img = synth.make_frame(120, 160, seed=11)
    r, o = R.forward(weights, img), O.frontend_forward(weights, img)
    assert np.array_equal(r["pixels_in"], o["pixels_in"])
    for k, tol in [("score", 2e-5), ("semi_dust", 2e-4), ("dense_dust", 2e-5), ("heat_log", 2e-4), ("desc_sampled", 5e-6)]:
        np.testing.assert_allclose(r[k], o[k], rtol=0, atol=tol, err_msg=k)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixture_self_consistency(name, golden):
    g = golden(name)
    nf = int(g["nfeatures"])
    H, W = g["frames"].shape[1:]
    for t in range(2):
        kp = g[f"f{t}_kp_xy"].astype(int)
        assert len(kp) == int(g[f"f{t}_n"]) <= nf + 1
        key = kp[:, 1] * W + kp[:, 0]
        assert np.all(np.diff(key) > 0)                                   # raster order
        assert kp[:, 0].min() >= 8 and kp[:, 0].max() < W - 8 and kp[:, 1].min() >= 8 and kp[:, 1].max() < H - 8
        occ = g[f"f{t}_occ_grid"]
        assert np.array_equal(occ[kp[:, 1] // 8, kp[:, 0] // 8], np.arange(len(kp)))
        d = np.abs(kp[:, None, :] - kp[None, :, :]).max(-1) + 100 * np.eye(len(kp), dtype=int)
        assert d.min() > 4                                                # NMS radius
    assert WEIGHTS.endswith(".spw")


def _ref_post():
    from oracle import ref_post as RP
    if not RP.available():
        pytest.skip("oracle/_ref/libsppost_ref.so not built (run oracle/ref_build.sh where /root/reference exists)")
    return RP


@pytest.mark.parametrize("seed", range(8))
def test_reference_nms_pins_oracle(seed):
    """The reference's OWN nms() (sp_extractor.cpp:161-250, compiled verbatim into oracle/_ref) against the C
    restatement: identical keypoints, raster order, occ_grid and gathered descriptor rows -- dense and sparse
    candidate sets, the nfeatures cap (kept count > nf stops the loop), points inside the border band, tiny images."""
    RP = _ref_post()
    rng = np.random.RandomState(seed)
    H, W = [(96, 128), (480, 752), (64, 64), (240, 320)][seed % 4]
    hc, wc = H // 8, W // 8
    cells = np.flatnonzero(rng.rand(hc * wc) < [0.9, 0.35, 1.0, 0.6][seed % 4])
    pos = rng.randint(0, 64, len(cells))
    pts = np.stack([(cells % wc) * 8 + pos % 8, (cells // wc) * 8 + pos // 8], 1).astype(np.float32)
    order = O.sort_desc(rng.permutation(len(cells)).astype(np.float32))
    pts = pts[order]
    desc = rng.randn(len(pts), 256).astype(np.float32)
    for nf in (800, 25, 3):
        sel, occ = O.nms(pts, nf, W, H)
        kps_ref, occ_ref, desc_ref = RP.nms(pts, desc, nf, W, H)
        assert np.array_equal(pts[sel], kps_ref)
        assert np.array_equal(occ, occ_ref)
        assert np.array_equal(desc[sel], desc_ref)
    k0, o0, _ = RP.nms(np.zeros((0, 2), np.float32), None, 800, W, H)          # no candidates at all
    s0, oc0 = O.nms(np.zeros((0, 2), np.float32), 800, W, H)
    assert len(k0) == 0 and len(s0) == 0 and np.array_equal(o0, oc0) and np.all(o0 == -1)


@pytest.mark.parametrize("seed", range(6))
def test_reference_covariance_pins_oracle(seed):
    """The reference's OWN computeCovariance() (sp_extractor.cpp:252-340) against the C restatement, bit for bit:
    smooth maps with overlapping basins (the visited map is shared between keypoints, so the keypoint ORDER matters),
    plateaus and exact zeros (the `0 < heat < parent` test), keypoints on the image edge (the `xx > 0` / `yy > 0`
    boundary tests never visit row / column 0)."""
    RP = _ref_post()
    rng = np.random.RandomState(100 + seed)
    H, W = [(64, 96), (120, 160), (48, 48)][seed % 3]
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    heat = np.zeros((H, W), np.float32)
    n = [12, 60, 5][seed % 3]
    cx, cy = rng.randint(0, W, n), rng.randint(0, H, n)
    for x, y in zip(cx, cy):
        s = rng.uniform(0.8, 1.6)      # narrow peaks, cut to 0 below 5 %: the flood re-pushes a pixel once per uphill path
        blob = np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * s * s))           # (duplicates, :296-300), wide smooth basins blow up
        heat = np.maximum(heat, (rng.uniform(0.3, 1.0) * np.where(blob > 0.05, blob, 0.0)).astype(np.float32))
    if seed % 2:
        heat = np.round(heat * 16) / 16                                       # plateaus and exact zeros
    heat = heat.astype(np.float32)
    kps = np.stack([cx, cy], 1).astype(np.float32)
    kps = kps[np.lexsort((kps[:, 0], kps[:, 1]))]                              # raster order, as nms() emits them
    for k in (kps, kps[::-1].copy()):                                          # and a different order: results change, parity must not
        r0, c0, i0 = O.covariance(heat, k)
        r1, c1, i1 = RP.covariance(heat, k)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1) and np.array_equal(i0, i1)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reference_post_pins_golden(name, golden):
    """The committed golden fixtures re-derived with the reference's own nms(): the golden candidate lists, sorted as
    SPExtractor::operator() sorts them (:489-498), give the golden keypoints and occ_grid -- including the cap case."""
    RP = _ref_post()
    g = golden(name)
    H, W = g["frames"].shape[1:]
    nf = int(g["nfeatures"])
    for t in range(2):
        pts = g[f"f{t}_cand_pixels"].astype(np.float32).T.reshape(-1, 2) if g[f"f{t}_cand_pixels"].shape[0] == 2 else g[f"f{t}_cand_pixels"].astype(np.float32)
        order = O.sort_desc(g[f"f{t}_cand_score"])
        kps, occ, _ = RP.nms(pts[order], None, nf, W, H)
        assert np.array_equal(kps.astype(np.int16), g[f"f{t}_kp_xy"])
        assert np.array_equal(occ, g[f"f{t}_occ_grid"])


@pytest.mark.parametrize("seed,n1,n2", [(1, 300, 320), (2, 801, 801), (3, 40, 7), (4, 5, 60)])
def test_reference_bruteforce_pins_mirror_logic(seed, n1, n2):
    """The reference's OWN SearchByBruteForce overloads (sp_matcher.cpp:1642-1674, sp_matcher_loop.cpp:334-376), compiled
    verbatim into oracle/_ref around a BFMatcher stand-in: which rows enter the matcher and how matches map back.
    (KeyFrame*, Frame&) keeps key-frame rows with a map point that is not bad and all frame rows; (KeyFrame*, KeyFrame*)
    keeps rows with a map point on both sides -- bad ones included -- and writes vpMatches12[row of KF1] = map point of
    the matched KF2 row.  Expected values = the index arithmetic of sp_orb_slam_b200.SPMatcher.SearchByBruteForce and of
    cpp/sp_matcher.h on top of the oracle's mutual-NN."""
    from oracle import ref_post as RP
    if not RP.bf_available():
        pytest.skip("oracle/_ref/libspbf_ref.so not built (run oracle/ref_build.sh where /root/reference exists)")
    rng = np.random.RandomState(seed)
    d1 = rng.randn(n1, 256).astype(np.float32); d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    src = rng.randint(0, n1, n2)
    d2 = (d1[src] + 0.05 * rng.randn(n2, 256)).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    has1, bad1 = (rng.rand(n1) < 0.7).astype(np.uint8), (rng.rand(n1) < 0.15).astype(np.uint8)
    has2, bad2 = (rng.rand(n2) < 0.75).astype(np.uint8), (rng.rand(n2) < 0.15).astype(np.uint8)
    # (KeyFrame*, Frame&)
    idx_t = np.flatnonzero((has1 == 1) & (bad1 == 0))
    q2t, _, _ = O.match_mutual_nn(d2, d1[idx_t])
    exp = np.where(q2t >= 0, idx_t[np.maximum(q2t, 0)] if len(idx_t) else -1, -1)
    got = RP.bruteforce_kf_frame(d1, has1, bad1, d2)
    assert np.array_equal(got, exp) and (n1 < 100 or (exp >= 0).sum() > 10)
    # (KeyFrame*, KeyFrame*)
    idx_t, idx_q = np.flatnonzero(has1 == 1), np.flatnonzero(has2 == 1)
    q2t, _, _ = O.match_mutual_nn(d2[idx_q], d1[idx_t])
    exp = -np.ones(n1, np.int64)
    hit = np.flatnonzero(q2t >= 0)
    exp[idx_t[q2t[hit]]] = idx_q[hit]
    got, cnt = RP.bruteforce_kf_kf(d1, has1, bad1, d2, has2, bad2)
    assert np.array_equal(got, exp) and cnt == len(hit)


@pytest.mark.parametrize("name", ["g120x160", "g480x752"])
def test_reference_composed_extractor_reproduces_golden(name, weights, golden):
    """SPExtractor::operator() re-assembled from the reference's OWN compiled pieces -- SPFrontend::forward (libspref), the
    sort (:489-498, oracle / cv2-pinned), nms() and computeCovariance() (libsppost_ref) -- on the golden frames: keypoints,
    occ_grid, responses and covariances of the committed fixtures come out bit for bit (the fixtures were minted with the
    C restatements in place of the last two)."""
    from oracle import ref_frontend as R
    RP = _ref_post()
    if not R.available():
        pytest.skip("oracle/_ref not built")
    g = golden(name)
    nf = int(g["nfeatures"])
    H, W = g["frames"].shape[1:]
    for t in range(2):
        r = R.forward(weights, g["frames"][t])
        pts = r["pixels_in"].T.astype(np.float32) if r["pixels_in"].shape[0] == 2 else r["pixels_in"].astype(np.float32)
        order = O.sort_desc(r["score"])
        kps, occ, _ = RP.nms(pts[order], None, nf, W, H)
        assert np.array_equal(kps.astype(np.int16), g[f"f{t}_kp_xy"]) and np.array_equal(occ, g[f"f{t}_occ_grid"])
        _, heat_inv, _, _ = O.to_heat(r["heat_log"])
        resp, cov2, _ = RP.covariance(heat_inv, kps)
        assert np.array_equal(resp, g[f"f{t}_response"]) and np.array_equal(cov2, g[f"f{t}_cov2"])
