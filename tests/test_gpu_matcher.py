"""GPU tests of the tensor-core matcher behind SPMatcher::SearchByBruteForce (spfe_match_mutual_nn / spfe_match_knn2 and
the descriptor-set forms): the fp16 GEMM only NOMINATES, the answer must be the exact fp32 one -- also where the
nomination is blind (near-ties below the fp16 resolution, several of them inside one 256-column block, duplicates
across block boundaries).  Index results are compared bit for bit with the oracle (pinned to cv2.BFMatcher)."""
import numpy as np
import pytest

from conftest import WEIGHTS
from oracle import sp_oracle as O
from sp_orb_slam_b200 import SPExtractor, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    e = SPExtractor(800, 240, 320, WEIGHTS, max_batch=2, emit_heat=False, emit_cov=False)
    yield e
    e.close()


def unit(a):
    a = np.asarray(a, np.float32)
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def check_mutual(ex, q, t):
    got, dist = ex.match(q, t)
    ref, rdist, _ = O.match_mutual_nn(q, t)
    assert np.array_equal(got, ref)
    np.testing.assert_allclose(dist, rdist, atol=2e-6, rtol=2e-6)
    return got


def check_knn2(ex, q, t):
    idx, dist = ex.knn2(q, t)
    ridx, rdist = O.knn2(q, t)
    assert np.array_equal(idx, ridx)
    np.testing.assert_allclose(dist, rdist, atol=2e-6, rtol=2e-6)


def test_near_tie_clusters_inside_one_block(ex):
    """Clusters of 3 - 6 train rows that differ by 1e-4 (far below the fp16 dot-product resolution of 1e-3) sit inside one
    256-row block; the true nearest is planted LAST in each cluster, where a top-2 / top-3 nomination cannot see it."""
    rng = np.random.RandomState(3)
    nq, nt = 300, 1024
    q = unit(rng.randn(nq, 256))
    t = unit(rng.randn(nt, 256))
    for i in range(nq):
        size = 3 + i % 4
        base = (i * 3) % (nt - 8)
        for k in range(size):                              # k = size - 1 is q itself + the smallest perturbation
            noise = rng.randn(256).astype(np.float32) * np.float32(1e-4 * (size - k))
            t[base + k] = q[i] + noise
    t = unit(t)
    got = check_mutual(ex, q, t)
    assert (got >= 0).sum() > 50
    check_knn2(ex, q, t)


def test_duplicates_across_block_boundaries(ex):
    """Exact duplicates of a train row on both sides of the 256-row block boundaries: the first index wins (BFMatcher)."""
    rng = np.random.RandomState(5)
    q = unit(rng.randn(64, 256))
    t = unit(rng.randn(1100, 256))
    for i, pos in enumerate([255, 256, 511, 512, 767, 768, 1023, 1024]):
        t[pos] = q[i // 2]                                 # q[0] at 255 and 256, q[1] at 511 and 512, ...
    t[900] = q[0]
    got = check_mutual(ex, q, t)
    assert got[:4].tolist() == [255, 511, 767, 1023]
    check_knn2(ex, q, t)
    idx, dist = ex.knn2(q[:4], t)
    assert idx[:, 1].tolist() == [256, 512, 768, 1024] and np.all(dist[:4] < 1e-6)


@pytest.mark.parametrize("nq,nt", [(2001, 1777), (801, 801), (4096, 4096), (1, 300), (257, 255), (700, 3)])
def test_sizes_against_oracle(ex, nq, nt):
    rng = np.random.RandomState(nq + nt)
    q = unit(rng.randn(nq, 256))
    t = unit(np.concatenate([q[rng.permutation(nq)[:min(nq, nt)]] + 0.05 * rng.randn(min(nq, nt), 256).astype(np.float32),
                             rng.randn(max(nt - nq, 0), 256).astype(np.float32)])[:nt])
    check_mutual(ex, q, t)
    if nt >= 2:
        check_knn2(ex, q, t)


def test_non_unit_rows_take_the_exact_general_path(ex):
    rng = np.random.RandomState(9)
    q = rng.randn(500, 256).astype(np.float32) * 3.0       # not normalised: the tensor-core score bound does not apply
    t = np.concatenate([q[:300] + 0.1 * rng.randn(300, 256).astype(np.float32), rng.randn(200, 256).astype(np.float32)])
    check_mutual(ex, q, t)
    check_knn2(ex, q, t)
    check_mutual(ex, unit(q), t)                            # one unit set, one general set


def test_descriptor_sets_match_without_leaving_the_device(ex):
    """A key frame's rows uploaded once, the current frame's rows gathered device -> device from the extractor's slot:
    same results as the host-pointer entry on the same numbers."""
    frames = synth.make_stream(240, 320, 2, seed=19, n_shapes=260)
    outs = ex.extract_batch(list(frames))
    kf = ex.desc_set(1024).upload(outs[0]["desc"])
    cur = ex.desc_set(1024).from_frame(0, 1)
    assert kf.size() == outs[0]["n"] and cur.size() == outs[1]["n"]
    q2t, dist = ex.match_sets(cur, kf)
    ref, rdist, _ = O.match_mutual_nn(outs[1]["desc"], outs[0]["desc"])
    assert np.array_equal(q2t, ref) and (ref >= 0).mean() > 0.5
    np.testing.assert_allclose(dist, rdist, atol=2e-6, rtol=2e-6)
    idx, d2 = ex.knn2_sets(cur, kf)
    ridx, rd2 = O.knn2(outs[1]["desc"], outs[0]["desc"])
    assert np.array_equal(idx, ridx)
    rows = np.arange(0, outs[1]["n"], 3, dtype=np.int32)   # a subset (e.g. the key points that carry map points)
    sub = ex.desc_set(512).from_frame(0, 1, rows)
    q2t_s, _ = ex.match_sets(sub, kf)
    ref_s, _, _ = O.match_mutual_nn(outs[1]["desc"][rows], outs[0]["desc"])
    assert np.array_equal(q2t_s, ref_s)
    for s in (kf, cur, sub):
        s.close()


def test_stream_match_prev_is_exact_on_near_ties():
    """The in-pipeline matcher (SPFE_MATCH_PREV, top-2 nomination) uses the same re-scan rule: identical results to the
    oracle on a stream of repetitive texture (many near-identical descriptors)."""
    H, W = 240, 320
    tile = synth.make_frame(48, 64, seed=4, n_shapes=14)
    frame = np.tile(tile, (5, 5))                           # 25 copies of one patch: descriptors repeat up to border effects
    frames = np.stack([frame, np.roll(frame, 2, axis=1), np.roll(frame, 4, axis=1)])
    e = SPExtractor(800, H, W, WEIGHTS, max_batch=3, match_prev=True, emit_heat=False, emit_cov=False)
    outs = e.extract_batch(list(frames))
    for t in (1, 2):
        ref, rdist, _ = O.match_mutual_nn(outs[t]["desc"], outs[t - 1]["desc"])
        assert np.array_equal(outs[t]["match_prev"], ref)
        assert outs[t]["n"] > 100
    e.close()


def test_sets_left_alive_are_freed_with_the_context():
    e = SPExtractor(800, 64, 64, WEIGHTS, emit_heat=False, emit_cov=False)
    rng = np.random.RandomState(1)
    a = e.desc_set(300).upload(unit(rng.randn(300, 256)))
    b = e.desc_set(300).upload(unit(rng.randn(200, 256)))
    q2t, _ = e.match_sets(a, b)
    assert len(q2t) == 300
    b.close()
    b.close()                                               # double destroy: ignored
    e.close()                                               # `a` is still alive: freed by spfe_destroy
    a.close()                                               # dead handle, dead context: a no-op in the mirror


def test_matcher_entries_from_three_threads_while_extracting():
    """SearchByBruteForce runs on the tracking and the loop-closing thread, the FLANN replacement on the mapping thread,
    while the tracking thread keeps extracting (graph replays): concurrent calls give the single-threaded answers."""
    import threading
    H, W = 240, 320
    e = SPExtractor(800, H, W, WEIGHTS, emit_heat=False, emit_cov=True)
    frames = synth.make_stream(H, W, 4, seed=23, n_shapes=240)
    outs = [e.extract(f) for f in frames]                  # the third call on replays the captured graph
    cur = e.desc_set(1024).from_frame(0, 0)                 # slot 0 holds frames[3] now
    assert cur.size() == outs[3]["n"]
    q2t_set, _ = e.match_sets(cur, e.desc_set(1024).upload(outs[2]["desc"]))
    ref32, _, _ = O.match_mutual_nn(outs[3]["desc"], outs[2]["desc"])
    assert np.array_equal(q2t_set, ref32)
    want_m, _, _ = O.match_mutual_nn(outs[1]["desc"], outs[0]["desc"])
    want_k, _ = O.knn2(outs[2]["desc"], outs[1]["desc"])
    errors = []

    def worker(kind):
        try:
            for _ in range(25):
                if kind == "knn":
                    idx, _ = e.knn2(outs[2]["desc"], outs[1]["desc"])
                    assert np.array_equal(idx, want_k)
                else:
                    q2t, _ = e.match(outs[1]["desc"], outs[0]["desc"])
                    assert np.array_equal(q2t, want_m)
        except Exception as ex:                             # noqa: BLE001
            errors.append(repr(ex))
    threads = [threading.Thread(target=worker, args=(k,)) for k in ("mutual", "mutual", "knn")]
    for t in threads:
        t.start()
    for i in range(40):                                     # the tracking thread
        o = e.extract(frames[i % 4])
        assert np.array_equal(o["kp_xy"], outs[i % 4]["kp_xy"]) and np.array_equal(o["cov2"], outs[i % 4]["cov2"])
    for t in threads:
        t.join()
    assert not errors, errors
    e.close()
