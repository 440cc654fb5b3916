"""GPU tests of the throughput-mode output set (round 2): only n descriptor rows per frame cross PCIe, optionally as
fp16 (SPFE_DESC_F16), and heat_ / heat_inv_ stay on the device until spfe_fetch_heat asks for a frame (SPFE_LAZY_HEAT).
Every variant must deliver the bits of the eager fp32 path (fp16: the rounded bits)."""
import numpy as np
import pytest

from conftest import WEIGHTS
from sp_orb_slam_b200 import SPExtractor, SpfeError, capi, synth

pytestmark = pytest.mark.gpu

H, W, NF = 240, 320, 800


@pytest.fixture(scope="module")
def frames():
    return synth.make_stream(H, W, 5, seed=41, n_shapes=220)


@pytest.fixture(scope="module")
def eager(frames):
    ex = SPExtractor(NF, H, W, WEIGHTS, max_batch=5, emit_heat=True, emit_heat_inv=True, emit_cov=True, match_prev=True)
    outs = ex.extract_batch(list(frames))
    d2h = ex.last_d2h_bytes(0)
    ex.close()
    return outs, d2h


def test_only_valid_descriptor_rows_cross_pcie(frames, eager):
    outs, d2h = eager
    cap, cells, px = NF + 1, (H // 8) * (W // 8), H * W
    n = sum(o["n"] for o in outs)
    small = 5 * (4 + cap * 12 + cells * 10 + 2 * px * 4 + cap * 20 + cap * 8) + 4 + 4
    assert d2h == small + n * 1024                      # n rows of 1 KB, not cap rows
    assert all(0 < o["n"] < cap for o in outs)


def test_lazy_heat_and_fp16_descriptors_equal_eager_outputs(frames, eager):
    outs, d2h_eager = eager
    ex = SPExtractor(NF, H, W, WEIGHTS, max_batch=5, emit_heat=False, emit_heat_inv=False, emit_cov=True, match_prev=True,
                     lazy_heat=True, desc_f16=True)
    lean = ex.extract_batch(list(frames))
    assert ex.last_d2h_bytes(0) < d2h_eager / 3
    for t, (a, b) in enumerate(zip(outs, lean)):
        for k in ["kp_xy", "kp_score", "occ_grid", "dense_dust", "semi_dust", "cov2", "cov2_inv", "kp_response", "match_prev", "match_dist"]:
            assert np.array_equal(a[k], b[k]), k
        assert "heat" not in b
        assert np.array_equal(b["desc"], a["desc"].astype(np.float16).astype(np.float32))     # the fp32 rows, rounded once
        cos = np.einsum("ij,ij->i", a["desc"], b["desc"]) / (np.linalg.norm(a["desc"], axis=1) * np.linalg.norm(b["desc"], axis=1))
        assert (cos > 1 - 1e-6).all()                                                         # parity bar: 1 - 1e-3
        heat, heat_inv = ex.fetch_heat(0, t, heat=True, heat_inv=True)
        assert np.array_equal(heat, a["heat"]) and np.array_equal(heat_inv, a["heat_inv"])      # same kernel, same bits
    only_inv = ex.fetch_heat(0, 2, heat=False, heat_inv=True)
    assert only_inv[0] is None and np.array_equal(only_inv[1], outs[2]["heat_inv"])
    with pytest.raises(SpfeError) as e:
        ex.fetch_heat(0, 5)
    assert e.value.code == capi.ERR_STATE
    ex.close()


def test_lazy_heat_without_covariance(frames, eager):
    outs, _ = eager
    ex = SPExtractor(NF, H, W, WEIGHTS, max_batch=5, emit_heat=False, emit_heat_inv=False, emit_cov=False, lazy_heat=True)
    ex.extract_batch(list(frames))
    heat, _ = ex.fetch_heat(0, 4)
    assert np.array_equal(heat, outs[4]["heat"])
    ex.close()
    plain = SPExtractor(NF, H, W, WEIGHTS, emit_heat=False, emit_heat_inv=False, emit_cov=False)
    plain.extract_batch([frames[0]])
    with pytest.raises(SpfeError):                      # no heat maps on the device with these flags: fails loudly
        plain.fetch_heat(0, 0)
    plain.close()


def test_error_messages_are_per_thread():
    """spfe_last_error is the calling thread's last failure (the matcher entries run on three threads upstream)."""
    import threading
    ex = SPExtractor(NF, 64, 64, WEIGHTS, emit_heat=False, emit_cov=False)
    seen = {}

    def worker(name, slot):
        try:
            ex.sync(slot)
        except SpfeError as e:
            seen[name] = str(e)
    t = threading.Thread(target=worker, args=("bad", 7))
    t.start()
    t.join()
    assert "slot out of range" in seen["bad"]
    assert (ex._lib.spfe_last_error(ex._ctx) or b"") != seen["bad"].encode()      # this thread saw no failure of that call
    ex.close()


def test_extract_graph_replay_equals_call_by_call(frames, monkeypatch):
    """spfe_extract replays its single-frame launch plan as a CUDA graph from the third call on: every output must be
    bit-identical to the call-by-call enqueue (SPFE_GRAPH=0), also after a threshold change (re-capture) and when
    batched submits on the same slot are mixed in."""
    monkeypatch.setenv("SPFE_GRAPH", "0")
    plain = SPExtractor(NF, H, W, WEIGHTS, max_batch=2)
    ref = [plain.extract(f) for f in frames]
    plain.set_score_threshold(0.02)
    ref_hi = plain.extract(frames[0])
    plain.close()
    monkeypatch.setenv("SPFE_GRAPH", "1")
    ex = SPExtractor(NF, H, W, WEIGHTS, max_batch=2)
    keys = ["kp_xy", "kp_score", "desc", "occ_grid", "dense_dust", "semi_dust", "heat", "heat_inv", "cov2", "cov2_inv", "kp_response"]
    l0 = ex.launch_count()
    got = [ex.extract(f) for f in frames]                 # call 1 plain, call 2 captures + replays, 3.. replay
    per_call = (ex.launch_count() - l0) / len(frames)
    assert 20 <= per_call <= 40
    for a, b in zip(ref, got):
        assert a["n"] == b["n"]
        for k in keys:
            assert np.array_equal(a[k], b[k]), k
    batch = ex.extract_batch([frames[1], frames[2]])       # a batched submit on the same slot in between
    for a, b in zip(ref[1:3], batch):
        for k in keys:
            assert np.array_equal(a[k], b[k]), k
    again = ex.extract(frames[3])
    for k in keys:
        assert np.array_equal(again[k], ref[3][k]), k
    ex.set_score_threshold(0.02)                           # baked into the graph: must be re-captured
    hi = ex.extract(frames[0])
    assert hi["n"] == ref_hi["n"] < ref[0]["n"]
    for k in keys:
        assert np.array_equal(hi[k], ref_hi[k]), k
    ex.close()
