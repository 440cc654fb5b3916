"""Host-side Python mirror of the reference's extractor / matcher interface.

``SPExtractor`` follows ``orbslam::SPExtractor`` (reference
``orb_slam2/include/orb_slam/cv/sp_extractor.h:49-88``): construct with
``nfeatures``, call it with a CV_8UC1 image, get ``(keypoints, descriptors)``
and read the side outputs ``semi_dust_ / dense_dust_ / heat_ / heat_inv_ /
occ_grid_ / getCov() / getCov2Inv()`` afterwards.  ``SPMatcher`` follows
``orbslam::SPMatcher`` (``sp_matcher.h:13-103``) for the brute-force path.
Geometry and model path are constructor arguments here instead of the
reference's ``camera::`` / ``common::`` globals.

Everything runs through the C ABI in ``include/spfe.h``; nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class SpfeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"spfe error {code}: {msg}")
        self.code = code


def _as_np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if not ptr or n == 0:
        return np.zeros(shape, dtype)
    ct = {np.float32: C.c_float, np.int16: C.c_int16}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape).copy()


class SPExtractor:
    """Drop-in mirror of ``orbslam::SPExtractor`` (blocking, batch 1) plus the batched entry points."""

    def __init__(self, nfeatures: int, height: int, width: int, model_path: str, *, device_id: int = 0,
                 max_batch: int = 1, num_slots: int = 1, emit_heat: bool = True, emit_cov: bool = True,
                 match_prev: bool = False, emit_heat_inv: bool | None = None, lazy_heat: bool = False,
                 desc_f16: bool = False, exact: bool = False):
        self._lib = capi.load()
        self._ctx = C.c_void_p()
        cfg = capi.Config()
        self._lib.spfe_default_config(C.byref(cfg), height, width, nfeatures)
        cfg.device_id, cfg.max_batch, cfg.num_slots = device_id, max_batch, num_slots
        emit_heat_inv = emit_heat if emit_heat_inv is None else emit_heat_inv
        cfg.flags = ((capi.EMIT_HEAT if emit_heat else 0) | (capi.EMIT_COV if emit_cov else 0)
                     | (capi.MATCH_PREV if match_prev else 0) | (capi.EMIT_HEAT_INV if emit_heat_inv else 0)
                     | (capi.LAZY_HEAT if lazy_heat else 0) | (capi.DESC_F16 if desc_f16 else 0)
                     | (capi.EXACT if exact else 0))
        self.exact = exact
        self.lazy_heat, self.desc_f16 = lazy_heat, desc_f16
        self.emit_heat_inv = emit_heat_inv
        self.match_prev = match_prev
        self._path = str(model_path).encode()
        cfg.weights_path = self._path
        self.cfg = cfg
        rc = self._lib.spfe_create(C.byref(cfg), C.byref(self._ctx))
        if rc != capi.OK:
            msg = (self._lib.spfe_last_error(None) or b"").decode()
            self._ctx = C.c_void_p()
            raise SpfeError(rc, msg)
        self.height, self.width, self.nfeatures = height, width, nfeatures
        self.hc, self.wc = height // 8, width // 8
        self.max_batch, self.num_slots = max_batch, num_slots
        self.emit_heat, self.emit_cov = emit_heat, emit_cov
        self.cap = min(nfeatures + 1, self.hc * self.wc)
        # side outputs of the last operator() call (sp_extractor.h:69-77)
        self.semi_dust_ = self.dense_dust_ = self.heat_ = self.heat_inv_ = self.occ_grid_ = self.mask_ = None
        self._cov2 = self._cov2_inv = None
        self._keep = None

    # -- BaseExtractor scale getters (base_extractor.h:58-72): one level, factor 1.0
    def GetLevels(self): return 1
    def GetScaleFactor(self): return 1.0
    def GetScaleFactors(self): return [1.0]
    def GetInverseScaleFactors(self): return [1.0]
    def GetScaleSigmaSquares(self): return [1.0]
    def GetInverseScaleSigmaSquares(self): return [1.0]

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.spfe_destroy(self._ctx)
            self._ctx = C.c_void_p()
            for p in getattr(self, "_pinned", []):
                self._lib.spfe_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise SpfeError(rc, (self._lib.spfe_last_error(self._ctx) or b"").decode())
        return rc

    def _unpack(self, o: capi.FrameOut) -> dict:
        n, hc, wc, H, W = o.n, self.hc, self.wc, self.height, self.width
        if self.desc_f16:       # binary16 on the wire, widened here exactly like the C++ shim does
            desc = (np.ctypeslib.as_array(o.desc_f16, shape=(n * 256,)).view(np.float16).reshape(n, 256).astype(np.float32)
                    if n else np.zeros((0, 256), np.float32))
        else:
            desc = _as_np(o.desc, (n, 256), np.float32)
        d = dict(n=n, kp_xy=_as_np(o.kp_xy, (n, 2), np.float32), kp_score=_as_np(o.kp_score, (n,), np.float32),
                 desc=desc, occ_grid=_as_np(o.occ_grid, (hc, wc), np.int16),
                 dense_dust=_as_np(o.dense_dust, (hc, wc), np.float32), semi_dust=_as_np(o.semi_dust, (hc, wc), np.float32))
        if self.emit_heat:
            d["heat"] = _as_np(o.heat, (H, W), np.float32)
        if self.emit_heat_inv:
            d["heat_inv"] = _as_np(o.heat_inv, (H, W), np.float32)
        if self.match_prev:
            d["n_prev"] = o.n_prev
            d["match_prev"] = (np.ctypeslib.as_array(o.match_prev, shape=(max(n, 1),))[:n].copy() if n else np.zeros(0, np.int32))
            d["match_dist"] = _as_np(o.match_dist, (n,), np.float32)
        if self.emit_cov:
            d["kp_response"] = _as_np(o.kp_response, (n,), np.float32)
            d["cov2"] = _as_np(o.cov2, (n, 2), np.float32)
            d["cov2_inv"] = _as_np(o.cov2_inv, (n, 2), np.float32)
        return d

    # -- the reference operator(): image -> (keypoints, descriptors)
    def __call__(self, image: np.ndarray, mask=None):
        """``operator()(image, mask, keypoints, descriptors)`` -- mask is ignored, like the reference.

        Returns (keypoints [n,3] = x, y, response ; descriptors [n,256] f32)."""
        out = self.extract(image)
        self.semi_dust_, self.dense_dust_ = out["semi_dust"], out["dense_dust"]
        self.heat_, self.heat_inv_ = out.get("heat"), out.get("heat_inv")
        self.occ_grid_ = out["occ_grid"]
        self._cov2, self._cov2_inv = out.get("cov2"), out.get("cov2_inv")
        resp = out.get("kp_response", out["kp_score"])
        return np.concatenate([out["kp_xy"], resp[:, None]], 1), out["desc"]

    def getCov(self): return self._cov2
    def getCov2Inv(self): return self._cov2_inv
    def getHeatMap(self): return self.heat_
    def getMask(self): return self.mask_

    def extract(self, image: np.ndarray) -> dict:
        if image is None or image.size == 0:
            raise RuntimeError("input image is empty")                 # sp_extractor.cpp:364-365
        assert image.dtype == np.uint8 and image.ndim == 2             # sp_extractor.cpp:368
        if image.shape != (self.height, self.width):
            raise SpfeError(capi.ERR_INVALID, f"image is {image.shape}, extractor was built for {(self.height, self.width)}")
        img = np.ascontiguousarray(image) if image.strides[1] != 1 else image
        o = capi.FrameOut()
        self._check(self._lib.spfe_extract(self._ctx, img.ctypes.data_as(C.c_void_p), img.strides[0], C.byref(o)))
        return self._unpack(o)

    # -- batched / pipelined entry points (no reference equivalent: sp_extractor.cpp:70 "TODO: batch-size")
    def submit(self, slot: int, frames) -> None:
        frames = [np.ascontiguousarray(f) for f in frames]
        for f in frames:
            assert f.dtype == np.uint8 and f.shape == (self.height, self.width)
        self._keep = frames
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        self._check(self._lib.spfe_submit(self._ctx, slot, ptrs, len(frames), self.width))

    def submit_pinned(self, slot: int, frames: np.ndarray) -> None:
        """``frames``: C-contiguous [batch][H][W] u8, ideally page-locked (``pinned_frames``); DMA'd as is, so it must stay
        untouched until ``wait(slot)`` returns."""
        assert frames.dtype == np.uint8 and frames.flags.c_contiguous and frames.shape[1:] == (self.height, self.width)
        self._keep = frames
        self._check(self._lib.spfe_submit_pinned(self._ctx, slot, C.c_void_p(frames.ctypes.data), frames.shape[0]))

    def pinned_frames(self, n: int) -> np.ndarray:
        """Page-locked [n][H][W] u8 buffer (spfe_host_alloc); freed with the extractor."""
        nbytes = n * self.height * self.width
        p = self._lib.spfe_host_alloc(nbytes)
        if not p:
            raise SpfeError(capi.ERR_CUDA, "spfe_host_alloc failed")
        self._pinned = getattr(self, "_pinned", []) + [p]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).reshape(n, self.height, self.width)

    def wait(self, slot: int, n_frames: int, unpack: bool = True):
        outs = (capi.FrameOut * n_frames)()
        self._check(self._lib.spfe_wait(self._ctx, slot, outs))
        return [self._unpack(o) for o in outs] if unpack else outs

    def last_d2h_bytes(self, slot: int = 0) -> int:
        """Bytes the last ``wait(slot)`` moved device -> host."""
        return int(self._check(self._lib.spfe_last_d2h_bytes(self._ctx, slot)))

    def fetch_heat(self, slot: int, frame: int, heat: bool = True, heat_inv: bool = False):
        """``heat_`` / ``heat_inv_`` of one frame of the slot's last batch, copied on demand (spfe_fetch_heat)."""
        h = np.empty((self.height, self.width), np.float32) if heat else None
        hi = np.empty((self.height, self.width), np.float32) if heat_inv else None
        self._check(self._lib.spfe_fetch_heat(self._ctx, slot, frame, None if h is None else h.ctypes.data,
                                              None if hi is None else hi.ctypes.data))
        return h, hi

    def extract_batch(self, frames, slot: int = 0):
        self.submit(slot, frames)
        return self.wait(slot, len(frames))

    def submit_device(self, slot: int, d_ptr: int, batch: int) -> None:
        """Frames already in HBM ([batch][H][W] u8 at device address ``d_ptr``); results stay on the device."""
        self._check(self._lib.spfe_submit_device(self._ctx, slot, C.c_void_p(d_ptr), batch))

    def sync(self, slot: int) -> None:
        self._check(self._lib.spfe_slot_sync(self._ctx, slot))

    def set_score_threshold(self, thresh: float) -> None:
        """Detection threshold for the batches submitted from now on (the common cut of ``sharding.global_keypoint_budget``)."""
        self._check(self._lib.spfe_set_score_threshold(self._ctx, C.c_float(thresh)))

    def reset_stream(self, slot: int) -> None:
        self._check(self._lib.spfe_reset_stream(self._ctx, slot))

    def timer_start(self, slot: int) -> None:
        self._check(self._lib.spfe_timer_start(self._ctx, slot))

    def timer_stop(self, slot: int) -> float:
        ms = C.c_float()
        self._check(self._lib.spfe_timer_stop(self._ctx, slot, C.byref(ms)))
        return float(ms.value)

    def match(self, q: np.ndarray, t: np.ndarray):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 256)
        t = np.ascontiguousarray(t, np.float32).reshape(-1, 256)
        q2t = np.empty(max(len(q), 1), np.int32)
        dist = np.empty(max(len(q), 1), np.float32)
        self._check(self._lib.spfe_match_mutual_nn(self._ctx, q.ctypes.data_as(C.c_void_p), len(q),
                                                   t.ctypes.data_as(C.c_void_p), len(t),
                                                   q2t.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return q2t[:len(q)], dist[:len(q)]

    def knn2(self, q: np.ndarray, t: np.ndarray):
        """Exact 2-NN (spfe_match_knn2) -> (idx int32[nq,2], dist f32[nq,2]); -1 / 0 where no such row exists."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 256)
        t = np.ascontiguousarray(t, np.float32).reshape(-1, 256)
        idx = np.empty((max(len(q), 1), 2), np.int32)
        dist = np.empty((max(len(q), 1), 2), np.float32)
        self._check(self._lib.spfe_match_knn2(self._ctx, q.ctypes.data_as(C.c_void_p), len(q), t.ctypes.data_as(C.c_void_p), len(t),
                                              idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return idx[:len(q)], dist[:len(q)]

    # -- device-resident descriptor sets (spfe_desc_set_*)
    def desc_set(self, capacity: int) -> "DescSet":
        return DescSet(self, capacity)

    def match_sets(self, q: "DescSet", t: "DescSet"):
        n = q.size()
        q2t, dist = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float32)
        self._check(self._lib.spfe_match_mutual_nn_sets(self._ctx, q._h, t._h, q2t.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return q2t[:n], dist[:n]

    def knn2_sets(self, q: "DescSet", t: "DescSet"):
        n = q.size()
        idx, dist = np.empty((max(n, 1), 2), np.int32), np.empty((max(n, 1), 2), np.float32)
        self._check(self._lib.spfe_match_knn2_sets(self._ctx, q._h, t._h, idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p)))
        return idx[:n], dist[:n]

    def search_guided(self, qdesc, qxy, qradius, occ, kp_un, kdesc, *, mode: int, best_init: float, th_le: float, th_lt: float,
                      c2_adaptive: float = 0.0, qvalid=None, qblocks=None, kp_taken=None, min_x: float = 0.0, min_y: float = 0.0):
        """spfe_search_guided on plain arrays -> (q2kp int32[m], qdist f32[m], kp_taken_after u8[n]).  ``qdesc`` / ``kdesc``
        may be ``DescSet`` objects (spfe_search_guided_sets: the descriptors are on the device already)."""
        qset = qdesc if isinstance(qdesc, DescSet) else None
        kset = kdesc if isinstance(kdesc, DescSet) else None
        qdesc = None if qset else np.ascontiguousarray(qdesc, np.float32).reshape(-1, 256)
        m = qset.size() if qset else len(qdesc)
        kdesc = None if kset else np.ascontiguousarray(kdesc, np.float32).reshape(-1, 256)
        n = kset.size() if kset else len(kdesc)
        qxy = np.ascontiguousarray(qxy, np.float32).reshape(m, 2)
        qr = np.ascontiguousarray(np.broadcast_to(np.asarray(qradius, np.float32), (m,)))
        occ = np.ascontiguousarray(occ, np.int16)
        kp_un = None if kp_un is None else np.ascontiguousarray(kp_un, np.float32).reshape(n, 2)
        opt = [None if a is None else np.ascontiguousarray(a, np.uint8) for a in (qvalid, qblocks, kp_taken)]
        g = capi.GuidedSearch()
        g.struct_size, g.mode, g.m, g.n = C.sizeof(capi.GuidedSearch), mode, m, n
        ptr = lambda a: None if a is None or a.size == 0 else a.ctypes.data
        g.qdesc, g.qxy, g.qradius = ptr(qdesc), ptr(qxy), ptr(qr)
        g.qvalid, g.qblocks, g.kp_taken = ptr(opt[0]), ptr(opt[1]), ptr(opt[2])
        g.kdesc, g.kp_un, g.occ_grid = ptr(kdesc), ptr(kp_un), ptr(occ)
        g.grid_rows, g.grid_cols = occ.shape
        g.min_x, g.min_y, g.best_init, g.th_le, g.th_lt, g.c2_adaptive = min_x, min_y, best_init, th_le, th_lt, c2_adaptive
        q2kp = np.empty(max(m, 1), np.int32)
        qdist = np.empty(max(m, 1), np.float32)
        taken = np.zeros(max(n, 1), np.uint8)
        if qset or kset:
            self._check(self._lib.spfe_search_guided_sets(self._ctx, C.byref(g), qset._h if qset else None, kset._h if kset else None,
                                                          q2kp.ctypes.data_as(C.c_void_p), qdist.ctypes.data_as(C.c_void_p),
                                                          taken.ctypes.data_as(C.c_void_p)))
        else:
            self._check(self._lib.spfe_search_guided(self._ctx, C.byref(g), q2kp.ctypes.data_as(C.c_void_p),
                                                     qdist.ctypes.data_as(C.c_void_p), taken.ctypes.data_as(C.c_void_p)))
        return q2kp[:m], qdist[:m], taken[:n]

    def _dust_struct(self, Xw, dust, fx, fy, cx, cy, huber, chi2_inlier, iterations, slot, frame):
        Xw = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
        d = capi.DustPose()
        d.struct_size, d.n = C.sizeof(capi.DustPose), len(Xw)
        d.Xw = Xw.ctypes.data if len(Xw) else None
        if dust is not None:
            dust = np.ascontiguousarray(dust, np.float32)
            d.dust, d.rows, d.cols = dust.ctypes.data, dust.shape[0], dust.shape[1]
        else:
            d.dust, d.rows, d.cols, d.slot, d.frame = None, 0, 0, slot, frame
        d.fx, d.fy, d.cx, d.cy, d.huber_delta, d.chi2_inlier, d.iterations = fx, fy, cx, cy, huber, chi2_inlier, iterations
        return d, Xw, dust

    def dust_pose_optimize(self, pose7, Xw, fx, fy, cx, cy, *, dust=None, slot: int = 0, frame: int = 0, huber: float = 0.9,
                           chi2_inlier: float = 0.9, iterations: int = 40):
        """spfe_dust_pose_optimize on plain arrays.  ``dust=None`` uses the dense_dust map of ``frame`` of ``slot``'s last
        batch where it lies on the device.  -> dict(pose, visible, uv, n_inlier, n_iter, stats)."""
        d, Xw, dust = self._dust_struct(Xw, dust, fx, fy, cx, cy, huber, chi2_inlier, iterations, slot, frame)
        n = len(Xw)
        pose = np.ascontiguousarray(pose7, np.float64).copy()
        vis, uv, stats = np.zeros(max(n, 1), np.uint8), np.zeros((max(n, 1), 2), np.float32), np.zeros(3)
        ninl, nit = C.c_int32(0), C.c_int32(0)
        vp = C.c_void_p
        self._check(self._lib.spfe_dust_pose_optimize(self._ctx, C.byref(d), vp(pose.ctypes.data), vp(vis.ctypes.data),
                                                      vp(uv.ctypes.data), C.byref(ninl), C.byref(nit), vp(stats.ctypes.data)))
        return dict(pose=pose, visible=vis[:n], uv=uv[:n], n_inlier=ninl.value, n_iter=nit.value, stats=stats)

    def dust_pose_optimize_batch(self, problems, *, huber: float = 0.9, chi2_inlier: float = 0.9, iterations: int = 40):
        """spfe_dust_pose_optimize_batch: ``problems`` = list of dicts(pose, Xw, cam=(fx, fy, cx, cy), and either ``dust``
        (host map) or ``slot`` / ``frame``); all solves run in one launch, one CTA each.  -> list of result dicts."""
        cnt = len(problems)
        arr = (capi.DustPose * max(cnt, 1))()
        keep, poses = [], np.zeros((max(cnt, 1), 7))
        vis_p, uv_p = (C.c_void_p * max(cnt, 1))(), (C.c_void_p * max(cnt, 1))()
        vis, uv = [], []
        for i, pr in enumerate(problems):
            d, Xw, dust = self._dust_struct(pr["Xw"], pr.get("dust"), *pr["cam"], huber, chi2_inlier, iterations,
                                            pr.get("slot", 0), pr.get("frame", 0))
            arr[i] = d
            keep.append((Xw, dust))
            poses[i] = pr["pose"]
            vis.append(np.zeros(max(len(Xw), 1), np.uint8))
            uv.append(np.zeros((max(len(Xw), 1), 2), np.float32))
            vis_p[i], uv_p[i] = vis[i].ctypes.data, uv[i].ctypes.data
        ninl, nit = np.zeros(max(cnt, 1), np.int32), np.zeros(max(cnt, 1), np.int32)
        vp = C.c_void_p
        self._check(self._lib.spfe_dust_pose_optimize_batch(self._ctx, arr, cnt, vp(poses.ctypes.data), C.cast(vis_p, vp), C.cast(uv_p, vp),
                                                            vp(ninl.ctypes.data), vp(nit.ctypes.data)))
        return [dict(pose=poses[i].copy(), visible=vis[i][:len(keep[i][0])], uv=uv[i][:len(keep[i][0])], n_inlier=int(ninl[i]),
                     n_iter=int(nit[i])) for i in range(cnt)]

    def dust_linearize(self, pose7, Xw, fx, fy, cx, cy, *, dust=None, slot: int = 0, frame: int = 0, huber: float = 0.9, level=None):
        """spfe_dust_linearize on plain arrays -> dict(level, err, uv, J, H, b, chi2)."""
        d, Xw, dust = self._dust_struct(Xw, dust, fx, fy, cx, cy, huber, 0.9, 0, slot, frame)
        n = len(Xw)
        pose = np.ascontiguousarray(pose7, np.float64)
        level = np.zeros(max(n, 1), np.uint8) if level is None else np.ascontiguousarray(level, np.uint8).copy()
        err, uv, J, Hb = np.zeros(max(n, 1)), np.zeros((max(n, 1), 2), np.float32), np.zeros((max(n, 1), 6)), np.zeros(43)
        vp = C.c_void_p
        self._check(self._lib.spfe_dust_linearize(self._ctx, C.byref(d), vp(pose.ctypes.data), vp(level.ctypes.data), vp(err.ctypes.data),
                                                  vp(uv.ctypes.data), vp(J.ctypes.data), vp(Hb.ctypes.data)))
        return dict(level=level[:n], err=err[:n], uv=uv[:n], J=J[:n], H=Hb[:36].reshape(6, 6).copy(), b=Hb[36:42].copy(), chi2=float(Hb[42]))

    # -- introspection
    _DEBUG = {"conv1a": (lambda s: (s.height, s.width, 64), np.float16), "conv1b": (lambda s: (s.height // 2, s.width // 2, 64), np.float16),
              "conv2a": (lambda s: (s.height // 2, s.width // 2, 64), np.float16), "conv2b": (lambda s: (s.height // 4, s.width // 4, 64), np.float16),
              "conv3a": (lambda s: (s.height // 4, s.width // 4, 128), np.float16), "conv3b": (lambda s: (s.hc, s.wc, 128), np.float16),
              "conv4a": (lambda s: (s.hc, s.wc, 128), np.float16), "conv4b": (lambda s: (s.hc, s.wc, 128), np.float16),
              "heads": (lambda s: (s.hc, s.wc, 512), np.float16), "coarse": (lambda s: (s.hc, s.wc, 256), np.float16),
              "score": (lambda s: (s.hc, s.wc), np.float32), "argmax": (lambda s: (s.hc, s.wc), np.uint8),
              "semi_dust": (lambda s: (s.hc, s.wc), np.float32), "dense_dust": (lambda s: (s.hc, s.wc), np.float32),
              "heat_log": (lambda s: (s.height, s.width), np.float32), "heat": (lambda s: (s.height, s.width), np.float32),
              "heat_inv": (lambda s: (s.height, s.width), np.float32), "heat_minmax": (lambda s: (2,), np.float32),
              "count": (lambda s: (), np.int32), "kp_xy": (lambda s: (s.cap, 2), np.float32), "kp_score": (lambda s: (s.cap,), np.float32),
              "desc": (lambda s: (s.cap, 256), np.float32), "occ_grid": (lambda s: (s.hc, s.wc), np.int16),
              "match_prev": (lambda s: (s.cap,), np.int32), "match_dist": (lambda s: (s.cap,), np.float32),
              "cov_qlen": (lambda s: (s.cap,), np.int32), "cov_done": (lambda s: (s.cap,), np.int32),
              "cov_counters": (lambda s: (4,), np.int32), "cov_replayed": (lambda s: (2,), np.int32)}

    _XP_LAYERS = ("conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b", "heads")

    def debug_read(self, slot: int, name: str, batch: int) -> np.ndarray:
        shape_fn, dt = self._DEBUG[name]
        shape = tuple(shape_fn(self))
        xp = self.exact and name in self._XP_LAYERS   # exact mode: [hi 64 | lo 64] per 64-channel block -> hi + lo as fp32
        if xp:
            shape = shape[:-1] + (2 * shape[-1],)
        arr = np.empty((batch,) + shape, dt)
        self._check(self._lib.spfe_debug_read(self._ctx, slot, name.encode(), arr.ctypes.data_as(C.c_void_p), arr.nbytes))
        if xp:
            pairs = arr.reshape(arr.shape[:-1] + (shape[-1] // 128, 2, 64)).astype(np.float32)
            return (pairs[..., 0, :] + pairs[..., 1, :]).reshape(arr.shape[:-1] + (shape[-1] // 2,))
        return arr

    def launch_count(self) -> int:
        return int(self._lib.spfe_launch_count(self._ctx))

    def dom_timing(self, enable: bool) -> None:
        self._check(self._lib.spfe_dom_timing(self._ctx, 1 if enable else 0))

    def dom_time(self):
        """-> (average launch duration in ms of the dominant kernel over the last <= 64 batches, how many)."""
        ms, n = C.c_float(), C.c_int32()
        self._check(self._lib.spfe_dom_time(self._ctx, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def profile_device(self, slot: int, d_ptr: int, batch: int):
        st = (capi.StageTime * 32)()
        n = self._check(self._lib.spfe_profile_device(self._ctx, slot, C.c_void_p(d_ptr), batch, st, 32))
        return [dict(name=st[i].name.decode(), ms=st[i].ms, flop=st[i].flop, bytes=st[i].bytes) for i in range(n)]


class DescSet:
    """Device-resident descriptor rows (``spfe_desc_set``): upload once, match many times."""

    def __init__(self, ex: SPExtractor, capacity: int):
        self._ex, self._h = ex, C.c_void_p()
        ex._check(ex._lib.spfe_desc_set_create(ex._ctx, capacity, C.byref(self._h)))

    def upload(self, rows: np.ndarray) -> "DescSet":
        rows = np.ascontiguousarray(rows, np.float32).reshape(-1, 256)
        self._ex._check(self._ex._lib.spfe_desc_set_upload(self._ex._ctx, self._h, rows.ctypes.data_as(C.c_void_p), len(rows)))
        return self

    def from_frame(self, slot: int, frame: int, rows=None) -> "DescSet":
        if rows is None:
            self._ex._check(self._ex._lib.spfe_desc_set_from_frame(self._ex._ctx, self._h, slot, frame, None, 0))
        else:
            rows = np.ascontiguousarray(rows, np.int32)
            self._ex._check(self._ex._lib.spfe_desc_set_from_frame(self._ex._ctx, self._h, slot, frame, rows.ctypes.data_as(C.c_void_p), len(rows)))
        return self

    def size(self) -> int:
        return int(self._ex._lib.spfe_desc_set_size(self._h))

    def close(self):
        if self._h and self._ex._ctx.value:
            self._ex._lib.spfe_desc_set_destroy(self._ex._ctx, self._h)
        self._h = C.c_void_p()


class SPMatcher:
    """Mirror of ``orbslam::SPMatcher`` for the brute-force path (sp_matcher.h:16-19,48-49,86-93)."""

    TH_HIGH, TH_LOW, HISTO_LENGTH = 0.7, 0.3, 30           # sp_matcher.cpp:18-20

    def __init__(self, extractor: SPExtractor, nnratio: float = 0.6):
        self._ex = extractor
        self.mfNNratio = nnratio

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> float:
        a = np.ascontiguousarray(a, np.float32).ravel()
        b = np.ascontiguousarray(b, np.float32).ravel()
        assert a.size == 256 and b.size == 256
        return float(capi.load().spfe_l2(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)))

    def KnnMatchRatio(self, query_desc: np.ndarray, train_desc: np.ndarray, ratio: float = 0.7):
        """``flann->knnMatch(query, matches, 2)`` + the ratio test of SearchForTriByFlann / SearchByFlann
        (sp_matcher.cpp:197-206, :266-270) with an exact 2-NN.  -> (train index or -1 per query row, idx[nq,2], dist[nq,2])."""
        idx, dist = self._ex.knn2(query_desc, train_desc)
        good = (idx[:, 0] >= 0) & (idx[:, 1] >= 0) & (dist[:, 0] < np.float32(ratio) * dist[:, 1])
        return np.where(good, idx[:, 0], -1), idx, dist

    @staticmethod
    def RadiusByViewingCos(view_cos):
        """sp_matcher.cpp:434-439."""
        return np.where(np.asarray(view_cos, np.float32) > np.float32(0.998), np.float32(2.5), np.float32(4.0))

    def SearchByProjectionMapPoints(self, frame: dict, mp_desc, proj_xy, view_cos, *, th: float = 1.0, th_dist: float = 0.7,
                                    in_view=None, observed=None, c2_adaptive: float = 0.0):
        """``SearchByProjection(Frame &F, const vector<MapPoint*>&, th, th_dist)`` (sp_matcher.cpp:344-432) on arrays.
        ``frame``: dict with ``desc`` [n,256], ``kp_un`` [n,2] (mvKeysUn), ``occ_grid``, optional ``taken`` [n] (keypoints
        that already carry an observed map point).  Per map point: ``mp_desc`` (getDescTrack), ``proj_xy`` (mTrackProjX/Y),
        ``view_cos`` (mTrackViewCos), ``in_view`` (mbTrackInView && !isBad()), ``observed`` (Observations() > 0);
        ``c2_adaptive`` = tracking::dust::c2_thresh when tracking::map::match_adaptive.  Returns (mp2kp, nmatches):
        F.mvpMapPoints[mp2kp[i]] = vpMapPoints[i] for mp2kp[i] >= 0, applied in order."""
        r = self.RadiusByViewingCos(view_cos)
        if th != 1.0:
            r = r * np.float32(th)
        q2kp, _, _ = self._ex.search_guided(mp_desc, proj_xy, r, frame["occ_grid"], frame["kp_un"], frame["desc"],
                                            mode=capi.GUIDED_AREA, best_init=256.0, th_le=th_dist, th_lt=0.7,
                                            c2_adaptive=c2_adaptive, qvalid=in_view, qblocks=observed, kp_taken=frame.get("taken"))
        return q2kp, int((q2kp >= 0).sum())

    def SearchByProjectionLastFrame(self, frame: dict, last_desc, proj_xy, *, th: float, valid=None, observed=None):
        """``SearchByProjection(Frame &Cur, const Frame &Last, th, bMono)`` (sp_matcher.cpp:1439-1543), monocular: the
        caller projects the last frame's map points into the current frame (:1464-1482, ``valid`` = has a map point,
        not an outlier, in front of the camera, inside the image); radius = th * mvScaleFactors[0] = th."""
        q2kp, _, _ = self._ex.search_guided(last_desc, proj_xy, np.float32(th), frame["occ_grid"], frame["kp_un"], frame["desc"],
                                            mode=capi.GUIDED_AREA, best_init=float(np.finfo(np.float32).max), th_le=self.TH_HIGH,
                                            th_lt=-np.inf, qvalid=valid, qblocks=observed, kp_taken=frame.get("taken"))
        return q2kp, int((q2kp >= 0).sum())

    def DustAssociate(self, frame: dict, mp_desc, dust_proj_uv, *, in_view=None):
        """Patch-wise association of dust tracking (tracker_dust.cpp:112-172): ``dust_proj_uv`` in occ_grid cell units."""
        q2kp, _, _ = self._ex.search_guided(mp_desc, dust_proj_uv, np.float32(0), frame["occ_grid"], None, frame["desc"],
                                            mode=capi.GUIDED_DUST_CELLS, best_init=0.75, th_le=-np.inf, th_lt=0.75, qvalid=in_view)
        return q2kp, int((q2kp >= 0).sum())

    def SearchByBruteForce(self, desc1: np.ndarray, valid1: np.ndarray, desc2: np.ndarray, valid2: np.ndarray | None = None):
        """Both reference overloads on plain arrays.

        (KeyFrame*, Frame&) [sp_matcher.cpp:1642-1674]: ``desc1/valid1`` = key-frame descriptors and "has a good map
        point" flags (train side), ``desc2`` = all frame descriptors (query side), ``valid2=None``.  Returns
        ``matches12`` of length len(desc2): index into desc1 of the matched map point, -1 = none.
        (KeyFrame*, KeyFrame*) [sp_matcher_loop.cpp:334-376]: pass ``valid2`` too; returns ``(matches12, n)`` with
        matches12 of length len(desc1): index into desc2, -1 = none.
        """
        idx_t = np.flatnonzero(valid1)
        if valid2 is None:
            q2t, _ = self._ex.match(desc2, desc1[idx_t])
            out = np.full(len(desc2), -1, np.int64)
            hit = q2t >= 0
            out[hit] = idx_t[q2t[hit]]
            return out
        idx_q = np.flatnonzero(valid2)
        q2t, _ = self._ex.match(desc2[idx_q], desc1[idx_t])
        out = np.full(len(desc1), -1, np.int64)
        hit = np.flatnonzero(q2t >= 0)
        out[idx_t[q2t[hit]]] = idx_q[hit]
        return out, int(len(hit))


class Optimizer:
    """Mirror of the one ``orbslam::Optimizer`` entry that sits on the dust-tracking path
    (orb_slam2/src/mapping/optimizer_dust.cpp:170-293), on plain arrays."""

    HUBER_DELTA, CHI2_INLIER, ITERATIONS = 0.9, 0.9, 40    # optimizer_dust.cpp:219, :253, :246

    def __init__(self, extractor: SPExtractor):
        self._ex = extractor

    def PoseOptimizationDust(self, Tcw_pose7, map_points_xyz, fx: float, fy: float, cx: float, cy: float, *, dust=None,
                             slot: int = 0, frame: int = 0):
        """``PoseOptimizationDust(Frame*, mps, is_visible)``: ``fx .. cy`` are the FRAME's intrinsics in pixels (the /8 and
        -3.5 of optimizer_dust.cpp:222-225 are applied here); ``Tcw_pose7`` = (qx, qy, qz, qw, tx, ty, tz).
        Returns (n_inlier, pose7, is_visible[n], dust_proj_uv[n, 2])."""
        r = self._ex.dust_pose_optimize(Tcw_pose7, map_points_xyz, fx / 8.0, fy / 8.0, (cx - 3.5) / 8.0, (cy - 3.5) / 8.0,
                                        dust=dust, slot=slot, frame=frame, huber=self.HUBER_DELTA,
                                        chi2_inlier=self.CHI2_INLIER, iterations=self.ITERATIONS)
        return r["n_inlier"], r["pose"], r["visible"].astype(bool), r["uv"]
