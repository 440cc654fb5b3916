"""ctypes wrapper around oracle/_ref/libsppost_ref.so = the REFERENCE's own ``nms`` and ``computeCovariance``
(sp_extractor.cpp:161-340), compiled verbatim from /root/reference against ``oracle/ref_cv_stub.h`` by
``oracle/ref_build.sh``.  Test infrastructure only: it pins the C restatement ``oracle/sp_post.c``
(tests/test_oracle.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libsppost_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.spref_nms.restype = C.c_int
        _lib.spref_covariance.restype = None
    return _lib


def nms(pts_sorted: np.ndarray, desc: np.ndarray | None, num_features: int, W: int, H: int, border: int = 8, r: int = 4):
    """-> (kps_xy [N,2], occ_grid [H/8,W/8] int16, descriptors [N,256] or None), the outputs of the reference's nms()."""
    pts = np.ascontiguousarray(pts_sorted, np.float32).reshape(-1, 2)
    n = len(pts)
    d = None if desc is None else np.ascontiguousarray(desc, np.float32).reshape(n, 256)
    kps = np.zeros((n + 1, 2), np.float32)
    occ = np.zeros((H // 8, W // 8), np.int16)
    dout = None if d is None else np.zeros((n + 1, 256), np.float32)
    vp = C.c_void_p
    N = _load().spref_nms(vp(pts.ctypes.data), vp(d.ctypes.data) if d is not None else None, n, int(num_features), int(border), int(r),
                          int(W), int(H), vp(kps.ctypes.data), vp(occ.ctypes.data), vp(dout.ctypes.data) if dout is not None else None, n + 1)
    assert N >= 0, N
    return kps[:N].copy(), occ, None if dout is None else dout[:N].copy()


def covariance(heat_inv: np.ndarray, kps_xy: np.ndarray):
    h = np.ascontiguousarray(heat_inv, np.float32)
    k = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    n = len(k)
    resp, cov2, cov2_inv = np.zeros(max(n, 1), np.float32), np.zeros((max(n, 1), 2), np.float32), np.zeros((max(n, 1), 2), np.float32)
    vp = C.c_void_p
    _load().spref_covariance(vp(h.ctypes.data), h.shape[0], h.shape[1], vp(k.ctypes.data), n, vp(resp.ctypes.data), vp(cov2.ctypes.data), vp(cov2_inv.ctypes.data))
    return resp[:n], cov2[:n], cov2_inv[:n]


# ---- the reference's own EdgeSE3ProjectDustOnlyPose (types_dust_tracking.{h,cpp}), oracle/_ref/libspdust_ref.so ----
DUST_LIB = os.path.join(_HERE, "_ref", "libspdust_ref.so")
_dust_lib = None


def dust_available() -> bool:
    return os.path.exists(DUST_LIB)


def dust_edges(dust, pose7, Xw, fx, fy, cx, cy, level=None):
    """computeError() + linearizeOplus() of one reference edge per map point at the vertex estimate ``pose7``.
    -> dict(level, err, uv, J, thrown)."""
    global _dust_lib
    if _dust_lib is None:
        _dust_lib = C.CDLL(DUST_LIB)
        _dust_lib.spref_dust_edges.restype = C.c_int
    dust = np.ascontiguousarray(dust, np.float32)
    Xw = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
    n = len(Xw)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    level = np.zeros(max(n, 1), np.uint8) if level is None else np.ascontiguousarray(level, np.uint8).copy()
    err, uv, J = np.zeros(max(n, 1)), np.zeros((max(n, 1), 2), np.float32), np.zeros((max(n, 1), 6))
    vp, dbl = C.c_void_p, C.c_double
    rc = _dust_lib.spref_dust_edges(vp(dust.ctypes.data), dust.shape[0], dust.shape[1], vp(pose7.ctypes.data), vp(Xw.ctypes.data), n,
                                    dbl(fx), dbl(fy), dbl(cx), dbl(cy), vp(level.ctypes.data), vp(err.ctypes.data), vp(uv.ctypes.data),
                                    vp(J.ctypes.data))
    return dict(level=level[:n], err=err[:n], uv=uv[:n], J=J[:n], thrown=rc != 0)


# ---- the reference's own guided-search loops (frame.cpp:382-474, sp_matcher.cpp:344-439, tracker_dust.cpp:105-172),
# ---- oracle/_ref/libspguided_ref.so
GUIDED_LIB = os.path.join(_HERE, "_ref", "libspguided_ref.so")
_guided_lib = None


def guided_available() -> bool:
    return os.path.exists(GUIDED_LIB)


def _guided():
    global _guided_lib
    if _guided_lib is None:
        _guided_lib = C.CDLL(GUIDED_LIB)
    return _guided_lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def features_in_area(occ, kp_un, x, y, r, min_x=0.0, min_y=0.0):
    occ = np.ascontiguousarray(occ, np.int16)
    kp_un = _f32(kp_un, (-1, 2))
    out = np.zeros(occ.size + 4, np.int32)
    vp, f = C.c_void_p, C.c_float
    n = _guided().spref_features_in_area(vp(occ.ctypes.data), occ.shape[0], occ.shape[1], vp(kp_un.ctypes.data), len(kp_un), f(x), f(y), f(r),
                                         f(min_x), f(min_y), vp(out.ctypes.data))
    return out[:n].copy()


def search_by_projection(qdesc, qxy, view_cos, occ, kp_un, kdesc, *, th, th_dist, in_view=None, bad=None, nobs=None, kp_taken=None,
                         c2_adaptive=0.0, min_x=0.0, min_y=0.0):
    """The reference's SPMatcher::SearchByProjection(Frame&, MapPoints, th, th_dist) -> (kp2mp [n], nmatches)."""
    qdesc, kdesc = _f32(qdesc, (-1, 256)), _f32(kdesc, (-1, 256))
    m, n = len(qdesc), len(kdesc)
    qxy, kp_un, view_cos = _f32(qxy, (m, 2)), _f32(kp_un, (n, 2)), _f32(view_cos, (m,))
    occ = np.ascontiguousarray(occ, np.int16)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    in_view, bad, kp_taken = u8(in_view), u8(bad), u8(kp_taken)
    nobs = None if nobs is None else np.ascontiguousarray(nobs, np.int32)
    kp2mp = np.full(max(n, 1), -1, np.int32)
    vp, f = C.c_void_p, C.c_float
    p = lambda a: None if a is None else vp(a.ctypes.data)
    nm = _guided().spref_search_by_projection(m, p(qdesc), p(qxy), p(view_cos), p(in_view), p(bad), p(nobs), p(kdesc), p(kp_un), n, p(occ),
                                              occ.shape[0], occ.shape[1], p(kp_taken), f(min_x), f(min_y), f(th), f(th_dist), f(c2_adaptive),
                                              p(kp2mp))
    return kp2mp[:n], nm


def dust_associate(qdesc, quv, occ, kdesc, *, in_view=None, bad=None):
    """The reference's dust-track association block -> (kp2mp [n], n_matches, dust_match [m])."""
    qdesc, kdesc = _f32(qdesc, (-1, 256)), _f32(kdesc, (-1, 256))
    m, n = len(qdesc), len(kdesc)
    quv = _f32(quv, (m, 2))
    occ = np.ascontiguousarray(occ, np.int16)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    in_view, bad = u8(in_view), u8(bad)
    kp2mp, dm = np.full(max(n, 1), -1, np.int32), np.zeros(max(m, 1), np.uint8)
    vp = C.c_void_p
    p = lambda a: None if a is None else vp(a.ctypes.data)
    nm = _guided().spref_dust_associate(m, p(qdesc), p(quv), p(in_view), p(bad), p(kdesc), n, p(occ), occ.shape[0], occ.shape[1], p(kp2mp), p(dm))
    return kp2mp[:n], nm, dm[:m]


def search_by_projection_last(qdesc, Xw, occ, kp_un, kdesc, *, th, Tcw_cur, Tcw_last, K, bounds, has_mp=None, outlier=None, nobs=None,
                              kp_taken=None):
    """The reference's SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono=true) -> (kp2mp [n], nmatches)."""
    qdesc, kdesc = _f32(qdesc, (-1, 256)), _f32(kdesc, (-1, 256))
    m, n = len(qdesc), len(kdesc)
    Xw, kp_un = _f32(Xw, (m, 3)), _f32(kp_un, (n, 2))
    occ = np.ascontiguousarray(occ, np.int16)
    u8 = lambda a: None if a is None else np.ascontiguousarray(a, np.uint8)
    has_mp, outlier, kp_taken = u8(has_mp), u8(outlier), u8(kp_taken)
    nobs = None if nobs is None else np.ascontiguousarray(nobs, np.int32)
    Tc, Tl, K, bounds = _f32(Tcw_cur, (4, 4)), _f32(Tcw_last, (4, 4)), _f32(K, (4,)), _f32(bounds, (4,))
    kp2mp = np.full(max(n, 1), -1, np.int32)
    vp = C.c_void_p
    p = lambda a: None if a is None else vp(a.ctypes.data)
    nm = _guided().spref_search_by_projection_last(m, p(qdesc), p(Xw), p(has_mp), p(outlier), p(nobs), p(Tc), p(Tl), p(K), p(bounds), p(kdesc),
                                                   p(kp_un), n, p(occ), occ.shape[0], occ.shape[1], p(kp_taken), C.c_float(th), p(kp2mp))
    return kp2mp[:n], nm


# ---- the reference's own SearchByBruteForce overloads (sp_matcher.cpp:1642-1674, sp_matcher_loop.cpp:334-376),
# ---- oracle/_ref/libspbf_ref.so
BF_LIB = os.path.join(_HERE, "_ref", "libspbf_ref.so")
_bf_lib = None


def bf_available() -> bool:
    return os.path.exists(BF_LIB)


def _bf():
    global _bf_lib
    if _bf_lib is None:
        _bf_lib = C.CDLL(BF_LIB)
        _bf_lib.spref_bruteforce_kf_frame.restype = None
    return _bf_lib


def bruteforce_kf_frame(desc1, has_mp1, bad1, desc2):
    """SearchByBruteForce(KeyFrame*, Frame&, vpMatches12) -> matches12 [len(desc2)]: key-frame row or -1."""
    d1, d2 = _f32(desc1, (-1, 256)), _f32(desc2, (-1, 256))
    h1, b1 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(bad1, np.uint8)
    out = np.full(max(len(d2), 1), -1, np.int32)
    vp = C.c_void_p
    _bf().spref_bruteforce_kf_frame(vp(d1.ctypes.data), vp(h1.ctypes.data), vp(b1.ctypes.data), len(d1), vp(d2.ctypes.data), len(d2), vp(out.ctypes.data))
    return out[:len(d2)]


def bruteforce_kf_kf(desc1, has_mp1, bad1, desc2, has_mp2, bad2):
    """SearchByBruteForce(KeyFrame*, KeyFrame*, vpMatches12) -> (matches12 [len(desc1)]: row of KF2 or -1, count)."""
    d1, d2 = _f32(desc1, (-1, 256)), _f32(desc2, (-1, 256))
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    h1, b1, h2, b2 = u8(has_mp1), u8(bad1), u8(has_mp2), u8(bad2)
    out = np.full(max(len(d1), 1), -1, np.int32)
    vp = C.c_void_p
    n = _bf().spref_bruteforce_kf_kf(vp(d1.ctypes.data), vp(h1.ctypes.data), vp(b1.ctypes.data), len(d1), vp(d2.ctypes.data), vp(h2.ctypes.data),
                                     vp(b2.ctypes.data), len(d2), vp(out.ctypes.data))
    return out[:len(d1)], n


FLANN_LIB = os.path.join(_HERE, "_ref", "libspflann_ref.so")
_flann_lib = None


def flann_available() -> bool:
    return os.path.exists(FLANN_LIB)


def search_tri_flann(desc1, has_mp1, kp1, cov2inv1, desc2, has_mp2, kp2, cov2inv2, F12, Cw1, R2w, t2w, intr2):
    """The reference's own SPMatcher::SearchForTriByFlann (sp_matcher.cpp:183-262, with CheckDistEpipolarLine :441-469)
    compiled verbatim, cv::FlannBasedMatcher replaced by an exact k-NN stand-in.  -> (pairs int64 [k, 2] = (row of KF1, row
    of KF2), nmatches)."""
    global _flann_lib
    if _flann_lib is None:
        _flann_lib = C.CDLL(FLANN_LIB)
    d1, d2 = _f32(desc1, (-1, 256)), _f32(desc2, (-1, 256))
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    h1, h2, k1, k2, c1, c2 = u8(has_mp1), u8(has_mp2), f32(kp1), f32(kp2), f32(cov2inv1), f32(cov2inv2)
    F, cw, R, t, intr = f32(F12), f32(Cw1), f32(R2w), f32(t2w), f32(intr2)
    pairs = np.zeros((max(len(d1), 1), 2), np.int64)
    npairs = C.c_int(0)
    vp = C.c_void_p
    n = _flann_lib.spref_search_tri_flann(vp(d1.ctypes.data), vp(h1.ctypes.data), vp(k1.ctypes.data), vp(c1.ctypes.data), len(d1),
                                          vp(d2.ctypes.data), vp(h2.ctypes.data), vp(k2.ctypes.data), vp(c2.ctypes.data), len(d2),
                                          vp(F.ctypes.data), vp(cw.ctypes.data), vp(R.ctypes.data), vp(t.ctypes.data), vp(intr.ctypes.data),
                                          vp(pairs.ctypes.data), C.byref(npairs))
    return pairs[:npairs.value].copy(), n
