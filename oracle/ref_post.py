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
