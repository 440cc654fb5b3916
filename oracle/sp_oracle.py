"""CPU oracle for SPExtractor / SPMatcher -- TEST INFRASTRUCTURE ONLY.

fp32 restatement of the reference path, function by function:

  frontend_forward   <- orb_slam2/src/cv/sp_extractor.cpp:79-159  (SPFrontend::forward)
  extract            <- orb_slam2/src/cv/sp_extractor.cpp:361-514 (SPExtractor::operator())
  match_mutual_nn    <- orb_slam2/src/cv/sp_matcher.cpp:1642-1674 (SearchByBruteForce core)
  dust_linearize / dust_pose_optimize
                     <- orb_slam2/src/optimization/types_dust_tracking.cpp:36-141 (EdgeSE3ProjectDustOnlyPose) and
                        orb_slam2/src/mapping/optimizer_dust.cpp:170-293 (PoseOptimizationDust); C in oracle/dust_pose.c

The network half runs on torch CPU (the reference runs the same ATen ops
through libtorch); the post-processing half is the C restatement in
``oracle/sp_post.c`` (built by ``oracle/build.py`` with gcc).  Parity status:
unpinned by the reference (no tests / goldens exist upstream); see
``oracle/__init__.py`` for what this oracle is pinned against instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCORE_THRESH = 0.007   # sp_extractor.cpp:122 (hard-coded)
NMS_RADIUS = 4         # sp_extractor.cpp:502
BORDER = 8             # sp_extractor.cpp:502
CELL = 8               # sp_extractor.cpp:354
HEAT_CLAMP = 0.001     # sp_extractor.cpp:129


def build_post(force: bool = False) -> str:
    """Compile oracle/sp_post.c -> oracle/_build/libsporacle.so (gcc, no deps)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsporacle.so")
    srcs = [os.path.join(_HERE, "sp_post.c"), os.path.join(_HERE, "dust_pose.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, *srcs, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_post())
        _LIB.orc_nms.restype = C.c_int
        _LIB.orc_l2.restype = C.c_float
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ----------------------------------------------------------------------------
# network half (torch CPU fp32)
# ----------------------------------------------------------------------------
def frontend_forward(weights: dict, img_u8: np.ndarray, keep_layers: bool = False) -> dict:
    """SPFrontend::forward on one H x W u8 frame (H, W multiples of 8)."""
    import torch
    import torch.nn.functional as F

    H, W = img_u8.shape
    hc, wc = H // CELL, W // CELL
    w = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()}
    layers = {}

    def conv(x, name, pad):
        return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=1, padding=pad)

    with torch.no_grad():
        # sp_extractor.cpp:386-390: convertTo(CV_32FC1, 1/255) then [1,1,H,W]
        x = (torch.from_numpy(img_u8.astype(np.float32)) * np.float32(1.0 / 255.0)).view(1, 1, H, W)
        for blk, pool in (("1", True), ("2", True), ("3", True), ("4", False)):   # :81-94
            x = torch.relu(conv(x, f"conv{blk}a", 1))
            if keep_layers:
                layers[f"conv{blk}a"] = x[0].numpy().copy()
            x = torch.relu(conv(x, f"conv{blk}b", 1))
            if pool:
                x = F.max_pool2d(x, 2, 2)
            if keep_layers:
                layers[f"conv{blk}b"] = x[0].numpy().copy()
        cPa = torch.relu(conv(x, "convPa", 1))                                     # :96
        semi = conv(cPa, "convPb", 0).squeeze()                                    # :97  [65,hc,wc]
        cDa = torch.relu(conv(x, "convDa", 1))                                     # :99
        coarse = conv(cDa, "convDb", 0)                                            # :100 [1,256,hc,wc]
        dn = torch.norm(coarse, 2, 1)                                              # :102
        coarse = coarse.div(torch.unsqueeze(dn, 1))                                # :103
        dense = torch.softmax(semi, 0)                                             # :105
        semi_dust, dense_dust, nodust = semi[-1], dense[-1], dense[:-1]            # :106-108
        score_map, indices = nodust.max(0)                                         # :112-114
        # :64-73,117-119: candidate pixel of cell (cy,cx) with channel i -> (cx*8 + i%8, cy*8 + i//8)
        cy, cx = torch.meshgrid(torch.arange(hc), torch.arange(wc), indexing="ij")
        px = (cx * CELL + indices % CELL)
        py = (cy * CELL + indices // CELL)
        mask = score_map >= SCORE_THRESH                                           # :122
        pixels_in = torch.stack([px[mask], py[mask]]).to(torch.float32)            # :123-125 [2,n]
        score = score_map[mask]                                                    # :126
        heat_log = F.pixel_shuffle(torch.log(torch.clamp(nodust, min=HEAT_CLAMP)).unsqueeze(0), CELL)  # :129-131
        x_s = pixels_in[0].div(W / 2.0) - 1.0                                      # :137
        y_s = pixels_in[1].div(H / 2.0) - 1.0                                      # :138
        samp = torch.stack([x_s, y_s], -1).view(1, 1, -1, 2)                       # :142-143
        if samp.shape[2] > 0:
            desc = F.grid_sample(coarse, samp, mode="bilinear", padding_mode="zeros",
                                 align_corners=True).squeeze(2).squeeze(0)         # :145-146 [256,n]
            desc = desc.div(torch.norm(desc, 2, 0, True))                          # :148
        else:
            desc = torch.zeros(256, 0)
    out = dict(semi_dust=semi_dust.numpy().copy(), dense_dust=dense_dust.numpy().copy(),
               pixels_in=pixels_in.numpy().copy(), score=score.numpy().copy(),
               desc_sampled=desc.numpy().copy(), heat_log=heat_log[0, 0].numpy().copy(),
               # extra taps for kernel bring-up / parity margins (not reference outputs)
               semi=semi.numpy().copy(), coarse=coarse[0].numpy().copy(),
               score_map=score_map.numpy().copy(), argmax=indices.numpy().astype(np.int32),
               nodust=nodust.numpy().copy())
    if keep_layers:
        layers["convPa"] = cPa[0].numpy().copy()
        layers["convDa"] = cDa[0].numpy().copy()
        out["layers"] = layers
    return out


# ----------------------------------------------------------------------------
# post-processing half (C restatement)
# ----------------------------------------------------------------------------
def to_heat(heat_log: np.ndarray):
    h = np.ascontiguousarray(heat_log, np.float32)
    heat, heat_inv = np.empty_like(h), np.empty_like(h)
    mn, mx = C.c_double(), C.c_double()
    _lib().orc_to_heat(_p(h, C.c_float), h.size, _p(heat, C.c_float), _p(heat_inv, C.c_float),
                       C.byref(mn), C.byref(mx))
    return heat, heat_inv, mn.value, mx.value


def sort_desc(score: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(score, np.float32)
    order = np.empty(s.size, np.int32)
    _lib().orc_sort_desc(_p(s, C.c_float), s.size, _p(order, C.c_int32))
    return order


def nms(pts_sorted: np.ndarray, num_features: int, W: int, H: int, border: int = BORDER, r: int = NMS_RADIUS):
    """pts_sorted: n x 2 float32 (x, y) in descending score order -> (sel indices, occ_grid)."""
    pts = np.ascontiguousarray(pts_sorted, np.float32).reshape(-1, 2)
    sel = np.empty(max(len(pts), 1), np.int32)
    occ = np.empty((H // CELL, W // CELL), np.int16)
    n = _lib().orc_nms(_p(pts, C.c_float), len(pts), num_features, border, r, W, H,
                       _p(sel, C.c_int32), _p(occ, C.c_int16))
    return sel[:n].copy(), occ


def covariance(heat_inv: np.ndarray, kps_xy: np.ndarray):
    h = np.ascontiguousarray(heat_inv, np.float32)
    k = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    n = len(k)
    resp = np.empty(n, np.float32)
    cov2 = np.empty((n, 2), np.float32)
    cov2_inv = np.empty((n, 2), np.float32)
    _lib().orc_covariance(_p(h, C.c_float), h.shape[0], h.shape[1], _p(k, C.c_float), n,
                          _p(resp, C.c_float), _p(cov2, C.c_float), _p(cov2_inv, C.c_float))
    return resp, cov2, cov2_inv


def postprocess(fwd: dict, H: int, W: int, num_features: int) -> dict:
    """Everything SPExtractor::operator() does after the forward (:426-513)."""
    heat, heat_inv, mn, mx = to_heat(fwd["heat_log"])
    pts = np.ascontiguousarray(fwd["pixels_in"].T)            # n x 2
    desc = np.ascontiguousarray(fwd["desc_sampled"].T)        # n x 256
    order = sort_desc(fwd["score"])
    sel, occ = nms(pts[order], num_features, W, H)
    src = order[sel]                                          # index into the unsorted candidate list
    kps = pts[src].copy()
    resp, cov2, cov2_inv = covariance(heat_inv, kps)
    return dict(n=len(src), kp_xy=kps, kp_response=resp, desc=desc[src].copy(), occ_grid=occ,
                score=fwd["score"][src].copy(), cov2=cov2, cov2_inv=cov2_inv,
                heat=heat, heat_inv=heat_inv, heat_min=mn, heat_max=mx,
                dense_dust=fwd["dense_dust"], semi_dust=fwd["semi_dust"])


def extract(weights: dict, img_u8: np.ndarray, num_features: int = 800, keep_forward: bool = False) -> dict:
    if img_u8 is None or img_u8.size == 0:
        raise RuntimeError("input image is empty")            # sp_extractor.cpp:364-365
    assert img_u8.dtype == np.uint8 and img_u8.ndim == 2      # :368
    H, W = img_u8.shape
    fwd = frontend_forward(weights, img_u8)
    out = postprocess(fwd, H, W, num_features)
    if keep_forward:
        out["forward"] = fwd
    return out


# ----------------------------------------------------------------------------
# matcher
# ----------------------------------------------------------------------------
def l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32).ravel()
    b = np.ascontiguousarray(b, np.float32).ravel()
    return float(_lib().orc_l2(_p(a, C.c_float), _p(b, C.c_float), a.size))


def match_mutual_nn(q: np.ndarray, t: np.ndarray):
    """-> (q2t int32[nq] with -1 = unmatched, dist float32[nq], second-best dist float32[nq])."""
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 256) if q.size else np.zeros((0, 256), np.float32)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 256) if t.size else np.zeros((0, 256), np.float32)
    q2t = np.empty(len(q), np.int32)
    dist = np.empty(len(q), np.float32)
    sec = np.empty(len(q), np.float32)
    _lib().orc_match_mutual(_p(q, C.c_float), len(q), _p(t, C.c_float), len(t), 256,
                            _p(q2t, C.c_int32), _p(dist, C.c_float), _p(sec, C.c_float))
    return q2t, dist, sec


# ----------------------------------------------------------------------------
# guided (cell-grid) searches -- SURVEY.md section 8(f) rank 2
# ----------------------------------------------------------------------------
FLT_MAX = float(np.finfo(np.float32).max)
TH_HIGH, TH_LOW = 0.7, 0.3      # sp_matcher.cpp:18-19


def features_in_area(occ: np.ndarray, kp_un: np.ndarray, x: float, y: float, r: float, min_x: float = 0.0, min_y: float = 0.0):
    """Frame::GetFeaturesInArea (frame.cpp:382-420) -> keypoint indices in the reference's order."""
    occ = np.ascontiguousarray(occ, np.int16)
    kp_un = np.ascontiguousarray(kp_un, np.float32).reshape(-1, 2)
    out = np.empty(occ.size + 4, np.int32)
    L = _lib()
    L.orc_features_in_area.restype = C.c_int
    n = L.orc_features_in_area(_p(occ, C.c_int16), occ.shape[0], occ.shape[1], _p(kp_un, C.c_float), C.c_float(x), C.c_float(y),
                               C.c_float(r), C.c_float(min_x), C.c_float(min_y), _p(out, C.c_int32))
    return out[:n].copy()


def search_guided(qdesc, qxy, qr, occ, kp_un, kdesc, *, mode: int, best_init: float, th_le: float, th_lt: float, c2: float = 0.0,
                  qvalid=None, qblocks=None, kp_taken=None, min_x: float = 0.0, min_y: float = 0.0):
    """Generic greedy guided search (orc_search_guided).  -> (q2kp int32[m], qdist f32[m], kp_taken_after u8[n])."""
    qdesc = np.ascontiguousarray(qdesc, np.float32).reshape(-1, 256)
    m = len(qdesc)
    qxy = np.ascontiguousarray(qxy, np.float32).reshape(m, 2)
    qr = np.ascontiguousarray(np.broadcast_to(np.asarray(qr, np.float32), (m,)))
    occ = np.ascontiguousarray(occ, np.int16)
    kp_un = np.ascontiguousarray(kp_un, np.float32).reshape(-1, 2)
    kdesc = np.ascontiguousarray(kdesc, np.float32).reshape(-1, 256)
    n = len(kdesc)
    qvalid = np.ones(m, np.uint8) if qvalid is None else np.ascontiguousarray(qvalid, np.uint8)
    qblocks = np.ones(m, np.uint8) if qblocks is None else np.ascontiguousarray(qblocks, np.uint8)
    taken = np.zeros(n, np.uint8) if kp_taken is None else np.array(kp_taken, np.uint8)
    q2kp = np.empty(max(m, 1), np.int32)
    qdist = np.empty(max(m, 1), np.float32)
    _lib().orc_search_guided(m, _p(qdesc, C.c_float), _p(qvalid, C.c_uint8), _p(qblocks, C.c_uint8), _p(qxy, C.c_float),
                             _p(qr, C.c_float), mode, _p(occ, C.c_int16), occ.shape[0], occ.shape[1], _p(kp_un, C.c_float),
                             _p(kdesc, C.c_float), n, _p(taken, C.c_uint8), C.c_float(min_x), C.c_float(min_y),
                             C.c_float(best_init), C.c_float(th_le), C.c_float(th_lt), C.c_float(c2),
                             _p(q2kp, C.c_int32), _p(qdist, C.c_float))
    return q2kp[:m], qdist[:m], taken


def search_by_projection_map_points(qdesc, proj_xy, radius, occ, kp_un, kdesc, *, th_dist: float, in_view=None, observed=None,
                                    kp_taken=None, c2_adaptive: float = 0.0):
    """SPMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, th_dist), sp_matcher.cpp:344-432.
    radius[i] = RadiusByViewingCos(mTrackViewCos) * th * mvScaleFactors[level] (host-side, :362-374)."""
    return search_guided(qdesc, proj_xy, radius, occ, kp_un, kdesc, mode=0, best_init=256.0, th_le=th_dist, th_lt=0.7,
                         c2=c2_adaptive, qvalid=in_view, qblocks=observed, kp_taken=kp_taken)


def search_by_projection_last_frame(qdesc, proj_xy, radius, occ, kp_un, kdesc, *, valid=None, observed=None, kp_taken=None):
    """SPMatcher::SearchByProjection(Frame &Cur, const Frame &Last, th, bMono), sp_matcher.cpp:1439-1543 (monocular:
    mvuRight < 0, so the stereo test at :1512-1517 never fires)."""
    return search_guided(qdesc, proj_xy, radius, occ, kp_un, kdesc, mode=0, best_init=FLT_MAX, th_le=TH_HIGH, th_lt=-np.inf,
                         qvalid=valid, qblocks=observed, kp_taken=kp_taken)


def dust_associate(qdesc, proj_uv_cells, occ, kdesc, *, in_view=None):
    """Patch-wise association of dust tracking, tracker_dust.cpp:112-172: 2x2 occ_grid cells at floor(dust_proj),
    best descriptor distance < 0.75, the matched cell is cleared."""
    n = len(np.asarray(kdesc).reshape(-1, 256))
    return search_guided(qdesc, proj_uv_cells, 0.0, occ, np.zeros((n, 2), np.float32), kdesc, mode=1, best_init=0.75,
                         th_le=-np.inf, th_lt=0.75, qvalid=in_view)


def knn2(q: np.ndarray, t: np.ndarray):
    """Exact 2-NN under L2 (orc_knn2) -> (idx int32[nq,2], dist f32[nq,2])."""
    q = np.ascontiguousarray(q, np.float32).reshape(-1, 256)
    t = np.ascontiguousarray(t, np.float32).reshape(-1, 256) if np.size(t) else np.zeros((0, 256), np.float32)
    idx = np.empty((max(len(q), 1), 2), np.int32)
    dist = np.empty((max(len(q), 1), 2), np.float32)
    _lib().orc_knn2(_p(q, C.c_float), len(q), _p(t, C.c_float), len(t), 256, _p(idx, C.c_int32), _p(dist, C.c_float))
    return idx[:len(q)], dist[:len(q)]


# ----------------------------------------------------------------------------
# dust-map pose optimisation -- C restatement in oracle/dust_pose.c
# ----------------------------------------------------------------------------
class _DustCam(C.Structure):
    _fields_ = [("dust", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("huber", C.c_double)]


def _dust_cam(dust, fx, fy, cx, cy, huber):
    dust = np.ascontiguousarray(dust, np.float32)
    return dust, _DustCam(dust.ctypes.data, dust.shape[0], dust.shape[1], fx, fy, cx, cy, huber)


def dust_linearize(dust, pose7, Xw, fx, fy, cx, cy, *, huber: float = 0.9, level=None):
    """One computeActiveErrors + buildSystem pass of the PoseOptimizationDust graph at ``pose7`` = (qx, qy, qz, qw, tx,
    ty, tz).  -> dict(level, err, uv, J, H, b, chi2, thrown)."""
    dust, cam = _dust_cam(dust, fx, fy, cx, cy, huber)
    Xw = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
    n = len(Xw)
    pose7 = np.ascontiguousarray(pose7, np.float64)
    level = np.zeros(n, np.uint8) if level is None else np.ascontiguousarray(level, np.uint8).copy()
    err, uv, J, Hb = np.zeros(n), np.zeros((n, 2), np.float32), np.zeros((n, 6)), np.zeros(43)
    vp = C.c_void_p
    rc = _lib().orc_dust_linearize(C.byref(cam), vp(pose7.ctypes.data), vp(Xw.ctypes.data), n, vp(level.ctypes.data),
                                   vp(err.ctypes.data), vp(uv.ctypes.data), vp(J.ctypes.data), vp(Hb.ctypes.data))
    return dict(level=level, err=err, uv=uv, J=J, H=Hb[:36].reshape(6, 6).copy(), b=Hb[36:42].copy(), chi2=float(Hb[42]), thrown=rc != 0)


def dust_pose_optimize(dust, pose7, Xw, fx, fy, cx, cy, *, huber: float = 0.9, iterations: int = 40, chi2_inlier: float = 0.9):
    """optimizer.optimize(40) on the PoseOptimizationDust graph + its inlier read-out.
    -> dict(pose, visible, uv, n_inlier, n_iter (-1: linearizeOplus threw), err, level, stats=(lambda, chi2, trials))."""
    dust, cam = _dust_cam(dust, fx, fy, cx, cy, huber)
    Xw = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3)
    n = len(Xw)
    pose = np.ascontiguousarray(pose7, np.float64).copy()
    level, err, uv, vis = np.zeros(n, np.uint8), np.zeros(n), np.zeros((n, 2), np.float32), np.zeros(n, np.uint8)
    ninl, stats = C.c_int(0), np.zeros(3)
    vp = C.c_void_p
    it = _lib().orc_dust_optimize(C.byref(cam), vp(pose.ctypes.data), vp(Xw.ctypes.data), n, int(iterations), C.c_double(chi2_inlier),
                                  vp(level.ctypes.data), vp(err.ctypes.data), vp(uv.ctypes.data), vp(vis.ctypes.data),
                                  C.byref(ninl), vp(stats.ctypes.data))
    return dict(pose=pose, visible=vis, uv=uv, n_inlier=ninl.value, n_iter=int(it), err=err, level=level, stats=stats)
