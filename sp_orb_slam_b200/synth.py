"""Deterministic synthetic grayscale camera streams.

EuRoC / Tsukuba images are not available offline, so every test and bench line
runs on synthetic stand-ins of identical geometry: a mid-grey canvas with
random filled rectangles, triangles and thick lines, a 3x3 Gaussian blur
(sigma 0.8) and additive N(0, 2) sensor noise, clipped to u8.  A *stream* is
one scene viewed through a window that drifts a few pixels per frame, so that
frame t and frame t-1 share most of their corners (the situation
``SPMatcher::SearchByBruteForce`` is used in, reference
``orb_slam2/src/tracking/tracker.cpp:372-417``).

Pure numpy, no OpenCV: it must run identically in the build container and on
the GPU box.
"""
from __future__ import annotations

import numpy as np

_MARGIN = 48  # canvas margin so the viewing window can drift


def _gauss3(img: np.ndarray, sigma: float = 0.8) -> np.ndarray:
    k = np.exp(-0.5 * (np.arange(-1, 2) / sigma) ** 2)
    k = (k / k.sum()).astype(np.float32)
    p = np.pad(img, 1, mode="edge")
    tmp = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * tmp[:-2] + k[1] * tmp[1:-1] + k[2] * tmp[2:]


def render_scene(height: int, width: int, n_shapes: int, seed: int) -> np.ndarray:
    """Float32 canvas of (height + 2*margin, width + 2*margin), values 0..255."""
    rng = np.random.RandomState(seed)
    H, W = height + 2 * _MARGIN, width + 2 * _MARGIN
    img = np.full((H, W), 128.0, np.float32)
    for _ in range(n_shapes):
        kind = rng.randint(3)
        val = float(rng.randint(0, 256))
        cx, cy = rng.randint(0, W), rng.randint(0, H)
        if kind == 0:  # axis-aligned rectangle
            hw, hh = rng.randint(3, 40), rng.randint(3, 40)
            img[max(cy - hh, 0):cy + hh, max(cx - hw, 0):cx + hw] = val
        else:
            if kind == 1:  # triangle
                pts = np.stack([cx + rng.randint(-40, 41, 3), cy + rng.randint(-40, 41, 3)], 1)
            else:  # thick line == thin quadrilateral
                ang = rng.uniform(0, np.pi)
                ln, th = rng.randint(10, 80), rng.randint(1, 4)
                d = np.array([np.cos(ang), np.sin(ang)])
                n = np.array([-d[1], d[0]])
                c = np.array([cx, cy], np.float64)
                pts = np.stack([c - d * ln - n * th, c + d * ln - n * th,
                                c + d * ln + n * th, c - d * ln + n * th])
            x0, x1 = int(max(np.floor(pts[:, 0].min()), 0)), int(min(np.ceil(pts[:, 0].max()) + 1, W))
            y0, y1 = int(max(np.floor(pts[:, 1].min()), 0)), int(min(np.ceil(pts[:, 1].max()) + 1, H))
            if x1 <= x0 or y1 <= y0:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1]
            inside_pos = np.ones(yy.shape, bool)
            inside_neg = np.ones(yy.shape, bool)
            m = len(pts)
            for i in range(m):
                ax, ay = pts[i]
                bx, by = pts[(i + 1) % m]
                cr = (bx - ax) * (yy - ay) - (by - ay) * (xx - ax)
                inside_pos &= cr >= 0
                inside_neg &= cr <= 0
            mask = inside_pos | inside_neg
            img[y0:y1, x0:x1][mask] = val
    return img


def frame_from_scene(scene: np.ndarray, height: int, width: int, seed: int, t: int) -> np.ndarray:
    """Frame ``t`` of a stream: drifted window of the scene + blur + noise -> u8."""
    rng = np.random.RandomState((seed * 7919 + t * 104729 + 17) % (2 ** 31 - 1))
    # smooth drift of a few pixels per frame, bounded by the canvas margin
    ox = _MARGIN + int(round((_MARGIN - 8) * np.sin(0.11 * t + 0.3 * seed)))
    oy = _MARGIN + int(round((_MARGIN - 8) * np.sin(0.07 * t + 0.5 * seed)))
    win = scene[oy:oy + height, ox:ox + width]
    out = _gauss3(win) + rng.normal(0.0, 2.0, win.shape).astype(np.float32)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def make_frame(height: int, width: int, seed: int = 1234, n_shapes: int | None = None, t: int = 0) -> np.ndarray:
    """One u8 frame; ``n_shapes`` defaults to a density that gives ~500 keypoints at 752x480."""
    if n_shapes is None:
        n_shapes = max(8, int(400 * (height * width) / (752 * 480)))
    scene = render_scene(height, width, n_shapes, seed)
    return frame_from_scene(scene, height, width, seed, t)


def make_stream(height: int, width: int, n_frames: int, seed: int = 1234, n_shapes: int | None = None) -> np.ndarray:
    """[n_frames, height, width] u8: consecutive views of one drifting scene."""
    if n_shapes is None:
        n_shapes = max(8, int(400 * (height * width) / (752 * 480)))
    scene = render_scene(height, width, n_shapes, seed)
    return np.stack([frame_from_scene(scene, height, width, seed, t) for t in range(n_frames)])
