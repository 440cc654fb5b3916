"""ctypes wrapper around oracle/_ref/libspref.so = the REFERENCE's own
``SPFrontend`` (sp_extractor.cpp:16-159) compiled against this image's libtorch
by ``oracle/ref_build.sh``.  Test infrastructure only; used to pin
``oracle/sp_oracle.frontend_forward`` and as the ``kind: "reference"`` CPU
baseline of bench.py."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .weights import LAYERS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libspref.so")
_lib = None
_loaded_for = None


def available() -> bool:
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (makes libtorch's dependencies resident)
        _lib = C.CDLL(LIB)
    return _lib


def forward(weights: dict, img_u8: np.ndarray, threads: int = 0) -> dict:
    global _loaded_for
    L = _load()
    H, W = img_u8.shape
    key = (id(weights), H, W, threads)
    if _loaded_for != key:
        arrs = []
        for name in LAYERS:
            arrs += [np.ascontiguousarray(weights[f"{name}.weight"], np.float32), np.ascontiguousarray(weights[f"{name}.bias"], np.float32)]
        ptrs = (C.c_void_p * 24)(*[a.ctypes.data for a in arrs])
        assert L.spref_load(ptrs, H, W, threads) == 0
        _loaded_for = key
    hc, wc = H // 8, W // 8
    mx = hc * wc
    semi_dust = np.empty((hc, wc), np.float32)
    dense_dust = np.empty((hc, wc), np.float32)
    pixels = np.empty(2 * mx, np.float32)
    score = np.empty(mx, np.float32)
    desc = np.empty(256 * mx, np.float32)
    heat = np.empty((H, W), np.float32)
    img = np.ascontiguousarray(img_u8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = L.spref_forward(p(img), p(semi_dust), p(dense_dust), p(pixels), p(score), p(desc), p(heat), mx)
    assert n >= 0
    return dict(semi_dust=semi_dust, dense_dust=dense_dust, pixels_in=pixels[:2 * n].reshape(2, n).copy(),
                score=score[:n].copy(), desc_sampled=desc[:256 * n].reshape(256, n).copy(), heat_log=heat)
