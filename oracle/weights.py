"""Weight readers for the oracle (test infrastructure, see oracle/__init__.py).

The reference loads ``orb_ros/data/models/superpoint.pt`` with ``torch::load``
(reference ``orb_slam2/src/cv/sp_extractor.cpp:354-355``).  That file is a
PyTorch-1.0 "legacy" TorchScript archive: a zip whose entries are *stored*
(uncompressed); ``superpoint/model.json`` maps ``convXX.{bias,weight}`` to
``superpoint/tensors/<id>`` (raw little-endian fp32, contiguous).  torch >= 2
refuses to load it, so it is parsed directly.

``.spw`` is this repo's own flat container for the same 24 tensors:
    magic 'SPW1' | u32 n | n x { char name[32] | u32 ndim | u32 dims[4] | f32 data[] }
"""
from __future__ import annotations

import json
import struct
import zipfile

import numpy as np

LAYERS = ["conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b",
          "conv4a", "conv4b", "convPa", "convPb", "convDa", "convDb"]


def read_legacy_pt(path: str) -> dict[str, np.ndarray]:
    z = zipfile.ZipFile(path)
    root = z.namelist()[0].split("/")[0]
    meta = json.loads(z.read(f"{root}/model.json"))
    tensors = meta["tensors"]
    out: dict[str, np.ndarray] = {}
    for sub in meta["mainModule"]["submodules"]:
        for p in sub.get("parameters", []):
            t = tensors[int(p["tensorId"])]
            assert t["dataType"] == "FLOAT"
            dims = [int(d) for d in t["dims"]]
            raw = z.read(f"{root}/{t['data']['key']}")
            arr = np.frombuffer(raw, "<f4", count=int(np.prod(dims)), offset=int(t["offset"]) * 4)
            out[f"{sub['name']}.{p['name']}"] = arr.reshape(dims).copy()
    return out


def write_spw(path: str, w: dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"SPW1" + struct.pack("<I", len(w)))
        for name, a in w.items():
            a = np.ascontiguousarray(a, "<f4")
            dims = list(a.shape) + [1] * (4 - a.ndim)
            f.write(name.encode().ljust(32, b"\0") + struct.pack("<5I", a.ndim, *dims))
            f.write(a.tobytes())


def read_spw(path: str) -> dict[str, np.ndarray]:
    buf = open(path, "rb").read()
    assert buf[:4] == b"SPW1", "not an SPW1 file"
    (n,) = struct.unpack_from("<I", buf, 4)
    off, out = 8, {}
    for _ in range(n):
        name = buf[off:off + 32].rstrip(b"\0").decode()
        ndim, *dims = struct.unpack_from("<5I", buf, off + 32)
        off += 52
        cnt = int(np.prod(dims[:ndim]))
        out[name] = np.frombuffer(buf, "<f4", cnt, off).reshape(dims[:ndim]).copy()
        off += 4 * cnt
    return out


def random_weights(seed: int = 0) -> dict[str, np.ndarray]:
    """Random-init weights of the reference architecture (sp_extractor.cpp:16-43)."""
    plan = [("conv1a", 1, 64, 3), ("conv1b", 64, 64, 3), ("conv2a", 64, 64, 3), ("conv2b", 64, 64, 3),
            ("conv3a", 64, 128, 3), ("conv3b", 128, 128, 3), ("conv4a", 128, 128, 3), ("conv4b", 128, 128, 3),
            ("convPa", 128, 256, 3), ("convPb", 256, 65, 1), ("convDa", 128, 256, 3), ("convDb", 256, 256, 1)]
    rng = np.random.RandomState(seed)
    out = {}
    for name, ci, co, k in plan:
        std = np.sqrt(2.0 / (ci * k * k))
        out[f"{name}.bias"] = (rng.randn(co) * 0.05).astype(np.float32)
        out[f"{name}.weight"] = (rng.randn(co, ci, k, k) * std).astype(np.float32)
    return out
