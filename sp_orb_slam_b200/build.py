"""In-tree build of libspfe.so (nvcc, sm_100a only) and of the C++ shim demo."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libspfe.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return [os.path.join(CSRC, "spfe.cu"), os.path.join(CSRC, "weights.cc")]


def _deps():
    deps = [os.path.join(ROOT, "include", "spfe.h")]
    for d, _, fs in os.walk(CSRC):
        deps += [os.path.join(d, f) for f in fs]
    return deps


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB, *_sources()]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


def build_shim(force: bool = False) -> str:
    """C++ shim (orbslam::SPExtractor / SPMatcher over the C ABI) + its self-test binary."""
    out = os.path.join(LIB_DIR, "shim_selftest")
    src = os.path.join(PKG, "cpp", "shim_selftest.cc")
    shim = os.path.join(PKG, "cpp", "sp_shim.cc")
    deps = [src, shim, os.path.join(PKG, "cpp", "sp_extractor.h"), os.path.join(PKG, "cpp", "sp_matcher.h"),
            os.path.join(PKG, "cpp", "mini_cv.h"), os.path.join(PKG, "cpp", "optimizer_dust.h"),
            os.path.join(PKG, "cpp", "dust_pose_selftest.cc"), LIB]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
                           "-o", out, src, shim, "-L", LIB_DIR, "-lspfe", "-Wl,-rpath,$ORIGIN"])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "cpp"),
                           "-o", os.path.join(LIB_DIR, "dust_pose_selftest"), os.path.join(PKG, "cpp", "dust_pose_selftest.cc"), shim,
                           "-L", LIB_DIR, "-lspfe", "-Wl,-rpath,$ORIGIN"])
    build_stream_bench(force=True)
    return out


def build_stream_bench(force: bool = False) -> str:
    """Native throughput harness over the C ABI (cpp/stream_bench.cc)."""
    out = os.path.join(LIB_DIR, "stream_bench")
    src = os.path.join(PKG, "cpp", "stream_bench.cc")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", out, src,
                           "-L", LIB_DIR, "-lspfe", "-Wl,-rpath,$ORIGIN"])
    return out


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose=True)
    print(LIB)
