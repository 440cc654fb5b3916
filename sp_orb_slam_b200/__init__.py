"""B200-native SuperPoint front-end (extract + brute-force match) for sp_orb_slam.

The product is the C-ABI library ``lib/libspfe.so`` (hand-written sm_100a CUDA,
``include/spfe.h``) plus the C++ shim in ``cpp/`` that re-creates the reference's
``orbslam::SPExtractor`` / ``orbslam::SPMatcher``.  This Python package is the
host-side mirror used by the tests and the benchmark.
"""
from .extractor import Optimizer, SPExtractor, SPMatcher, SpfeError  # noqa: F401

__all__ = ["Optimizer", "SPExtractor", "SPMatcher", "SpfeError"]
