"""CPU oracle for the SuperPoint extract + match hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status: **parity unpinned by the reference** -- the reference ships no
tests, golden vectors or fixtures for this path.  The
oracle is instead pinned against (i) the reference's own code compiled in this
container by ``oracle/ref_build.sh`` into ``oracle/_ref/``: its ``SPFrontend``
against libtorch (network half), its ``nms`` / ``computeCovariance`` verbatim
against a cv / Eigen stand-in (``ref_cv_stub.h``) and its
``EdgeSE3ProjectDustOnlyPose`` verbatim against a g2o stand-in
(``ref_g2o_stub.h``) and its guided-search loops (``Frame::GetFeaturesInArea``,
``SPMatcher::SearchByProjection(Frame&, MapPoints)`` and ``(Cur, Last)``, the dust-track association
block) and both ``SearchByBruteForce`` overloads verbatim against class skeletons
(``ref_guided_driver.cc``, ``ref_bf_driver.cc``), and (ii) OpenCV (``cv2.BFMatcher``, ``cv2.sortIdx``,
``cv2.minMaxLoc``) for the third-party arithmetic the reference calls; the
resulting vectors are committed under ``tests/golden/``.
"""
