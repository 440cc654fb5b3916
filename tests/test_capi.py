"""CPU tests of the C-ABI library: it loads, exports every symbol include/spfe.h
declares, validates arguments, parses weight files and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, WEIGHTS
from sp_orb_slam_b200 import SPExtractor, SPMatcher, SpfeError, capi


@pytest.fixture(scope="module")
def lib():
    return capi.load()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "spfe.h")).read()
    declared = sorted(set(re.findall(r"\b(spfe_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"libspfe.so does not export {name}"
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts(lib):
    cfg = capi.Config()
    lib.spfe_default_config(C.byref(cfg), 480, 752, 800)
    assert cfg.struct_size == C.sizeof(capi.Config)
    assert (cfg.height, cfg.width, cfg.max_keypoints) == (480, 752, 800)
    assert abs(cfg.score_thresh - 0.007) < 1e-9 and cfg.nms_radius == 4 and cfg.border == 8   # sp_extractor.cpp:122,502
    assert cfg.flags == capi.EMIT_HEAT | capi.EMIT_HEAT_INV | capi.EMIT_COV   # everything operator() fills


def test_create_validates_arguments(lib):
    ctx = C.c_void_p()
    cfg = capi.Config()
    lib.spfe_default_config(C.byref(cfg), 480, 750, 800)          # width not a multiple of 8
    cfg.weights_path = WEIGHTS.encode()
    assert lib.spfe_create(C.byref(cfg), C.byref(ctx)) == capi.ERR_INVALID
    assert b"multiples of 8" in lib.spfe_last_error(None)
    lib.spfe_default_config(C.byref(cfg), 480, 752, 800)
    assert lib.spfe_create(C.byref(cfg), C.byref(ctx)) == capi.ERR_WEIGHTS   # NULL path
    cfg.struct_size = 8
    assert lib.spfe_create(C.byref(cfg), C.byref(ctx)) == capi.ERR_INVALID
    assert not ctx.value


def test_no_cpu_fallback():
    """Without a B200 the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(SpfeError) as e:
        SPExtractor(800, 480, 752, WEIGHTS)
    assert e.value.code == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_weight_readers(lib, weights):
    err = C.create_string_buffer(256)
    assert lib.spfe_check_weights(WEIGHTS.encode(), err, 256) == 1300865
    assert lib.spfe_check_weights(b"/nonexistent/superpoint.pt", err, 256) == capi.ERR_WEIGHTS and b"cannot open" in err.value
    ref = "/root/reference/orb_ros/data/models/superpoint.pt"
    if os.path.exists(ref):                                   # the reference's own legacy archive, read by the C++ parser
        assert lib.spfe_check_weights(ref.encode(), err, 256) == 1300865


def test_weight_reader_rejects_garbage(lib, tmp_path):
    err = C.create_string_buffer(256)
    p = tmp_path / "bad.spw"
    p.write_bytes(b"SPW1" + b"\x05\x00\x00\x00" + b"x" * 40)
    assert lib.spfe_check_weights(str(p).encode(), err, 256) == capi.ERR_WEIGHTS
    p.write_bytes(b"not a model file at all")
    assert lib.spfe_check_weights(str(p).encode(), err, 256) == capi.ERR_WEIGHTS
    from oracle import weights as OW
    w = OW.random_weights(0)
    del w["convDb.bias"]
    OW.write_spw(str(p), w)
    assert lib.spfe_check_weights(str(p).encode(), err, 256) == capi.ERR_WEIGHTS and b"convDb" in err.value


def test_descriptor_distance_host():
    rng = np.random.RandomState(0)
    a, b = rng.randn(256).astype(np.float32), rng.randn(256).astype(np.float32)
    d = SPMatcher.DescriptorDistance(a, b)
    assert abs(d - float(np.sqrt(((a - b) ** 2).sum()))) < 1e-4
    from oracle import sp_oracle as O
    assert d == O.l2(a, b)                                      # same fp32 summation order
    assert (SPMatcher.TH_HIGH, SPMatcher.TH_LOW, SPMatcher.HISTO_LENGTH) == (0.7, 0.3, 30)


def test_shim_compiles():
    """The C++ drop-in classes (orbslam::SPExtractor / SPMatcher) build against the C ABI."""
    from sp_orb_slam_b200 import build
    exe = build.build_shim()
    assert os.path.exists(exe)


def test_base_extractor_matches_reference(tmp_path):
    """The shim's stand-alone copy of BaseExtractor (scale-pyramid getters that Frame reads, frame.cpp:211-217) against the
    REFERENCE's own class compiled verbatim (oracle/_ref/ref_base_probe, base_extractor.h:7-95): the SuperPoint
    configuration (1 level, factor 1.0) and ORB-style pyramids print the same levels, factors, sigmas and per-level
    feature budgets."""
    import subprocess
    from sp_orb_slam_b200 import build
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_base_probe")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ref_base_probe not built (run oracle/ref_build.sh where /root/reference exists)")
    build.build_lib()
    exe = str(tmp_path / "base_probe_shim")
    cpp = os.path.join(ROOT, "sp_orb_slam_b200", "cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", cpp,
                           os.path.join(ROOT, "tests", "cpp", "base_probe_shim.cc"), os.path.join(cpp, "sp_shim.cc"),
                           "-o", exe, "-L", build.LIB_DIR, "-lspfe", f"-Wl,-rpath,{build.LIB_DIR}"])
    for cfg in (("800", "1.0", "1"), ("2000", "1.0", "1"), ("1000", "1.2", "8"), ("2000", "1.2", "8"), ("500", "1.5", "4"), ("1200", "2.0", "3"), ("7", "1.1", "12")):
        a = subprocess.run([ref, *cfg], capture_output=True, text=True, check=True).stdout
        b = subprocess.run([exe, *cfg], capture_output=True, text=True, check=True).stdout
        assert a == b, (cfg, a, b)


def test_abi_struct_layouts_match_ctypes(tmp_path):
    """sizeof / offsetof of every struct in include/spfe.h as gcc lays them out, against the ctypes mirrors in
    sp_orb_slam_b200/capi.py (field names, order, offsets, total size)."""
    import subprocess
    structs = {"spfe_config": capi.Config, "spfe_frame_out": capi.FrameOut, "spfe_guided_search": capi.GuidedSearch,
               "spfe_dust_pose": capi.DustPose, "spfe_stage_time": capi.StageTime}
    src = ['#include <stddef.h>', '#include <stdio.h>', '#include "spfe.h"', 'int main(void) {']
    for cname, ct in structs.items():
        src.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, *_ in ct._fields_:
            src.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src += ['  return 0;', '}']
    (tmp_path / "probe.c").write_text("\n".join(src))
    exe = str(tmp_path / "probe")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(tmp_path / "probe.c"), "-o", exe])
    got = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, *_ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)
    # and no field of the C structs is missing from the mirrors
    hdr = open(os.path.join(ROOT, "include", "spfe.h")).read()
    for cname, ct in structs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, re.S).group(1)
        n_decl = 0
        for decl in re.sub(r"/\*.*?\*/", "", body, flags=re.S).split(";"):
            if decl.strip():
                n_decl += len(decl.split(","))
        assert n_decl == len(ct._fields_), cname


def test_round2_entries_reject_null_arguments(lib):
    """The entries added in round 2 validate their arguments before touching CUDA (no GPU needed): NULL context / set."""
    ctx_null = C.c_void_p()
    assert lib.spfe_desc_set_size(None) == capi.ERR_INVALID
    assert lib.spfe_last_d2h_bytes(None, 0) == capi.ERR_INVALID
    assert lib.spfe_fetch_heat(None, 0, 0, None, None) == capi.ERR_INVALID
    out = C.c_void_p()
    assert lib.spfe_desc_set_create(None, 16, C.byref(out)) == capi.ERR_INVALID and not out.value
    assert lib.spfe_desc_set_upload(None, None, None, 0) == capi.ERR_INVALID
    assert lib.spfe_desc_set_from_frame(None, None, 0, 0, None, 0) == capi.ERR_INVALID
    assert lib.spfe_match_mutual_nn_sets(None, None, None, None, None) == capi.ERR_INVALID
    assert lib.spfe_match_knn2_sets(None, None, None, None, None) == capi.ERR_INVALID
    assert lib.spfe_search_guided_sets(None, None, None, None, None, None, None) == capi.ERR_INVALID
    lib.spfe_desc_set_destroy(None, None)                      # a no-op, not a crash
    assert capi.LAZY_HEAT == 16 and capi.DESC_F16 == 32 and capi.EXACT == 64
    hdr = open(os.path.join(ROOT, "include", "spfe.h")).read()
    for name, val in (("SPFE_LAZY_HEAT", 4), ("SPFE_DESC_F16", 5), ("SPFE_EXACT", 6)):
        assert re.search(rf"{name}\s*=\s*1u\s*<<\s*{val}\b", hdr), name
