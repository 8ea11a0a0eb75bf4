"""The C-ABI library loads, exports every symbol include/aruco_b200.h declares, and has no CPU path."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "aruco_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"AB_API\s+[\w\s\*]+?\b(ab_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("ab_create", "ab_destroy", "ab_set_params", "ab_detect_batch", "ab_enqueue_batch_device", "ab_fetch_results",
                 "ab_load_hrm_dictionary", "ab_set_decoder_callback", "ab_threshold", "ab_detect_rectangles", "ab_warp",
                 "ab_get_thresholded", "ab_get_candidates", "ab_calculate_extrinsics"):
        assert must in syms
    assert len(syms) >= 28


def test_library_exports_every_declared_symbol(built):
    from aruco_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export " + name
        assert name in _lib.SYMBOLS, "binding table lacks " + name
    assert b"sm_100a" in lib.ab_version()


def test_struct_layouts(built):
    from aruco_b200 import _lib
    assert C.sizeof(_lib.ab_marker) == 96
    p = _lib.ab_params()
    assert _lib.load().ab_default_params(C.byref(p)) == 0
    # ctor defaults of the reference, src/markerdetector.cpp:235-249
    assert (p.thres_method, p.thres_param1, p.thres_param2, p.corner_method) == (1, 7.0, 7.0, 3)
    assert abs(p.min_size - 0.04) < 1e-7 and p.max_size == 0.5 and p.warp_size == 56 and abs(p.border_dist - 0.025) < 1e-7
    assert (p.locked_corners, p.erosion, p.decoder) == (0, 0, 0)


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from aruco_b200 import ArucoError, MarkerDetector, _lib
    h = C.c_void_p()
    assert _lib.load().ab_create(0, C.byref(h)) == _lib.AB_E_NO_DEVICE
    with pytest.raises(ArucoError):
        MarkerDetector()


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under aruco_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "aruco_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn), errors="ignore") as f:
                    txt = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), fn
                assert "liboracle" not in txt and "oracle/" not in txt, fn


def test_no_source_names_the_banned_batch_copy_calls():
    banned = ["cudaMemcpy" + "BatchAsync", "cudaMemcpy3D" + "BatchAsync", "cuMemcpy" + "BatchAsync", "cuMemcpy3D" + "BatchAsync"]
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "__pycache__")]
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")) or fn == "Makefile":
                with open(os.path.join(dirpath, fn), errors="ignore") as f:
                    txt = f.read()
                for b in banned:
                    assert b not in txt, (fn, b)
