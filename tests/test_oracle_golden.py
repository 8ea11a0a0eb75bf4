"""Pins the oracles: both CPU oracles must reproduce the reference's own golden files
(testdata/{single,hrm,board,chessboard}/expected.yml -> tests/golden/expected.json, test/core_tests.cpp:77-358).

The goldens were written by the authors' OpenCV 3.x build in 2015; cross-version drift of the OpenCV
primitives is ~4e-4 px / 4e-5, far inside the north-star tolerances used here."""
import numpy as np
import pytest

from conftest import CORNER_TOL, intrinsics, needs_cv2
from oracle import native
from oracle.cv2_oracle import DEC_HRM, Params

GOLDEN_POSE_ATOL = 1e-4  # absolute, on Rvec/Tvec components of magnitude 0.05..8

CASES = [("single", Params(), True), ("hrm", Params(p1=21, p2=7, warp_size=48, min_size=0.005, decoder=DEC_HRM), True),
         ("board", Params(), False), ("chessboard", Params(), False)]


def _check(markers, golden, with_pose):
    assert [m["id"] for m in markers] == [g["id"] for g in golden]
    for m, g in zip(markers, golden):
        assert np.abs(m["corners"] - np.array(g["corners"], np.float32)).max() < CORNER_TOL
        if with_pose:
            assert np.abs(m["rvec"] - np.array(g["rvec"])).max() < GOLDEN_POSE_ATOL
            assert np.abs(m["tvec"] - np.array(g["tvec"])).max() < GOLDEN_POSE_ATOL


@pytest.mark.parametrize("name,P,cam", CASES)
def test_native_oracle_reproduces_reference_golden(built, frames, expected, name, P, cam):
    K, D = intrinsics(expected, name) if cam else (None, None)
    hrm = native.dict_from_yaml_text(expected["dictionaries"]["d4x4_100"]) if P.decoder == DEC_HRM else None
    r = native.detect(frames[name], P, K, D, 1.0 if cam else -1.0, hrm)
    _check(r["markers"], expected["goldens"][name]["markers"], cam)


@needs_cv2
@pytest.mark.parametrize("name,P,cam", CASES)
def test_cv2_oracle_reproduces_reference_golden(frames, expected, name, P, cam):
    from oracle import cv2_oracle as o
    K, D = intrinsics(expected, name) if cam else (None, None)
    hrm = o.HrmDictionary.from_yaml_text(expected["dictionaries"]["d4x4_100"]) if P.decoder == DEC_HRM else None
    r = o.detect(frames[name], P, K, D, 1.0 if cam else -1.0, hrm)
    _check(r["markers"], expected["goldens"][name]["markers"], cam)


def test_refine_fail_regression(built, frames, expected):
    """Aruco.RefineFail (test/core_tests.cpp:360-382): the 1-point-side frame must not crash."""
    hrm = native.dict_from_yaml_text(expected["dictionaries"]["d4x4_100"])
    r = native.detect(frames["refine_fail"], Params(p1=21, p2=7, warp_size=48, min_size=0.005, decoder=DEC_HRM), None, None, -1, hrm)
    assert len(r["markers"]) == 13


@needs_cv2
def test_bgr_golden_single(frames, expected):
    """The reference test loads the colour frame (cvtColor BGR2GRAY front step, markerdetector.cpp:307-310)."""
    from oracle import cv2_oracle as o
    K, D = intrinsics(expected, "single")
    r = o.detect(frames["single_bgr"], Params(), K, D, 1.0)
    _check(r["markers"], expected["goldens"]["single"]["markers"], True)


@needs_cv2
@pytest.mark.parametrize("name,cfgname", [("board", "board_pix"), ("chessboard", "chessboard_pix")])
def test_board_pose_golden_cv2_oracle(frames, expected, name, cfgname):
    """Aruco.Board / Aruco.Multi (test/core_tests.cpp:164-228): markers without camera, then the board pose."""
    from oracle import cv2_oracle as o
    K, D = intrinsics(expected, name)
    r = o.detect(frames[name], Params())
    b = o.board_detect(r["markers"], expected["boards"][cfgname], K, D, 1.0)
    g = expected["goldens"][name]
    assert len(b["markers"]) == len(g["markers"])
    assert np.abs(b["rvec"] - np.array(g["rvec"])).max() < GOLDEN_POSE_ATOL and np.abs(b["tvec"] - np.array(g["tvec"])).max() < GOLDEN_POSE_ATOL
