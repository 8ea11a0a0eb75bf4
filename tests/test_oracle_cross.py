"""The dependency-free C++ oracle against real OpenCV (cv2 4.13): primitive by primitive and end to end."""
import ctypes as C

import numpy as np
import pytest

from conftest import CORNER_TOL, POSE_RTOL, intrinsics, needs_cv2, rel_err
from oracle import native
from oracle.cv2_oracle import DEC_HRM, FIXED_THRES, SUBPIX, NONE, Params

pytestmark = needs_cv2


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("k,c", [(3, 7), (7, 7), (7, 6.5), (21, 7), (35, 3), (8, 7)])
def test_adaptive_threshold_matches_cv2(built, frames, k, c):
    import cv2
    from oracle import cv2_oracle as o
    lib = native.load()
    rng = np.random.default_rng(k)
    for img in (frames["single"], rng.integers(0, 256, (97, 131), dtype=np.uint8)):
        img = np.ascontiguousarray(img)
        out = np.empty_like(img)
        lib.orc_threshold(P(img), img.shape[1], img.shape[0], 1, float(k), float(c), P(out))
        assert (out == o.threshold(img, o.ADPT_THRES, k, c)).all()


def test_find_contours_matches_cv2(built, frames):
    import cv2
    from oracle import cv2_oracle as o
    lib = native.load()
    rng = np.random.default_rng(3)
    imgs = [((rng.random((120, 160)) < p) * 255).astype(np.uint8) for p in (0.3, 0.5, 0.7)]
    imgs += [o.threshold(frames[k], o.ADPT_THRES, 7, 7) for k in ("single", "hrm")]
    for img in imgs:
        img = np.ascontiguousarray(img)
        lens = np.zeros(100000, np.int32)
        pts = np.zeros((1000000, 2), np.int32)
        n = lib.orc_find_contours(P(img), img.shape[1], img.shape[0], len(lens), len(pts), P(lens), P(pts))
        ref, _ = cv2.findContours(img.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        assert n == len(ref)
        off = 0
        for i, c in enumerate(ref):
            assert lens[i] == len(c) and (pts[off:off + lens[i]] == c.reshape(-1, 2)).all()
            off += lens[i]


def test_approx_poly_matches_cv2(built, frames):
    import cv2
    from oracle import cv2_oracle as o
    lib = native.load()
    th = o.threshold(frames["board"], o.ADPT_THRES, 7, 7)
    cs, _ = cv2.findContours(th.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    n_checked = 0
    for c in cs:
        n = len(c)
        if n < 12:
            continue
        for f in (0.05, 0.02):
            ref = cv2.approxPolyDP(c, n * f, True).reshape(-1, 2)
            pts = np.ascontiguousarray(c.reshape(-1, 2).astype(np.int32))
            out = np.zeros((n, 2), np.int32)
            k = lib.orc_approx_poly(P(pts), n, n * f, P(out), n)
            assert k == len(ref) and (out[:k] == ref).all()
            n_checked += 1
    assert n_checked > 300


CASES = [("single", Params(), True), ("single", Params(corner_method=SUBPIX), True), ("single", Params(corner_method=NONE), True),
         ("board", Params(erosion=True), False), ("chessboard", Params(thres_method=FIXED_THRES, p1=100), True),
         ("hrm", Params(p1=21, p2=7, warp_size=48, min_size=0.005, decoder=DEC_HRM), True),
         ("single", Params(p1_range=1), True), ("board", Params(p1_range=2), False),
         ("single", Params(thres_method=2), True), ("chessboard", Params(thres_method=2), False)]


@pytest.mark.parametrize("name,prm,cam", CASES)
def test_native_oracle_equals_cv2_oracle(built, frames, expected, name, prm, cam):
    from oracle import cv2_oracle as o
    K, D = intrinsics(expected, name) if cam else (None, None)
    text = expected["dictionaries"]["d4x4_100"]
    hn = native.dict_from_yaml_text(text) if prm.decoder == DEC_HRM else None
    hc = o.HrmDictionary.from_yaml_text(text) if prm.decoder == DEC_HRM else None
    a = native.detect(frames[name], prm, K, D, 1.0 if cam else -1.0, hn)
    b = o.detect(frames[name], prm, K, D, 1.0 if cam else -1.0, hc)
    assert (a["thres"] == b["thres"]).all()
    assert a["n_contours"] == b["n_contours"]
    oq = np.array([c["quad"] for c in b["candidates"]]).reshape(-1, 4, 2)
    assert a["quads"].shape == oq.shape and (a["quads"] == oq).all()
    assert list(a["ids"]) == [c["id"] for c in b["candidates"]]
    for i, c in enumerate(b["candidates"]):
        assert (a["canon"][i] == c["canon"]).all()
        if c["id"] >= 0:
            assert a["nrot"][i] == c["nrot"]
    assert [m["id"] for m in a["markers"]] == [m["id"] for m in b["markers"]]
    for x, y in zip(a["markers"], b["markers"]):
        assert np.abs(x["corners"] - y["corners"]).max() < CORNER_TOL
        if cam:
            assert rel_err(x["rvec"], y["rvec"]) < POSE_RTOL and rel_err(x["tvec"], y["tvec"]) < POSE_RTOL


def test_native_oracle_equals_cv2_oracle_synthetic_1080p(built):
    """End to end on a C3-shaped frame.  LINES corners differ by ~1e-3 px between OpenCV's f32 SVD line fit and
    an f64 fit; planar PnP of a near-frontal marker has two minima, and such a corner change can flip which one
    cv2's 20-iteration LM reaches.  So the pose is checked strictly on IDENTICAL corners (cv2.solvePnP on the
    native oracle's corners), and directly for the well-conditioned majority."""
    import cv2
    from oracle import cv2_oracle as o
    from aruco_b200 import synth
    g, truth = synth.render_frame(1920, 1080, 50, seed=5, sigma=2.0)
    K, D = synth.camera_for(1920, 1080)
    a = native.detect(g, Params(), K, D, 0.05)
    b = o.detect(g, Params(), K, D, 0.05)
    assert (a["thres"] == b["thres"]).all()
    assert [m["id"] for m in a["markers"]] == [m["id"] for m in b["markers"]]
    assert set(m["id"] for m in a["markers"]) <= set(truth["ids"]) and len(a["markers"]) >= 45
    direct = 0
    for x, y in zip(a["markers"], b["markers"]):
        assert np.abs(x["corners"] - y["corners"]).max() < CORNER_TOL
        ok, rv, tv = cv2.solvePnP(o.object_points(0.05), x["corners"].reshape(4, 1, 2), K, D.reshape(1, 5))
        assert rel_err(x["rvec"], rv.ravel()) < POSE_RTOL and rel_err(x["tvec"], tv.ravel()) < POSE_RTOL
        direct += rel_err(x["rvec"], y["rvec"]) < POSE_RTOL and rel_err(x["tvec"], y["tvec"]) < POSE_RTOL
    assert direct >= 0.85 * len(a["markers"])


@pytest.mark.parametrize("name,kw", [("single", dict(corner_method=1)), ("board", dict(corner_method=1)),
                                     ("chessboard", dict(corner_method=2, locked_corners=True)),
                                     ("chessboard", dict(corner_method=1, locked_corners=True))])
def test_harris_and_locked_corners_native_equals_cv2(built, frames, name, kw):
    from oracle import cv2_oracle as o
    prm = Params(**kw)
    a = native.detect(frames[name], prm)
    b = o.detect(frames[name], prm)
    assert [m["id"] for m in a["markers"]] == [m["id"] for m in b["markers"]] and len(a["markers"]) >= 6
    for x, y in zip(a["markers"], b["markers"]):
        assert np.abs(x["corners"] - y["corners"]).max() < CORNER_TOL
