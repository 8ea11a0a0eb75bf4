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


def test_jacobi_svd_restatement_is_cv2_solve(built):
    """interpolate2Dline / getCrossPoint call cv::solve(DECOMP_SVD) on CV_32F systems (markerdetector.cpp:112,124,138).
    Below 25 rows cv2.solve runs OpenCV's own Jacobi SVD: both restatements (numpy in cv2_oracle, C++ in the port) are
    BIT-identical to it for every row count 2..24; from 25 rows on (where this wheel hands over to LAPACK) the two
    restatements stay bit-identical to each other."""
    import ctypes as C
    import cv2
    from oracle import cv2_oracle as o
    lib = native.load()
    rng = np.random.default_rng(0)
    for N in list(range(2, 25)) * 6 + [25, 60, 333, 1900] * 4:
        x0, y0, sl = rng.integers(0, 3800), rng.integers(0, 2100), rng.uniform(-0.95, 0.95)
        xs = (x0 + np.arange(N)).astype(np.float32)
        ys = np.round(y0 + sl * np.arange(N) + rng.normal(0, 0.4, N)).astype(np.float32)
        if rng.random() < 0.3:
            xs = xs + rng.normal(0, 0.3, N).astype(np.float32)  # undistorted (non-integer) coordinates
        A = np.ascontiguousarray(np.stack([xs, np.ones(N, np.float32)], 1))
        B = np.ascontiguousarray(ys)
        py = o.jacobi_svd_solve_f32(A, B.reshape(-1, 1))
        cc = np.zeros(2, np.float32)
        lib.orc_svd_solve_f32(A.ctypes.data_as(C.c_void_p), N, B.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p))
        assert (py == cc).all(), N
        if N < 25:
            ref = cv2.solve(A, B.reshape(-1, 1), flags=cv2.DECOMP_SVD)[1].ravel()
            assert (ref == py).all(), (N, ref, py)
    for _ in range(200):  # the 2x2 systems of getCrossPoint
        A = np.array([[rng.uniform(-1, 1), -1.0], [-1.0, rng.uniform(-1, 1)]], np.float32)
        if rng.random() < 0.5:
            A = A[::-1].copy()
        B = rng.uniform(-4000, 4000, 2).astype(np.float32)
        ref = cv2.solve(A, B.reshape(2, 1), flags=cv2.DECOMP_SVD)[1].ravel()
        cc = np.zeros(2, np.float32)
        lib.orc_svd_solve_f32(A.ctypes.data_as(C.c_void_p), 2, B.ctypes.data_as(C.c_void_p), cc.ctypes.data_as(C.c_void_p))
        assert (ref == cc).all()


def test_native_oracle_equals_cv2_oracle_synthetic_1080p(built):
    """End to end on C3-shaped frames (seed 5 holds near-frontal markers whose planar PnP is bistable).  Both oracles fit
    the LINES sides with OpenCV's f32 Jacobi SVD, so corners are bit-identical and EVERY pose agrees within 1e-4."""
    import cv2
    from oracle import cv2_oracle as o
    from aruco_b200 import synth
    for seed in (5, 8):
        g, truth = synth.render_frame(1920, 1080, 50, seed=seed, sigma=2.0)
        K, D = synth.camera_for(1920, 1080)
        a = native.detect(g, Params(), K, D, 0.05)
        b = o.detect(g, Params(), K, D, 0.05)
        assert (a["thres"] == b["thres"]).all()
        assert [m["id"] for m in a["markers"]] == [m["id"] for m in b["markers"]]
        assert set(m["id"] for m in a["markers"]) <= set(truth["ids"]) and len(a["markers"]) >= 45
        for x, y in zip(a["markers"], b["markers"]):
            assert (x["corners"] == y["corners"]).all()
            ok, rv, tv = cv2.solvePnP(o.object_points(0.05), x["corners"].reshape(4, 1, 2), K, D.reshape(1, 5))
            assert rel_err(x["rvec"], rv.ravel()) < POSE_RTOL and rel_err(x["tvec"], tv.ravel()) < POSE_RTOL
            assert rel_err(x["rvec"], y["rvec"]) < POSE_RTOL and rel_err(x["tvec"], y["tvec"]) < POSE_RTOL


def test_lapack_line_fit_is_within_the_corner_bar(built):
    """cv2.solve of this wheel (LAPACK sgesdd from 25 rows on) against OpenCV's Jacobi: LINES corners differ by ~1e-3 px
    (f32 rounding of an ill-scaled system), far inside the 0.01 px bar; the pose of a few near-frontal markers is
    sensitive to that (two close minima of planar PnP) -- >= 97 % still agree within 1e-4."""
    from oracle import cv2_oracle as o
    from aruco_b200 import synth
    K, D = synth.camera_for(1920, 1080)
    tot = agree = 0
    worst = 0.0
    for seed in (0, 1, 5, 8):
        g, _ = synth.render_frame(1920, 1080, 50, seed=seed, sigma=2.0)
        a = o.detect(g, Params(), K, D, 0.05)
        o.LINES_SOLVER = "cv2"
        try:
            b = o.detect(g, Params(), K, D, 0.05)
        finally:
            o.LINES_SOLVER = "jacobi"
        assert [m["id"] for m in a["markers"]] == [m["id"] for m in b["markers"]]
        for x, y in zip(a["markers"], b["markers"]):
            worst = max(worst, float(np.abs(x["corners"] - y["corners"]).max()))
            tot += 1
            agree += rel_err(x["rvec"], y["rvec"]) < POSE_RTOL and rel_err(x["tvec"], y["tvec"]) < POSE_RTOL
    assert worst < CORNER_TOL and agree >= 0.97 * tot, (worst, agree, tot)


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
