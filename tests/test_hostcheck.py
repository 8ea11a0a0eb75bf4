"""Per-thread kernel logic (the AB_HD functions of aruco_b200/csrc/*.cuh, compiled for the host by
tests/hostcheck) against real OpenCV.  These are the exact functions the kernels execute per thread."""
import ctypes as C

import numpy as np
import pytest

from conftest import ROOT, intrinsics, needs_cv2

pytestmark = needs_cv2


def P(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def hc(built):
    lib = C.CDLL(ROOT + "/tests/_build/libhostcheck.so")
    lib.hc_approx_poly.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    lib.hc_solve_pnp.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_void_p, C.c_void_p]
    return lib


def walk_contours(hc, img, mn=-1, mx=1 << 30):
    H, W = img.shape
    lens = np.zeros(400000, np.int32)
    pts = np.zeros(2 * 3000000, np.int32)
    tot, nc = C.c_int(), C.c_int()
    n = hc.hc_find_contours(P(img), W, H, mn, mx, len(lens), len(pts) // 2, P(lens), P(pts), C.byref(tot), C.byref(nc))
    assert n >= 0
    out, off = [], 0
    for i in range(n):
        out.append(pts[2 * off:2 * (off + lens[i])].reshape(-1, 2).copy())
        off += lens[i]
    return out, tot.value, nc.value


def thresholded(frames):
    from oracle import cv2_oracle as o
    return {k: o.threshold(frames[k], o.ADPT_THRES, 7, 7) for k in ("single", "hrm", "board", "chessboard", "refine_fail")}


def test_border_walk_equals_findcontours(hc, frames):
    """Every contour, every point, OpenCV's order -- noise images, edge cases and the reference frames."""
    import cv2
    rng = np.random.default_rng(0)
    imgs = [((rng.random((200, 300)) < p) * 255).astype(np.uint8) for p in (0.2, 0.3, 0.5, 0.7, 0.9)]
    imgs += [np.zeros((10, 10), np.uint8), np.full((10, 10), 255, np.uint8), np.eye(9, dtype=np.uint8) * 255]
    one = np.zeros((5, 7), np.uint8)
    one[2, 3] = 255
    ring = np.zeros((9, 33), np.uint8)
    ring[2:7, 1:32] = 255
    ring[4, 3:30] = 0
    imgs += [one, ring, ring.T.copy()]
    imgs += list(thresholded(frames).values())
    for img in imgs:
        img = np.ascontiguousarray(img)
        mine, total, ncand = walk_contours(hc, img)
        ref, _ = cv2.findContours(img.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        assert total == len(ref) == len(mine)
        assert ncand >= total
        for a, b in zip(mine, ref):
            b = b.reshape(-1, 2)
            assert a.shape == b.shape and (a == b).all()


def test_border_walk_length_filter_4k(hc):
    import cv2
    from aruco_b200 import synth
    from oracle import cv2_oracle as o
    g, _ = synth.render_frame(3840, 2160, 100, 1, 2.0)
    th = o.threshold(g, o.ADPT_THRES, 7, 7)
    mine, _, _ = walk_contours(hc, th, 614, 7680)
    ref = [c.reshape(-1, 2) for c in cv2.findContours(th.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)[0] if 614 < len(c) < 7680]
    assert len(mine) == len(ref) > 100
    assert all(a.shape == b.shape and (a == b).all() for a, b in zip(mine, ref))


def test_polygon_fit_and_convexity_equal_cv2(hc, frames):
    import cv2
    rng = np.random.default_rng(1)
    imgs = list(thresholded(frames).values()) + [((rng.random((300, 300)) < 0.55) * 255).astype(np.uint8)]
    n_checked = n_quads = 0
    for im in imgs:
        for c in cv2.findContours(im.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)[0]:
            n = len(c)
            if n < 12:
                continue
            for f in (0.05, 0.02):
                ref = cv2.approxPolyDP(c, n * f, True).reshape(-1, 2)
                pts = np.ascontiguousarray(c.reshape(-1, 2).astype(np.int32))
                out = np.zeros((n, 2), np.int32)
                k = hc.hc_approx_poly(P(pts), n, n * f, P(out), n)
                assert k == len(ref) and (out[:k] == ref).all()
                n_checked += 1
                if len(ref) == 4:
                    n_quads += 1
                    q = np.ascontiguousarray(ref.astype(np.int32))
                    assert bool(hc.hc_is_convex4(P(q))) == bool(cv2.isContourConvex(ref.reshape(-1, 1, 2)))
    assert n_checked > 2000 and n_quads > 300
    for _ in range(5000):
        q = rng.integers(0, 40, (4, 2)).astype(np.int32)
        assert bool(hc.hc_is_convex4(P(q))) == bool(cv2.isContourConvex(q.reshape(-1, 1, 2)))


@pytest.mark.parametrize("S", [28, 48, 56, 80, 100])
def test_homography_and_nn_warp_equal_cv2(hc, frames, S):
    import cv2
    from oracle import cv2_oracle as o
    grey = np.ascontiguousarray(frames["board"])
    H, W = grey.shape
    cands, _ = o.detect_rectangles(o.threshold(grey, o.ADPT_THRES, 7, 7), 0.04, 0.5)
    assert len(cands) >= 24
    dst = np.array([[0, 0], [S - 1, 0], [S - 1, S - 1], [0, S - 1]], np.float32)
    for c in cands:
        q = np.ascontiguousarray(c["corners"])
        Mref = cv2.getPerspectiveTransform(q, dst)
        M = np.zeros(9)
        assert hc.hc_perspective(P(q), S, P(M)) == 1
        assert (M.reshape(3, 3) == Mref).all()  # bit-identical f64
        out = np.zeros((S, S), np.uint8)
        hc.hc_warp(P(grey), W, H, P(q), S, P(out))
        ref = cv2.warpPerspective(grey, Mref, (S, S), flags=cv2.INTER_NEAREST)
        assert (out == ref).all()
        t, _ = cv2.threshold(ref, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        assert hc.hc_otsu(P(np.ascontiguousarray(ref)), S * S) == int(t)


def test_fiducidal_decode_equals_oracle(hc, frames):
    from oracle import cv2_oracle as o
    for name in ("single", "board", "chessboard"):
        r = o.detect(frames[name], o.Params())
        for c in r["candidates"]:
            nrot = C.c_int()
            i = hc.hc_fid_decode(P(np.ascontiguousarray(c["canon"])), 56, C.byref(nrot))
            assert i == c["id"] and (i < 0 or nrot.value == c["nrot"])


def test_pose_equals_cv2_solvepnp(hc, expected):
    """CvLevMarq-schedule LM against cv2.solvePnP (ITERATIVE): goldens (with distortion) + synthetic frames,
    including near-frontal markers where the pose has two minima."""
    import cv2
    from aruco_b200 import synth
    from oracle import cv2_oracle as o
    worst = 0.0
    for name in ("single", "hrm"):
        K, D = intrinsics(expected, name)
        for m in expected["goldens"][name]["markers"]:
            c = np.array(m["corners"], np.float32)
            _, rv, tv = cv2.solvePnP(o.object_points(1.0), c.reshape(4, 1, 2), K, D.reshape(1, 5))
            r, t = np.zeros(3), np.zeros(3)
            assert hc.hc_solve_pnp(P(np.ascontiguousarray(K.ravel())), P(D), P(c), 1.0, P(r), P(t)) == 1
            worst = max(worst, np.abs(r - rv.ravel()).max() / np.abs(rv).max(), np.abs(t - tv.ravel()).max() / np.abs(tv).max())
    K, D = synth.camera_for(1920, 1080)
    n = 0
    for s in (5, 6):
        g, _ = synth.render_frame(1920, 1080, 50, seed=s, sigma=2.0)
        for m in o.detect(g, o.Params(), K, D, 0.05, keep=False)["markers"]:
            r, t = np.zeros(3), np.zeros(3)
            hc.hc_solve_pnp(P(np.ascontiguousarray(K.ravel())), P(D), P(np.ascontiguousarray(m["corners"])), 0.05, P(r), P(t))
            worst = max(worst, np.abs(r - m["rvec"]).max() / np.abs(m["rvec"]).max(), np.abs(t - m["tvec"]).max() / np.abs(m["tvec"]).max())
            n += 1
    assert n >= 90 and worst < 1e-4


def test_undistort_points_bit_identical(hc, expected):
    import cv2
    K, D = intrinsics(expected, "single")
    rng = np.random.default_rng(2)
    pts = (rng.random((2000, 2)) * [640, 480]).astype(np.float32)
    out = np.zeros_like(pts)
    hc.hc_undistort_px(P(np.ascontiguousarray(K.ravel())), P(D), P(pts), 2000, P(out))
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D.reshape(1, 5), None, K).reshape(-1, 2)
    assert (out == ref).all()


def test_cross_point_bit_identical_to_cv2_solve(hc):
    """getCrossPoint (markerdetector.cpp:132-139) = Matx22f::solve(DECOMP_SVD): the device function against cv2.solve."""
    import cv2
    rng = np.random.default_rng(3)
    for _ in range(500):
        k = rng.integers(0, 4)
        l1 = np.array([rng.uniform(-1, 1), -1.0, rng.uniform(-4000, 4000)], np.float32)
        l2 = np.array([-1.0, rng.uniform(-1, 1), rng.uniform(-4000, 4000)], np.float32)
        if k == 1:
            l1, l2 = l2, l1
        elif k == 2:
            l2 = np.array([rng.uniform(-1, 1), -1.0, rng.uniform(-4000, 4000)], np.float32)
        elif k == 3:
            l1 = np.array([-1.0, rng.uniform(-1, 1), rng.uniform(-4000, 4000)], np.float32)
        A = np.array([[l1[0], l1[1]], [l2[0], l2[1]]], np.float32)
        B = np.array([[-l1[2]], [-l2[2]]], np.float32)
        ref = cv2.solve(A, B, flags=cv2.DECOMP_SVD)[1].ravel()
        out = np.zeros(2, np.float32)
        hc.hc_cross_point(P(l1), P(l2), P(out))
        assert (out == ref).all(), (l1, l2, out, ref)


def test_rotate_x_axis(hc):
    from oracle import cv2_oracle as o
    rng = np.random.default_rng(4)
    for _ in range(50):
        r = rng.normal(0, 1.2, 3)
        ref = o.rotate_x_axis(r.copy())
        mine = r.copy()
        hc.hc_rotate_x_axis(P(mine))
        assert np.abs(mine - ref).max() < 1e-5  # f32 rotation matrices in the reference (utils.cpp:17)


def test_board_pose_equals_cv2(hc, expected):
    """N-point planar pose (solve_pnp_planar) vs cv2.solvePnP on the golden board corners, pixel and meter configs."""
    import cv2
    hc.hc_solve_pnp_planar.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    for name, cfgname, size in (("board", "board_pix", 1.0), ("chessboard", "chessboard_pix", 1.0), ("board", "board_meters", -1.0)):
        g, cfg = expected["goldens"][name], expected["boards"][cfgname]
        K, D = intrinsics(expected, name)
        pts = {m["id"]: np.array(m["corners"], np.float32) for m in cfg["markers"]}
        first = np.array(cfg["markers"][0]["corners"], np.float64)
        mpp = size / np.linalg.norm(first[0] - first[1]) if cfg["mInfoType"] == 0 else 1.0
        obj, img = [], []
        for m in g["markers"]:
            for p in range(4):
                img.append(m["corners"][p])
                obj.append((pts[m["id"]][p].astype(np.float64) * mpp).astype(np.float32))
        obj, img = np.ascontiguousarray(np.array(obj, np.float32)), np.ascontiguousarray(np.array(img, np.float32))
        r, t = np.zeros(3), np.zeros(3)
        assert hc.hc_solve_pnp_planar(P(np.ascontiguousarray(K.ravel())), P(D), P(obj), P(img), len(obj), P(r), P(t)) == 1
        _, rv, tv = cv2.solvePnP(obj, img.reshape(-1, 1, 2), K, D.reshape(1, 5))
        assert np.abs(r - rv.ravel()).max() / np.abs(rv).max() < 1e-6 and np.abs(t - tv.ravel()).max() / np.abs(tv).max() < 1e-6
        if cfg["mInfoType"] == 0:
            assert np.abs(r - np.array(g["rvec"])).max() < 1e-4 and np.abs(t - np.array(g["tvec"])).max() < 1e-4
