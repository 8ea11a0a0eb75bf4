"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): binarised image, candidate set and ids bit-exact; refined corners within
0.01 px; Rvec/Tvec within 1e-4 relative.  The primary checker is the dependency-free C++ oracle
(oracle/aruco_oracle.cpp); where cv2 is importable the cv2-driven oracle (real OpenCV) is checked as well."""
import numpy as np
import pytest

from conftest import CORNER_TOL, POSE_RTOL, have_cv2, intrinsics, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det(built):
    from aruco_b200 import MarkerDetector
    return MarkerDetector(0)


def configure(det, P, expected=None):
    from aruco_b200 import FiducidalMarkers, HighlyReliableMarkers
    det.setThresholdMethod(P.thres_method)
    det.setThresholdParams(P.p1, P.p2)
    det.setCornerRefinementMethod(P.corner_method)
    det.setMinMaxSize(P.min_size, P.max_size)
    det.setWarpSize(P.warp_size)
    det.enableErosion(P.erosion)
    det.setYPerpendicular(P.set_y_perpendicular)
    det.setThresholdParamRange(getattr(P, "p1_range", 0))
    det.setMakerDetectorFunction(HighlyReliableMarkers.detect if P.decoder == 1 else FiducidalMarkers.detect)


def oracle_pose(corners, K, D, size):
    """Pose of the oracle for GIVEN corners (strict pose-kernel parity, independent of corner differences)."""
    import ctypes as C
    from oracle import native
    lib = native.load()
    r, t = np.zeros(3), np.zeros(3)
    c = np.ascontiguousarray(corners, np.float32)
    Kf = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    Df = np.ascontiguousarray(np.asarray(D, np.float32).reshape(-1)[:5]) if D is not None else None
    lib.orc_solve_pnp(Kf.ctypes.data_as(C.c_void_p), Df.ctypes.data_as(C.c_void_p) if Df is not None else None,
                      c.ctypes.data_as(C.c_void_p), float(size), r.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    return r, t


def cv2_pose(corners, K, D, size):
    """cv2.solvePnP (real OpenCV) for GIVEN corners, as markerdetector.cpp:458 / marker.cpp:118 call it."""
    import cv2
    from oracle import cv2_oracle as o
    ok, rv, tv = cv2.solvePnP(o.object_points(size), np.asarray(corners, np.float32).reshape(4, 1, 2), np.asarray(K, np.float32),
                              np.asarray(D, np.float32).reshape(1, -1) if D is not None else None)
    return rv.ravel(), tv.ravel()


def cv2_pose_or_unconverged(m, K, D, size, stats=None):
    """GPU pose against OpenCV's solvePnP on the SAME corners.  They agree within 1e-4 except where OpenCV's own
    Levenberg-Marquardt was cut off at its 20-iteration limit in mid-descent (ill-conditioned quads: the returned pose
    then depends on the last bits of every accept/reject test).  Such a marker is accepted only if that is PROVEN:
    the two poses have the same reprojection error to 1e-3 relative, and restarting OpenCV from its own answer moves the
    pose by more than ten times the GPU/OpenCV difference.  Returns True if the strict 1e-4 gate held."""
    import cv2
    from oracle import cv2_oracle as o
    rr, tt = cv2_pose(m.corners, K, D, size)
    e = max(rel_err(m.Rvec, rr), rel_err(m.Tvec, tt))
    if e < POSE_RTOL:
        return True
    obj, Kd = o.object_points(size), np.asarray(K, np.float64)
    Dd = np.asarray(D, np.float64).reshape(1, -1) if D is not None else None

    def rms(rv, tv):
        pr = cv2.projectPoints(obj, np.asarray(rv, np.float64).reshape(3, 1), np.asarray(tv, np.float64).reshape(3, 1), Kd, Dd)[0]
        return float(np.sqrt(((pr.reshape(4, 2) - m.corners) ** 2).sum()))

    e_gpu, e_cv = rms(m.Rvec, m.Tvec), rms(rr, tt)
    ok, r2, t2 = cv2.solvePnP(obj, np.asarray(m.corners, np.float32).reshape(4, 1, 2), np.asarray(K, np.float32),
                              np.asarray(D, np.float32).reshape(1, -1) if D is not None else None,
                              rvec=rr.reshape(3, 1).copy(), tvec=tt.reshape(3, 1).copy(), useExtrinsicGuess=True)
    moved = max(rel_err(r2.ravel(), rr), rel_err(t2.ravel(), tt))
    assert abs(e_gpu - e_cv) <= 1e-3 * e_cv and moved > 10 * e, \
        "pose differs from cv2.solvePnP on identical corners: rel %.3g, rms %.6f vs %.6f, restart moves %.3g" % (e, e_gpu, e_cv, moved)
    if stats is not None:
        stats["unconverged_cv2"] = stats.get("unconverged_cv2", 0) + 1
    return False


def check_frame(det, grey, P, K=None, D=None, size=-1.0, hrm_text=None, frame=0, markers=None, min_direct_pose=0.99,
                stats=None):
    """Full per-stage comparison of one frame against the C++ oracle and the cv2 oracle (real OpenCV).
    Pose gate (north star: Rvec/Tvec within 1e-4 relative): the GPU pipeline's pose against each oracle PIPELINE's pose,
    for at least `min_direct_pose` of the frame's markers (`stats` accumulates the counts for multi-frame gates), and
    for EVERY marker against the oracle's and OpenCV's own solvePnP on the GPU's corners (so a marker that misses the
    direct gate is one whose pose is ill-conditioned in its corners, not one the pose kernel got wrong)."""
    from oracle import native
    hn = native.dict_from_yaml_text(hrm_text) if hrm_text else None
    ref = native.detect(grey, P, K, D, size, hn, cap=2048)
    if markers is None:
        markers = det.detect(grey, K, D, size, bool(P.set_y_perpendicular))
    assert (det.getThresholdedImage(frame) == ref["thres"]).all(), "binarised image not bit-exact"
    q, ids, nrot = det.getAllCandidates(frame)
    assert q.shape == ref["quads"].shape and (q == ref["quads"]).all(), "candidate set/order differs"
    assert list(ids) == list(ref["ids"]), "ids differ"
    assert all(a == b for a, b, i in zip(nrot, ref["nrot"], ids) if i >= 0)
    for i in range(len(ids)):
        assert (det.getCanonical(frame, i) == ref["canon"][i]).all(), "canonical image %d differs" % i
    assert [m.id for m in markers] == [m["id"] for m in ref["markers"]]
    posed = size > 0 and K is not None
    direct = 0
    for m, r in zip(markers, ref["markers"]):
        assert np.abs(m.corners - r["corners"]).max() < CORNER_TOL
        if posed:
            assert m.Rvec is not None and abs(m.ssize - size) < 1e-7
            if not P.set_y_perpendicular:  # (rotateXAxis is applied after the solve; covered by the direct check)
                rr, tt = oracle_pose(m.corners, K, D, size)
                assert rel_err(m.Rvec, rr) < POSE_RTOL and rel_err(m.Tvec, tt) < POSE_RTOL
            direct += rel_err(m.Rvec, r["rvec"]) < POSE_RTOL and rel_err(m.Tvec, r["tvec"]) < POSE_RTOL
        else:
            assert m.Rvec is None and m.ssize == -1.0
    if posed and markers:
        assert direct >= min_direct_pose * len(markers), "only %d/%d poses agree with the C++ oracle" % (direct, len(markers))
    if stats is not None:
        stats["markers"] = stats.get("markers", 0) + len(markers)
        stats["direct_port"] = stats.get("direct_port", 0) + direct
    if have_cv2():
        from oracle import cv2_oracle as o
        hc = o.HrmDictionary.from_yaml_text(hrm_text) if hrm_text else None
        b = o.detect(grey, P, K, D, size, hc)
        assert (det.getThresholdedImage(frame) == b["thres"]).all()
        oq = np.array([c["quad"] for c in b["candidates"]]).reshape(-1, 4, 2)
        assert q.shape == oq.shape and (q == oq).all()
        assert list(ids) == [c["id"] for c in b["candidates"]]
        for i, c in enumerate(b["candidates"]):
            assert (det.getContour(frame, i) == c["contour"]).all()
        assert [m.id for m in markers] == [m["id"] for m in b["markers"]]
        direct_cv = 0
        for m, r in zip(markers, b["markers"]):
            assert np.abs(m.corners - r["corners"]).max() < CORNER_TOL
            if posed:
                if not P.set_y_perpendicular:
                    rr, tt = cv2_pose(m.corners, K, D, size)
                    assert rel_err(m.Rvec, rr) < POSE_RTOL and rel_err(m.Tvec, tt) < POSE_RTOL
                direct_cv += rel_err(m.Rvec, r["rvec"]) < POSE_RTOL and rel_err(m.Tvec, r["tvec"]) < POSE_RTOL
        if posed and markers:
            assert direct_cv >= min_direct_pose * len(markers), "only %d/%d poses agree with the cv2 oracle" % (direct_cv, len(markers))
        if stats is not None:
            stats["direct_cv2"] = stats.get("direct_cv2", 0) + direct_cv
            stats["corner_max_cv2"] = max(stats.get("corner_max_cv2", 0.0),
                                          max([float(np.abs(m.corners - r["corners"]).max()) for m, r in zip(markers, b["markers"])] or [0.0]))
    return markers, ref


def P_(**kw):
    from oracle.cv2_oracle import Params
    return Params(**kw)


GOLDEN_CASES = [
    ("single", dict(), True), ("single", dict(corner_method=2), True), ("single", dict(corner_method=0), True),
    ("single", dict(erosion=True), True), ("single", dict(thres_method=0, p1=100), True), ("single", dict(p1=8, p2=6.5), True),
    ("single", dict(set_y_perpendicular=True), True),
    ("board", dict(), False), ("board", dict(), True), ("board", dict(corner_method=2), True), ("board", dict(erosion=True), True),
    ("board", dict(p1=21, p2=5), True), ("board", dict(warp_size=28, corner_method=0), False),
    ("chessboard", dict(), False), ("chessboard", dict(erosion=True, corner_method=2), True),
    ("chessboard", dict(thres_method=0, p1=90), True), ("chessboard", dict(min_size=0.01, max_size=0.9), True),
    # setThresholdParamRange: 3 / 5 threshold images per frame (utils/aruco_test.cpp:134 uses range 2)
    ("single", dict(p1_range=1), True), ("board", dict(p1_range=2), True), ("chessboard", dict(p1_range=2, corner_method=2), True),
    ("single", dict(p1_range=1, thres_method=0, p1=100), True), ("board", dict(p1_range=1, erosion=True), False),
    # ThresholdMethods::CANNY (cv::Canny(10, 220), markerdetector.cpp:664-675)
    ("single", dict(thres_method=2), True), ("board", dict(thres_method=2), True), ("chessboard", dict(thres_method=2, corner_method=2), True),
    # warp sizes whose S*S is not a multiple of 4 / of 7 (byte path of the canonical-image writer, ragged cells)
    ("single", dict(warp_size=57), True), ("board", dict(warp_size=35, corner_method=0), False), ("chessboard", dict(warp_size=99), True),
]


@pytest.mark.parametrize("name,kw,cam", GOLDEN_CASES)
def test_reference_frames_all_stages(det, frames, expected, name, kw, cam):
    P = P_(**kw)
    configure(det, P)
    K, D = intrinsics(expected, name) if cam else (None, None)
    check_frame(det, frames[name], P, K, D, 1.0 if cam else -1.0)


@pytest.mark.parametrize("name,cam", [("single", True), ("board", False), ("chessboard", False)])
def test_reference_goldens_through_cuda(det, frames, expected, name, cam):
    """The reference's own golden files (test/core_tests.cpp Aruco.Single / Board / Multi) reproduced by the GPU."""
    P = P_()
    configure(det, P)
    K, D = intrinsics(expected, name) if cam else (None, None)
    ms = det.detect(frames[name], K, D, 1.0 if cam else -1.0)
    gold = expected["goldens"][name]["markers"]
    assert [m.id for m in ms] == [g["id"] for g in gold]
    for m, g in zip(ms, gold):
        assert np.abs(m.corners - np.array(g["corners"], np.float32)).max() < CORNER_TOL
        if cam:
            assert np.abs(m.Rvec - np.array(g["rvec"])).max() < 1e-4 and np.abs(m.Tvec - np.array(g["tvec"])).max() < 1e-4


def test_hrm_golden_and_all_stages(det, frames, expected):
    """Aruco.HRM_Single (test/core_tests.cpp:310-358) + RefineFail (:360-382)."""
    from aruco_b200 import HighlyReliableMarkers
    text = expected["dictionaries"]["d4x4_100"]
    HighlyReliableMarkers.loadDictionary(text)
    P = P_(p1=21, p2=7, warp_size=48, min_size=0.005, decoder=1)
    configure(det, P)
    K, D = intrinsics(expected, "hrm")
    ms, _ = check_frame(det, frames["hrm"], P, K, D, 1.0, text)
    gold = expected["goldens"]["hrm"]["markers"]
    assert [m.id for m in ms] == [g["id"] for g in gold] == list(range(16))
    for m, g in zip(ms, gold):
        assert np.abs(m.corners - np.array(g["corners"], np.float32)).max() < CORNER_TOL
        assert np.abs(m.Rvec - np.array(g["rvec"])).max() < 1e-4 and np.abs(m.Tvec - np.array(g["tvec"])).max() < 1e-4
    ms, _ = check_frame(det, frames["refine_fail"], P, K, D, 1.0, text)
    assert len(ms) == 13


def test_bgr_front_step(det, frames, expected):
    """8UC3 input: cvtColor BGR2GRAY then the same path (markerdetector.cpp:307-310)."""
    configure(det, P_())
    K, D = intrinsics(expected, "single")
    ms = det.detect(frames["single_bgr"], K, D, 1.0)
    ref = det.detect(frames["single"], K, D, 1.0)  # frames['single'] = cv2 4.13 BGR2GRAY of the same PNG
    assert [m.id for m in ms] == [m.id for m in ref] and all((a.corners == b.corners).all() for a, b in zip(ms, ref))


@pytest.mark.parametrize("W,H,n,seed,sigma,kw", [
    (1920, 1080, 50, 11, 2.0, dict(corner_method=2)),   # C3: 1080p, 50 markers, SUBPIX + PnP
    (1920, 1080, 50, 12, 4.0, dict()),                  # stress: contour storm at sigma 4
    (3840, 2160, 100, 13, 2.0, dict()),                 # C4: 4K, 100 markers, LINES + PnP
    (3840, 2160, 100, 14, 2.0, dict(corner_method=2, erosion=False)),
    (1280, 720, 24, 15, 1.5, dict(erosion=True)),       # C2-shaped: 1280x720, erosion enabled
])
def test_synthetic_configs_all_stages(det, W, H, n, seed, sigma, kw):
    from aruco_b200 import synth
    g, truth = synth.render_frame(W, H, n, seed, sigma, marker_px=100 if W == 1280 else None)
    K, D = synth.camera_for(W, H)
    P = P_(**kw)
    configure(det, P)
    ms, _ = check_frame(det, g, P, K, D, 0.05)
    if not kw.get("erosion"):
        assert len(ms) >= 0.9 * n and set(m.id for m in ms) <= set(truth["ids"])


@pytest.mark.parametrize("W,H,n,frames_n,kw", [
    (1920, 1080, 50, 22, dict(corner_method=2)),  # C3: SUBPIX + PnP, >= 1000 markers
    (1920, 1080, 50, 22, dict()),                 # 1080p LINES (the frames with near-frontal, bistable markers)
    (3840, 2160, 100, 11, dict()),                # C4: LINES + PnP, >= 1000 markers
])
def test_pose_gate_c3_c4_against_both_oracles(det, W, H, n, frames_n, kw):
    """North-star pose gate at >= 1000 markers per config: Rvec/Tvec of the GPU pipeline within 1e-4 relative of the
    C++ oracle pipeline AND of the cv2 oracle pipeline (real OpenCV primitives; LINES fit = OpenCV's Jacobi SVD), for
    >= 99.9 % (C++ oracle) / >= 99.7 % (cv2 oracle) of the markers, every stage before it bit-exact (check_frame).  The
    few others are markers whose 4-point LM OpenCV itself leaves unconverged at its 20-iteration cap; each is proven to
    be one (cv2_pose_or_unconverged), at most 0.3 %."""
    from aruco_b200 import synth
    K, D = synth.camera_for(W, H)
    P = P_(**kw)
    configure(det, P)
    stats = {}
    for i in range(frames_n):
        g, _ = synth.render_frame(W, H, n, seed=700 + i, sigma=2.0)
        check_frame(det, g, P, K, D, 0.05, stats=stats)
    assert stats["markers"] >= 1000
    assert stats["direct_port"] >= 0.999 * stats["markers"], stats
    if have_cv2():
        assert stats["direct_cv2"] >= 0.997 * stats["markers"], stats
        assert stats.get("unconverged_cv2", 0) <= 0.003 * stats["markers"], stats
        if P.corner_method == 3:
            assert stats["corner_max_cv2"] < 1e-4, stats  # same f32 arithmetic as OpenCV's Jacobi: far below the 0.01 px bar


def test_pose_sensitivity_to_the_lapack_line_fit(det):
    """This wheel's cv2.solve switches to LAPACK sgesdd from 25 rows on (OpenBLAS HAL); its f32 rounding differs from
    OpenCV's own Jacobi SVD by ~1e-3 px in the LINES corners.  Against THAT oracle the corners still meet the 0.01 px bar
    and the poses agree directly for >= 98 %; every remaining marker is explained by its corners alone: OpenCV's own
    solvePnP on the GPU's corners returns the GPU's pose (checked for every marker inside check_frame)."""
    if not have_cv2():
        pytest.skip("cv2 needed")
    from aruco_b200 import synth
    from oracle import cv2_oracle as o
    W, H = 1920, 1080
    K, D = synth.camera_for(W, H)
    P = P_()
    configure(det, P)
    tot = agree = 0
    o.LINES_SOLVER = "cv2"
    try:
        for i in range(10):
            g, _ = synth.render_frame(W, H, 50, seed=i, sigma=2.0)
            ms = det.detect(g, K, D, 0.05)
            b = o.detect(g, P, K, D, 0.05)
            assert [m.id for m in ms] == [m["id"] for m in b["markers"]]
            for m, r in zip(ms, b["markers"]):
                assert np.abs(m.corners - r["corners"]).max() < CORNER_TOL
                cv2_pose_or_unconverged(m, K, D, 0.05)
                tot += 1
                agree += rel_err(m.Rvec, r["rvec"]) < POSE_RTOL and rel_err(m.Tvec, r["tvec"]) < POSE_RTOL
    finally:
        o.LINES_SOLVER = "jacobi"
    assert tot >= 450 and agree >= 0.98 * tot, (agree, tot)


def test_parked_walk_queue_overflow_finishes_in_place(built, monkeypatch):
    """When the queue of parked long walks is full, k_trace<false> finishes the walk itself and k_emit (two walkers per
    contour) writes it: same markers, same contours as with the default capacity (where k_trace<true> records the points
    of the walks it finishes and copies the winners' strips).  The same holds when the strip budget only pays for one CTA
    of k_trace<true>."""
    from aruco_b200 import MarkerDetector, synth
    g, _ = synth.render_frame(1920, 1080, 50, 21, 2.0)
    K, D = synth.camera_for(1920, 1080)
    ref_det = MarkerDetector()
    ref = ref_det.detect(g, K, D, 0.05)
    ref_contours = [ref_det.getContour(0, i) for i in range(len(ref_det.getAllCandidates(0)[0]))]
    monkeypatch.setenv("ARUCO_B200_CAP_LONG", "3")
    small = MarkerDetector()
    got = small.detect(g, K, D, 0.05)
    assert len(ref) >= 45 and [m.id for m in got] == [m.id for m in ref]
    assert all((a.corners == b.corners).all() and (a.Rvec == b.Rvec).all() for a, b in zip(got, ref))
    got_contours = [small.getContour(0, i) for i in range(len(small.getAllCandidates(0)[0]))]
    assert len(got_contours) == len(ref_contours) and all((a == b).all() for a, b in zip(got_contours, ref_contours))
    monkeypatch.delenv("ARUCO_B200_CAP_LONG")
    monkeypatch.setenv("ARUCO_B200_TRACE_REC_MB", "1")
    lean = MarkerDetector()
    got = lean.detect(g, K, D, 0.05)
    assert [m.id for m in got] == [m.id for m in ref] and all((a.corners == b.corners).all() for a, b in zip(got, ref))
    got_contours = [lean.getContour(0, i) for i in range(len(lean.getAllCandidates(0)[0]))]
    assert len(got_contours) == len(ref_contours) and all((a == b).all() for a, b in zip(got_contours, ref_contours))


@pytest.mark.parametrize("n,flips", [(4, 0), (5, 1), (6, 2), (8, 4)])
def test_hrm_synthetic_4k_with_bit_flips(det, expected, n, flips):
    """C5: HRM dictionaries on synthetic 4K frames; bit flips exercise the correction radius."""
    from aruco_b200 import HighlyReliableMarkers, synth
    text = expected["dictionaries"]["d%dx%d_100" % (n, n)]
    HighlyReliableMarkers.loadDictionary(text)
    codes = [l.split('"')[1] for l in text.splitlines() if l.startswith("marker_")]
    g, truth = synth.render_frame(3840, 2160, 100, seed=20 + n, sigma=2.0, hrm_codes=codes, hrm_n=n, flips=flips)
    K, D = synth.camera_for(3840, 2160)
    P = P_(p1=21, p2=7, warp_size=(n + 2) * 8, min_size=0.005, decoder=1)
    configure(det, P)
    # minSize 0.005 lets every isolated white cell through as a quad: > 512 candidates on some frames
    det.reserve(3840, 2160, 1, max_candidates=1024)
    ms, ref = check_frame(det, g, P, K, D, 0.05, text)
    assert len(ms) >= 80
    if flips:
        assert set(m.id for m in ms) <= set(truth["ids"])  # corrected back to the right dictionary entries


def test_batch_equals_single_and_is_order_stable(det):
    """A batch gives, frame by frame, exactly what single-frame calls give (byte-identical corners/poses)."""
    from aruco_b200 import synth
    configure(det, P_())
    K, D = synth.camera_for(1920, 1080)
    frames = np.stack([synth.render_frame(1920, 1080, 50, 30 + i, 2.0)[0] for i in range(5)])
    batch = det.detect_batch(frames, K, D, 0.05)
    for f in range(5):
        one = det.detect(frames[f], K, D, 0.05)
        assert [m.id for m in one] == [m.id for m in batch[f]]
        for a, b in zip(one, batch[f]):
            assert (a.corners == b.corners).all() and (a.Rvec == b.Rvec).all() and (a.Tvec == b.Tvec).all()


def test_full_size_batch_properties_4k(det):
    """BASELINE size: 64 x 4K frames in one launch (device-resident path). Size-independent properties:
    every frame of the batch equals its single-frame result; noise-free duplicates give identical output;
    ids are a subset of the rendered truth; chunked host path == device path."""
    import torch
    from aruco_b200 import synth
    configure(det, P_())
    K, D = synth.camera_for(3840, 2160)
    scenes = [synth.render_frame(3840, 2160, 100, 40 + i, 2.0) for i in range(4)]
    frames = np.stack([scenes[i % 4][0] for i in range(64)])
    dev = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    det.enqueue_device(dev.data_ptr(), 3840, 2160, 64, K, D, 0.05)
    res = det.fetch(64, 128)
    for f in range(64):
        assert [m.id for m in res[f]] == [m.id for m in res[f % 4]]
        assert all((a.corners == b.corners).all() and (a.Rvec == b.Rvec).all() for a, b in zip(res[f], res[f % 4]))
        assert set(m.id for m in res[f]) <= set(scenes[f % 4][1]["ids"]) and len(res[f]) >= 70
    host = det.detect_batch(frames, K, D, 0.05)  # chunked H2D path
    for f in range(64):
        assert [m.id for m in host[f]] == [m.id for m in res[f]]
        assert all((a.corners == b.corners).all() for a, b in zip(host[f], res[f]))
    # and the first scenes against the oracle
    from oracle import native
    for f in range(2):
        ref = native.detect(frames[f], P_(), K, D, 0.05, debug=False)["markers"]
        assert [m.id for m in res[f]] == [m["id"] for m in ref]
        assert all(np.abs(a.corners - b["corners"]).max() < CORNER_TOL for a, b in zip(res[f], ref))


def test_edge_cases(det):
    """Empty / saturated / tiny / ragged-width frames (no markers, no crash, thresholds still bit-exact)."""
    from oracle import native
    configure(det, P_())
    rng = np.random.default_rng(5)
    cases = [np.zeros((480, 640), np.uint8), np.full((480, 640), 255, np.uint8), rng.integers(0, 256, (64, 64), dtype=np.uint8),
             rng.integers(0, 256, (101, 333), dtype=np.uint8), rng.integers(0, 256, (37, 1021), dtype=np.uint8),
             (rng.random((240, 322)) < 0.5).astype(np.uint8) * 255]
    for img in cases:
        ms = det.detect(img)
        ref = native.detect(img, P_())
        assert (det.getThresholdedImage(0) == ref["thres"]).all()
        q, ids, _ = det.getAllCandidates(0)
        assert q.shape == ref["quads"].shape and (q == ref["quads"]).all() and list(ids) == list(ref["ids"])
        assert [m.id for m in ms] == [m["id"] for m in ref["markers"]]


def test_host_callback_decoder_matches_builtin(det, frames, expected):
    """setMakerDetectorFunction with a user function (markerdetector.h:78,243): called once per candidate in the
    reference's order with the S x S canonical image; results equal the device decoder's."""
    from oracle import native
    K, D = intrinsics(expected, "board")
    configure(det, P_())
    builtin = det.detect(frames["board"], K, D, 1.0)
    calls = []
    lib = native.load()

    def user_fn(canon):
        import ctypes as C
        assert canon.shape == (56, 56) and canon.dtype == np.uint8
        buf = np.ascontiguousarray(canon)
        calls.append(buf.copy())
        # decode with the oracle's Fiducidal decoder via a 1-candidate detect on the canonical image is overkill:
        # use the oracle's helper through the hostcheck-free path: threshold + cells in numpy
        from oracle import cv2_oracle as o
        if o.cv2 is None:
            pytest.skip("cv2 needed for the python decoder")
        mid, nrot, _, _ = o.fiducidal_detect(buf)
        return mid, nrot

    det.setMakerDetectorFunction(user_fn)
    custom = det.detect(frames["board"], K, D, 1.0)
    q, ids, _ = det.getAllCandidates(0)
    assert len(calls) == len(ids) == 30
    assert [m.id for m in custom] == [m.id for m in builtin]
    assert all((a.corners == b.corners).all() for a, b in zip(custom, builtin))
    for i, c in enumerate(calls):
        assert (c == det.getCanonical(0, i)).all()
    configure(det, P_())


def test_capacity_overflow_is_an_error_not_a_truncation(built, frames):
    from aruco_b200 import ArucoError, MarkerDetector
    d = MarkerDetector(0)
    d.reserve(640, 480, 1, max_quads=4, max_candidates=4)
    with pytest.raises(ArucoError) as e:
        d.detect(frames["board"])
    assert e.value.code == -3 and "overflow" in str(e.value)
    d.reserve(640, 480, 1, max_starts=100)
    with pytest.raises(ArucoError) as e:
        d.detect(frames["board"])
    assert e.value.code == -3
    d.reserve(640, 480, 1)
    assert len(d.detect(frames["board"])) == 24


@pytest.mark.parametrize("name,kw", [("single", dict(corner_method=1)), ("board", dict(corner_method=1)),
                                     ("chessboard", dict(corner_method=2, locked_corners=True)),
                                     ("chessboard", dict(corner_method=1, locked_corners=True)),
                                     ("single", dict(corner_method=2, locked_corners=True, p1=9))])
def test_harris_and_locked_corner_modes(det, frames, expected, name, kw):
    """HARRIS refinement (SubPixelCorner::RefineCorner, subpixelcorner.cpp:70-189) and enableLockedCornersMethod
    (findCornerMaxima, markerdetector.cpp:157-199) against the oracle."""
    P = P_(**kw)
    configure(det, P)
    det.enableLockedCornersMethod(bool(kw.get("locked_corners")))
    det.setCornerRefinementMethod(P.corner_method)
    K, D = intrinsics(expected, name)
    try:
        check_frame(det, frames[name], P, K, D, 1.0)
    finally:
        det.enableLockedCornersMethod(False)
        det.setCornerRefinementMethod(3)


def test_synthetic_1080p_locked_corners_and_harris(det):
    from aruco_b200 import synth
    g, _ = synth.render_frame(1920, 1080, 50, 21, 2.0)
    K, D = synth.camera_for(1920, 1080)
    for kw in (dict(corner_method=1), dict(corner_method=2, locked_corners=True)):
        P = P_(**kw)
        configure(det, P)
        det.enableLockedCornersMethod(bool(kw.get("locked_corners")))
        det.setCornerRefinementMethod(P.corner_method)
        try:
            check_frame(det, g, P, K, D, 0.05)
        finally:
            det.enableLockedCornersMethod(False)
            det.setCornerRefinementMethod(3)


def test_threshold_range_batch_4k(det):
    """Multi-threshold search on a batch: every frame equals its single-frame result and the oracle's ids."""
    from aruco_b200 import synth
    from oracle import native
    P = P_(p1_range=1)
    configure(det, P)
    K, D = synth.camera_for(1920, 1080)
    frames = np.stack([synth.render_frame(1920, 1080, 50, 60 + i, 2.0)[0] for i in range(3)])
    try:
        batch = det.detect_batch(frames, K, D, 0.05)
        for f in range(3):
            ref = native.detect(frames[f], P, K, D, 0.05, debug=False)["markers"]
            assert [m.id for m in batch[f]] == [m["id"] for m in ref]
            assert all(np.abs(a.corners - b["corners"]).max() < CORNER_TOL for a, b in zip(batch[f], ref))
            assert (det.getThresholdedImage(f) == native.detect(frames[f], P)["thres"]).all()
    finally:
        det.setThresholdParamRange(0)


def test_c4_full_batch_every_frame_against_oracle(det):
    """BASELINE.json configs[3] at full size: 256 distinct 3840x2160 frames (8 rendered scenes x per-frame noise,
    as bench.py builds them) in ONE device-resident batch; every frame's ids, corners and poses against the C++
    oracle run frame-parallel on the host cores."""
    import os
    import torch
    from aruco_b200 import synth
    from oracle import native
    W, H, B = 3840, 2160, 256
    configure(det, P_())
    K, D = synth.camera_for(W, H)
    scenes = [synth.render_frame(W, H, 100, seed=500 + i, as_float=True)[0] for i in range(8)]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(99)
    frames = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
    for i in range(B):
        clean = torch.from_numpy(scenes[i % 8]).cuda()
        frames[i] = torch.clamp(torch.round(clean + torch.randn((H, W), generator=gen, device="cuda") * 2.0), 0, 255).to(torch.uint8)
    torch.cuda.synchronize()
    det.enqueue_device(frames.data_ptr(), W, H, B, K, D, 0.05)
    res = det.fetch(B, 128)
    host = frames.cpu().numpy()
    del frames
    ref = native.detect_batch(host, P_(), K, D, 0.05, cap=128, threads=len(os.sched_getaffinity(0)))
    n_markers = direct = 0
    for f in range(B):
        assert [m.id for m in res[f]] == [m["id"] for m in ref[f]], "frame %d ids differ" % f
        for a, b in zip(res[f], ref[f]):
            assert np.abs(a.corners - b["corners"]).max() < CORNER_TOL
            n_markers += 1
            direct += rel_err(a.Rvec, b["rvec"]) < POSE_RTOL and rel_err(a.Tvec, b["tvec"]) < POSE_RTOL
    assert n_markers > 0.9 * 100 * B
    # GPU and oracle both run OpenCV's f32 Jacobi line fit (interpolate2Dline), so the poses agree directly
    assert direct >= 0.999 * n_markers, "%d/%d poses agree" % (direct, n_markers)
