"""-m gpu: the public workers and the error behaviour of the MarkerDetector mirror (markerdetector.h:129-280)."""
import numpy as np
import pytest

from conftest import POSE_RTOL, intrinsics, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def det(built):
    from aruco_b200 import MarkerDetector
    return MarkerDetector(0)


def test_defaults_and_setters_mirror_the_reference(built):
    from aruco_b200 import ArucoError, MarkerDetector
    d = MarkerDetector(0)
    assert d.getThresholdMethod() == MarkerDetector.ADPT_THRES and d.getThresholdParams() == (7.0, 7.0)
    assert d.getCornerRefinementMethod() == MarkerDetector.LINES and d.getWarpSize() == 56
    mn, mx = d.getMinMaxSize()
    assert abs(mn - 0.04) < 1e-7 and mx == 0.5
    d.setDesiredSpeed(0)
    assert (d.getWarpSize(), d.getCornerRefinementMethod()) == (56, MarkerDetector.SUBPIX)
    d.setDesiredSpeed(2)
    assert (d.getWarpSize(), d.getCornerRefinementMethod()) == (28, MarkerDetector.NONE)
    d.setDesiredSpeed(3)  # passes the clamp and changes nothing
    assert (d.getWarpSize(), d.getDesiredSpeed()) == (28, 3)
    d.setDesiredSpeed(7)
    assert d.getDesiredSpeed() == 2
    d.enableLockedCornersMethod(False)
    for bad in ((0, 0.5), (0.5, 0.4), (0.1, 1.5)):  # CV_Assert in setMinMaxSize (cpp:1031-1038)
        with pytest.raises(ArucoError):
            d.setMinMaxSize(*bad)
    with pytest.raises(ArucoError):  # CV_Assert(val >= 10) in setWarpSize (cpp:1047-1051)
        d.setWarpSize(9)
    assert d.getWarpSize() == 28
    with pytest.raises(ArucoError):  # CV_Assert(grey.type()==CV_8UC1) (cpp:644)
        d.thresHold(1, np.zeros((10, 10), np.float32))
    with pytest.raises(ArucoError):  # CV_Assert(points.size()==4) (cpp:685)
        d.warp(np.zeros((20, 20), np.uint8), 56, [[0, 0], [1, 0], [1, 1]])


@pytest.mark.parametrize("method,p1,p2", [(1, 7, 7), (1, -1, -1), (1, 2, 7), (1, 8, 6.5), (1, 21, 7), (1, 35, 3), (1, 61, 0), (0, 100, 0), (0, 127.5, 0), (2, 0, 0),
                                          # lane-paired kernel (K <= 11 while K^2*255 + |K^2*delta| < 2^15) and its fallbacks
                                          (1, 3, 7), (1, 5, 0), (1, 9, 7), (1, 11, 7), (1, 11, 20), (1, 7, -7), (1, 7, 300), (1, 7, -300), (1, 13, 7)])
def test_threshold_worker_bit_exact(det, frames, method, p1, p2):
    from oracle import native
    import ctypes as C
    lib = native.load()
    rng = np.random.default_rng(1)
    for img in (frames["hrm"], rng.integers(0, 256, (97, 131), dtype=np.uint8), rng.integers(0, 256, (480, 1000), dtype=np.uint8),
                rng.integers(0, 256, (131, 1536), dtype=np.uint8), rng.integers(100, 140, (300, 16), dtype=np.uint8)):
        img = np.ascontiguousarray(img)
        got = det.thresHold(method, img, p1, p2)
        q1, q2 = (7.0 if p1 == -1 else p1), (7.0 if p2 == -1 else p2)
        ref = np.empty_like(img)
        lib.orc_threshold(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], method, float(q1), float(q2), ref.ctypes.data_as(C.c_void_p))
        assert (got == ref).all()


@pytest.mark.parametrize("W", [320, 512, 640, 768, 960, 1024, 1280, 1536, 1920, 3840])
def test_tma_threshold_tiles_and_borders_bit_exact(det, W):
    """The TMA-staged kernel over every tile width it dispatches to (8*TO in 768 / 960 / 640 / 512 / 320: one tile, several
    tiles, box wider than the image), block sizes 3..21, heights that are not a multiple of the row group (a single partial
    group, 1 row), noise and hard edges at all four image borders -- binarised image AND the packed copy (through erosion,
    which reads the packed image and writes the u8 image) bit-exact against the oracle."""
    import ctypes as C
    from oracle import native
    lib = native.load()
    rng = np.random.default_rng(W)
    for H, k, delta in ((1, 7, 7), (13, 3, 2), (126, 5, 7), (127, 7, 7), (300, 9, -3), (253, 11, 7),
                        # the wide kernel (K >= 13: staged ring instead of the register ring, 32-bit window sums)
                        (1, 13, 7), (40, 15, 0), (127, 17, -2), (126, 19, 7), (260, 21, 7), (33, 21, 300)):
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        img[:, :2] = 255 * (H & 1)            # hard edges on the left / right / top / bottom borders
        img[:, -3:] = 0
        img[0, :] = 200
        img[-1, W // 3:] = 17
        got = det.thresHold(1, img, k, delta)
        ref = np.empty_like(img)
        lib.orc_threshold(img.ctypes.data_as(C.c_void_p), W, H, 1, float(k), float(delta), ref.ctypes.data_as(C.c_void_p))
        assert (got == ref).all(), (W, H, k)
    # the packed copy and the erosion that follows it (threshold kernel without its u8 store + k_erode)
    from oracle.cv2_oracle import Params
    img = rng.integers(0, 256, (139, W), dtype=np.uint8)
    det.enableErosion(True)
    try:
        det.detect(img)
        ref = native.detect(img, Params(erosion=True))
        assert (det.getThresholdedImage(0) == ref["thres"]).all()
    finally:
        det.enableErosion(False)


def test_detect_rectangles_and_warp_workers(det, frames):
    from oracle import native
    from oracle.cv2_oracle import Params
    ref = native.detect(frames["board"], Params())
    quads = det.detectRectangles(ref["thres"])
    assert quads.shape == ref["quads"].shape and (quads == ref["quads"]).all()
    for i in range(len(quads)):
        assert (det.warp(frames["board"], 56, quads[i]) == ref["canon"][i]).all()


def test_calculate_extrinsics_worker(det, expected):
    """Marker::calculateExtrinsics (marker.cpp:112-125) on the golden corners vs the golden poses."""
    from aruco_b200 import ArucoError, Marker
    K, D = intrinsics(expected, "single")
    gold = expected["goldens"]["single"]["markers"]
    ms = det.calculateExtrinsics([Marker(g["id"], g["corners"]) for g in gold], 1.0, K, D)
    for m, g in zip(ms, gold):
        assert rel_err(m.Rvec, g["rvec"]) < POSE_RTOL and rel_err(m.Tvec, g["tvec"]) < POSE_RTOL and m.ssize == 1.0
    with pytest.raises(ArucoError):  # CV_Assert(markerSizeMeters > 0) (marker.cpp:114)
        det.calculateExtrinsics(ms, -1.0, K, D)


def test_hrm_requires_dictionary(built):
    from aruco_b200 import ArucoError, HighlyReliableMarkers, MarkerDetector
    saved = HighlyReliableMarkers._dict
    HighlyReliableMarkers._dict = None
    try:
        with pytest.raises(ArucoError):
            MarkerDetector(0).setMakerDetectorFunction(HighlyReliableMarkers.detect)
    finally:
        HighlyReliableMarkers._dict = saved


def test_cpp_facade_aruco_simple(built, frames, expected, tmp_path):
    """The header-only C++ facade (include/aruco/markerdetector.hpp): a headless utils/aruco_simple.cpp on the
    reference's testdata/single frame must print the golden markers."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "tests", "_build", "aruco_simple")
    assert os.path.exists(exe)
    raw = tmp_path / "single.raw"
    frames["single"].tofile(str(raw))
    K, D = intrinsics(expected, "single")
    args = [exe, str(raw), "640", "480", repr(float(K[0, 0])), repr(float(K[1, 1])), repr(float(K[0, 2])), repr(float(K[1, 2]))]
    args += [repr(float(d)) for d in D] + ["1.0"]
    out = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    gold = expected["goldens"]["single"]["markers"]
    assert [int(r[0]) for r in rows] == [g["id"] for g in gold]
    for r, g in zip(rows, gold):
        c = np.array([float(v) for v in r[1:9]]).reshape(4, 2)
        assert np.abs(c - np.array(g["corners"])).max() < 0.01 and int(r[9]) == 1
        assert np.abs(np.array([float(v) for v in r[10:13]]) - np.array(g["rvec"])).max() < 1e-4
        assert np.abs(np.array([float(v) for v in r[13:16]]) - np.array(g["tvec"])).max() < 1e-4


@pytest.mark.parametrize("name,cfgfile", [("board", "board__board_pix.yml"), ("chessboard", "chessboard__chessboardinfo_pix.yml")])
def test_cpp_facade_aruco_simple_board_from_yaml(built, frames, expected, name, cfgfile, tmp_path):
    """Headless utils/aruco_simple_board.cpp on the C++ facade, configured only by the reference's YAML files
    (board configuration + intrinsics read by include/aruco/serialization.hpp); the Board it saves, read back by
    cv::FileStorage, must be the reference's golden board (Aruco.Board / Aruco.Multi)."""
    import os
    import subprocess
    import cv2
    from conftest import ROOT
    exe = os.path.join(ROOT, "tests", "_build", "aruco_simple_board")
    ydir = os.path.join(ROOT, "tests", "golden", "yaml")
    raw, out_yml = tmp_path / "frame.raw", tmp_path / "board.yml"
    frames[name].tofile(str(raw))
    H, W = frames[name].shape
    r = subprocess.run([exe, str(raw), str(W), str(H), os.path.join(ydir, cfgfile), os.path.join(ydir, name + "__intrinsics.yml"), "1.0",
                        str(out_yml)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    fs = cv2.FileStorage(str(out_yml), cv2.FILE_STORAGE_READ)
    b = fs.getNode("Board")
    g = expected["goldens"][name]
    assert np.abs(np.array([b.getNode("Rvec").at(k).real() for k in range(3)]) - np.array(g["rvec"])).max() < 1e-4
    assert np.abs(np.array([b.getNode("Tvec").at(k).real() for k in range(3)]) - np.array(g["tvec"])).max() < 1e-4
    ms = b.getNode("Markers")
    assert [int(ms.at(i).getNode("id").real()) for i in range(ms.size())] == [m["id"] for m in g["markers"]]
    for i, m in enumerate(g["markers"]):
        c = ms.at(i).getNode("corners")
        got = np.array([[c.at(k).at(0).real(), c.at(k).at(1).real()] for k in range(4)])
        assert np.abs(got - np.array(m["corners"])).max() < 0.01


def test_cpp_facade_aruco_create_board(built, expected, tmp_path):
    """Headless utils/aruco_create_board.cpp on the C++ facade: the image equals the reference's printed board
    (testdata/board/board.png), the YAML it writes is read back by cv::FileStorage with board_pix.yml's layout scaled to
    150-pixel markers, and the board is detected in its own picture."""
    import os
    import subprocess
    import cv2
    from conftest import ROOT
    exe = os.path.join(ROOT, "tests", "_build", "aruco_create_board")
    ids = [m["id"] for m in expected["boards"]["board_pix"]["markers"]]
    raw, yml = tmp_path / "board.raw", tmp_path / "board.yml"
    r = subprocess.run([exe, "4", "6", "150", "0", "30", str(raw), str(yml)] + [str(i) for i in ids], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    w, h, n, found, prob = r.stdout.split()
    assert (int(w), int(h), int(n), int(found), float(prob)) == (690, 1050, 24, 24, 1.0)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "render.npz"))["board_4x6_150_30"]
    assert np.array_equal(np.fromfile(str(raw), np.uint8).reshape(1050, 690), gold)
    fs = cv2.FileStorage(str(yml), cv2.FILE_STORAGE_READ)
    ms = fs.getNode("aruco_bc_markers")
    assert int(fs.getNode("aruco_bc_mInfoType").real()) == 0 and [int(ms.at(i).getNode("id").real()) for i in range(ms.size())] == ids
    c0 = ms.at(0).getNode("corners")
    want0 = np.array(expected["boards"]["board_pix"]["markers"][0]["corners"]) * 1.5  # same layout, 150 instead of 100 pixels
    assert np.array_equal(np.array([[c0.at(k).at(d).real() for d in range(3)] for k in range(4)]), want0)


@pytest.mark.parametrize("name,cfgname", [("board", "board_pix"), ("chessboard", "chessboard_pix")])
def test_board_detector_goldens(built, frames, expected, name, cfgname):
    """Aruco.Board / Aruco.Multi (test/core_tests.cpp:164-228) through the device: detect without camera, then
    BoardDetector::detect -> the golden board pose; also vs cv2 when importable, incl. the outlier re-solve."""
    from aruco_b200 import BoardConfiguration, BoardDetector
    bd = BoardDetector()
    cfg = BoardConfiguration.from_dict(expected["boards"][cfgname])
    K, D = intrinsics(expected, name)
    markers = bd.getMarkerDetector().detect(frames[name])
    prob, board = bd.detect(markers, cfg, K, D, 1.0)
    g = expected["goldens"][name]
    assert len(board) == len(g["markers"]) and abs(prob - len(board) / cfg.size()) < 1e-6
    assert [m.id for m in board] == [m["id"] for m in g["markers"]] and all(m.ssize == 1.0 for m in board)
    assert np.abs(board.Rvec - np.array(g["rvec"])).max() < 1e-4 and np.abs(board.Tvec - np.array(g["tvec"])).max() < 1e-4
    try:
        from oracle import cv2_oracle as o
        assert o.cv2 is not None
    except Exception:
        return
    ms = [{"id": m.id, "corners": m.corners} for m in markers]
    for thr, yperp in ((-1.0, False), (1.5, False), (0.8, True)):
        bd.set_repj_err_thres(thr)
        bd.setYPerpendicular(yperp)
        prob, board = bd.detect(markers, cfg, K, D, 1.0)
        ref = o.board_detect(ms, expected["boards"][cfgname], K, D, 1.0, thr, yperp)
        assert rel_err(board.Rvec, ref["rvec"]) < 1e-4 and rel_err(board.Tvec, ref["tvec"]) < 1e-4
    # no camera -> no pose, probability 0 (boarddetector.cpp:117-118)
    prob, board = bd.detect(markers, cfg)
    assert prob == 0.0 and board.Rvec is None and len(board) == len(g["markers"])
    # meters configuration needs no marker size
    cfgm = BoardConfiguration.from_dict(expected["boards"][cfgname.replace("_pix", "_meters")])
    bd.set_repj_err_thres(-1.0)
    bd.setYPerpendicular(False)
    prob, board = bd.detect(markers, cfgm, K, D)
    refm = o.board_detect(ms, expected["boards"][cfgname.replace("_pix", "_meters")], K, D)
    assert rel_err(board.Rvec, refm["rvec"]) < 1e-4 and rel_err(board.Tvec, refm["tvec"]) < 1e-4


def test_c_abi_strided_and_unaligned_frames(built, frames, expected):
    """The C ABI takes caller strides: padded rows, gaps between frames and a base pointer that is not 4-byte
    aligned (scalar load path of the threshold kernel) must give the result of the dense, aligned call."""
    import ctypes as C
    from aruco_b200 import MarkerDetector
    from aruco_b200._lib import ab_marker
    det = MarkerDetector(0)
    K, D = intrinsics(expected, "single")
    ref = [det.detect(frames[n], K, D, 1.0) for n in ("single", "board")]
    H, W = 480, 640
    row, gap = W + 37, 1234
    fstride = row * H + gap
    buf = np.zeros(1 + 2 * fstride, np.uint8)
    for off in (0, 1):  # aligned / unaligned base
        view = buf[off:]
        for i, n in enumerate(("single", "board")):
            for y in range(H):
                view[i * fstride + y * row:i * fstride + y * row + W] = frames[n][y]
        out = (ab_marker * (2 * 64))()
        cnt = (C.c_int32 * 2)()
        Kf, Df = np.ascontiguousarray(K.reshape(9)), np.ascontiguousarray(D)
        rc = det._lib.ab_detect_batch(det._h, C.c_void_p(view.ctypes.data), W, H, row, fstride, 2, Kf.ctypes.data_as(C.c_void_p),
                                      Df.ctypes.data_as(C.c_void_p), 1.0, out, 64, cnt)
        assert rc == 0, det._lib.ab_last_error(det._h)
        for i in range(2):
            assert cnt[i] == len(ref[i])
            for j, m in enumerate(ref[i]):
                o = out[i * 64 + j]
                assert o.id == m.id and (np.array(o.corners, np.float32).reshape(4, 2) == m.corners).all()
                assert (np.array(o.rvec) == m.Rvec).all()
    # device-resident frames with a row stride
    import torch
    dev = torch.zeros((2, H, row), dtype=torch.uint8, device="cuda")
    dev[0, :, :W] = torch.from_numpy(frames["single"]).cuda()
    dev[1, :, :W] = torch.from_numpy(frames["board"]).cuda()
    torch.cuda.synchronize()
    det.enqueue_device(dev.data_ptr(), W, H, 2, K, D, 1.0, row_stride=row, frame_stride=row * H)
    res = det.fetch(2, 64)
    for i in range(2):
        assert [m.id for m in res[i]] == [m.id for m in ref[i]]
        assert all((a.corners == b.corners).all() for a, b in zip(res[i], ref[i]))


def test_refine_candidate_lines_worker(det, frames, expected):
    """Public worker refineCandidateLines (markerdetector.h:280, cpp:931-997) against the cv2 oracle's restatement, with and
    without camera (the contour is undistorted only when both matrices are given)."""
    from conftest import have_cv2
    if not have_cv2():
        pytest.skip("cv2 needed")
    from oracle import cv2_oracle as o
    K, D = intrinsics(expected, "single")
    b = o.detect(frames["single"], o.Params(corner_method=0), None, None, -1.0)
    done = 0
    for c in b["candidates"]:
        if c["id"] < 0:
            continue
        for cam in ((None, None), (K, D)):
            ref = o.refine_candidate_lines(np.array(c["quad"], np.float32).reshape(4, 2), c["contour"], cam[0], cam[1])
            got = det.refineCandidateLines(c["quad"], c["contour"], cam[0], cam[1])
            assert np.abs(got - ref).max() < 1e-4, (got, ref)
        done += 1
    assert done == 6
    from aruco_b200 import ArucoError
    with pytest.raises(ArucoError):  # a corner that is not a point of the contour
        det.refineCandidateLines(np.array([[1, 1], [2, 2], [3, 3], [4, 4]], np.float32), b["candidates"][0]["contour"])


def test_two_batches_in_flight(built):
    """ab_enqueue_batch_device twice before ab_fetch_results: results come back in enqueue order and equal the one-at-a-time
    results byte for byte; a third enqueue is an error, not an overwrite."""
    import torch
    from aruco_b200 import ArucoError, MarkerDetector, synth
    K, D = synth.camera_for(1920, 1080)
    fa = np.stack([synth.render_frame(1920, 1080, 50, 80 + i, 2.0)[0] for i in range(3)])
    fb = np.stack([synth.render_frame(1920, 1080, 50, 90 + i, 2.0)[0] for i in range(3)])
    da, db = torch.from_numpy(fa).cuda(), torch.from_numpy(fb).cuda()
    torch.cuda.synchronize()
    d = MarkerDetector(0)
    d.set_stream(torch.cuda.current_stream().cuda_stream)  # the default stream: handle 0
    d.enqueue_device(da.data_ptr(), 1920, 1080, 3, K, D, 0.05)
    ref_a = d.fetch(3, 128)
    d.enqueue_device(db.data_ptr(), 1920, 1080, 3, K, D, 0.05)
    ref_b = d.fetch(3, 128)
    for rep in range(3):
        d.enqueue_device(da.data_ptr(), 1920, 1080, 3, K, D, 0.05)
        d.enqueue_device(db.data_ptr(), 1920, 1080, 3, K, D, 0.05)
        with pytest.raises(ArucoError) as e:
            d.enqueue_device(da.data_ptr(), 1920, 1080, 3, K, D, 0.05)
        assert e.value.code == -5
        got_a, got_b = d.fetch(3, 128), d.fetch(3, 128)
        for got, ref in ((got_a, ref_a), (got_b, ref_b)):
            for f in range(3):
                assert [m.id for m in got[f]] == [m.id for m in ref[f]] and len(got[f]) >= 45
                assert all((x.corners == y.corners).all() and (x.Rvec == y.Rvec).all() for x, y in zip(got[f], ref[f]))
        assert (d.getThresholdedImage(0) == MarkerDetector(0).thresHold(1, fb[0])).all()  # the getters read the batch last fetched


def test_canny_with_parameter_range_and_setter_rollback(det, frames):
    """CANNY ignores the threshold parameters: a parameter range yields the reference's result (its duplicates collapse in the
    too-near filter).  A rejected setter value is rolled back and does not poison later calls."""
    from aruco_b200 import ArucoError
    det.setThresholdMethod(2)
    try:
        plain = det.detect(frames["single"])
        det.setThresholdParamRange(2)
        ranged = det.detect(frames["single"])
        assert [m.id for m in plain] == [m.id for m in ranged] and len(plain) >= 5
        assert all((a.corners == b.corners).all() for a, b in zip(plain, ranged))
        with pytest.raises(ArucoError):
            det.setThresholdParamRange(99)
        det.setThresholdParams(7, 7)  # would re-push the rejected range if it had not been rolled back
    finally:
        det.setThresholdParamRange(0)
        det.setThresholdMethod(1)
