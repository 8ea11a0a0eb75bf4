"""marker / board generators.  The oracle restatement is pinned by the reference's own PNGs and
board YAMLs (CPU tests); the device path is compared with both (GPU tests, through the C ABI) and closed with a
render -> detect round trip."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def render_goldens():
    return np.load(os.path.join(ROOT, "tests", "golden", "render.npz"))


def board_ids(expected, name):
    return [m["id"] for m in expected["boards"][name]["markers"]]


def test_oracle_generators_reproduce_the_reference_pngs(render_goldens, expected):
    from oracle import cv2_oracle as o
    assert np.array_equal(o.create_marker_image(471, 500), render_goldens["marker_471_500"])            # Aruco.CreateMarker
    assert np.array_equal(o.create_marker_image(471, 500, locked=True), render_goldens["locked_marker_471_500"])
    img, used, _ = o.create_board_image(0, 4, 6, 150, 30, board_ids(expected, "board_pix"))
    assert np.array_equal(img, render_goldens["board_4x6_150_30"]) and used == board_ids(expected, "board_pix")
    img, used, _ = o.create_board_image(1, 5, 7, 300, 0, board_ids(expected, "chessboard_pix"))
    assert np.array_equal(img, render_goldens["chessboard_5x7_300"]) and used == board_ids(expected, "chessboard_pix")


def hrm_codes(expected, name="d4x4_100"):
    return [l.split('"')[1] for l in expected["dictionaries"][name].splitlines() if l.startswith("marker_")]


def read_board_yaml(name):
    import cv2
    fs = cv2.FileStorage(os.path.join(ROOT, "tests", "golden", "yaml", name), cv2.FILE_STORAGE_READ)
    ms = fs.getNode("aruco_bc_markers")
    ids = [int(ms.at(i).getNode("id").real()) for i in range(ms.size())]
    corners = np.array([[[ms.at(i).getNode("corners").at(k).at(d).real() for d in range(3)] for k in range(4)] for i in range(ms.size())], np.float32)
    return ids, corners


def test_oracle_hrm_board_reproduces_the_reference_png_and_yaml(render_goldens, expected):
    """testdata/hrm/boards/board4x4.{png,yml} = HighlyReliableMarkers::createBoardImage(4x4, d4x4_100).  The shipped YAML was
    written when MarkerCode folded ids with 1 << pos; the current source folds with 2 << pos (highlyreliablemarkers.cpp:176,
    SURVEY B.4), so today's getId() is exactly twice the stored id."""
    from oracle import cv2_oracle as o
    img, ids, corners = o.hrm_create_board_image(4, 4, hrm_codes(expected), 4)
    want_ids, want_c = read_board_yaml("hrm__boards__board4x4.yml")
    assert np.array_equal(img, render_goldens["hrm_board4x4"]) and np.array_equal(corners, want_c)
    assert ids == [2 * i for i in want_ids]


def test_oracle_board_configuration_matches_the_reference_yaml(expected):
    """board_pix.yml was written by createBoardImage(4x6, 100, 20): same ids -> same corner coordinates."""
    from oracle import cv2_oracle as o
    want = np.array([m["corners"] for m in expected["boards"]["board_pix"]["markers"]], np.float32)
    _, _, corners = o.create_board_image(0, 4, 6, 100, 20, board_ids(expected, "board_pix"))
    assert np.array_equal(corners, want)


@pytest.mark.gpu
def test_device_marker_images(built, render_goldens):
    from aruco_b200 import ArucoError, FiducidalMarkers
    from oracle import cv2_oracle as o
    assert np.array_equal(FiducidalMarkers.createMarkerImage(471, 500), render_goldens["marker_471_500"])
    assert np.array_equal(FiducidalMarkers.createMarkerImage(471, 500, False, True), render_goldens["locked_marker_471_500"])
    for mid, size, locked in [(0, 7, False), (1023, 56, False), (341, 175, True), (7, 99, False), (600, 1001, True), (5, 10, True)]:
        assert np.array_equal(FiducidalMarkers.createMarkerImage(mid, size, False, locked), o.create_marker_image(mid, size, locked))
    for mid in (0, 471, 1023):
        assert np.array_equal(FiducidalMarkers.getMarkerMat(mid), o.create_marker_image(mid, 7)[1:6, 1:6] // 255)
    with pytest.raises(ArucoError):
        FiducidalMarkers.createMarkerImage(1024, 100)        # CV_Assert(0 <= id && id < 1024)
    with pytest.raises(ArucoError):
        FiducidalMarkers.createMarkerImage(5, 100, True)     # watermark: not reproduced, refused loudly


@pytest.mark.gpu
def test_device_board_images(built, render_goldens, expected):
    from aruco_b200 import ArucoError, FiducidalMarkers
    from oracle import cv2_oracle as o
    ids = board_ids(expected, "board_pix")
    img, cfg = FiducidalMarkers.createBoardImage((4, 6), 150, 30, ids)
    assert np.array_equal(img, render_goldens["board_4x6_150_30"]) and cfg.ids == ids and cfg.mInfoType == 0
    img, cfg = FiducidalMarkers.createBoardImage((4, 6), 100, 20, ids)
    assert np.array_equal(cfg.objPoints, np.array([m["corners"] for m in expected["boards"]["board_pix"]["markers"]], np.float32))
    cids = board_ids(expected, "chessboard_pix")
    img, cfg = FiducidalMarkers.createBoardImage_ChessBoard((5, 7), 300, cids)
    assert np.array_equal(img, render_goldens["chessboard_5x7_300"]) and cfg.ids == cids
    rng = np.random.default_rng(3)
    pool = rng.permutation(1024)[:64].tolist()
    for kind, gw, gh, ms, md, center in [(0, 3, 2, 70, 11, True), (1, 4, 4, 56, 0, False), (1, 3, 5, 63, 0, True), (2, 5, 4, 49, 13, False),
                                         (2, 2, 2, 100, 0, True)]:
        want_img, want_ids, want_c = o.create_board_image(kind, gw, gh, ms, md, pool, center)
        if kind == 0:
            img, cfg = FiducidalMarkers.createBoardImage((gw, gh), ms, md, pool)
        elif kind == 1:
            img, cfg = FiducidalMarkers.createBoardImage_ChessBoard((gw, gh), ms, pool, center)
        else:
            img, cfg = FiducidalMarkers.createBoardImage_Frame((gw, gh), ms, md, pool, center)
        assert np.array_equal(img, want_img) and cfg.ids == want_ids and np.array_equal(cfg.objPoints, want_c)
    with pytest.raises(ArucoError):
        FiducidalMarkers.createBoardImage((4, 6), 100, 20, ids[:5])  # not enough ids


@pytest.mark.gpu
def test_device_hrm_marker_image_and_round_trip(built, expected):
    from aruco_b200 import HighlyReliableMarkers, MarkerDetector, render
    from oracle import cv2_oracle as o
    text = expected["dictionaries"]["d5x5_100"]
    codes = [l.split('"')[1] for l in text.splitlines() if l.startswith("marker_")]
    for code, pix in [(codes[0], 70), (codes[17], 100), (codes[99], 57)]:
        bits = [c == "1" for c in code]
        assert np.array_equal(render.hrmMarkerImage(code, pix), o.hrm_marker_image(bits, 5, pix))
    # round trip: a rendered HRM marker on a white page is detected with its dictionary index
    page = np.full((480, 640), 255, np.uint8)
    m = render.hrmMarkerImage(codes[42], 140)
    page[100:100 + m.shape[0], 200:200 + m.shape[1]] = m
    saved = HighlyReliableMarkers._dict
    try:
        HighlyReliableMarkers.loadDictionary(text)
        det = MarkerDetector()
        det.setMakerDetectorFunction(HighlyReliableMarkers.detect)
        det.setWarpSize(56)
        assert [mk.id for mk in det.detect(page)] == [42]
    finally:
        HighlyReliableMarkers._dict = saved


@pytest.mark.gpu
def test_device_hrm_board_image(built, render_goldens, expected):
    from aruco_b200 import render
    from oracle import cv2_oracle as o
    img, cfg = render.hrmCreateBoardImage((4, 4), hrm_codes(expected))
    want_ids, want_c = read_board_yaml("hrm__boards__board4x4.yml")
    assert np.array_equal(img, render_goldens["hrm_board4x4"]) and np.array_equal(cfg.objPoints, want_c)
    assert cfg.ids == [2 * i for i in want_ids]
    codes6 = hrm_codes(expected, "d6x6_100")
    img, cfg = render.hrmCreateBoardImage((5, 3), codes6)
    want_img, want_ids6, want_c6 = o.hrm_create_board_image(5, 3, codes6, 6)
    # BoardConfiguration::ids is vector<int>: getId() (unsigned) above 2^31 wraps negative, compare modulo 2^32
    assert np.array_equal(img, want_img) and [i & 0xFFFFFFFF for i in cfg.ids] == want_ids6 and np.array_equal(cfg.objPoints, want_c6)


@pytest.mark.gpu
def test_rendered_board_round_trip_full_size(built):
    """Size-independent property: every marker of a device-rendered 10x10 board (the C4 marker count) is detected with
    its id, and BoardDetector recovers a frontal pose from the generated configuration."""
    from aruco_b200 import BoardDetector, FiducidalMarkers
    ids = np.random.default_rng(11).permutation(1024)[:100].tolist()
    img, cfg = FiducidalMarkers.createBoardImage((10, 10), 175, 35, ids)
    page = np.full((img.shape[0] + 200, img.shape[1] + 200), 255, np.uint8)
    page[100:-100, 100:-100] = img
    bd = BoardDetector()
    markers = bd.getMarkerDetector().detect(page)
    assert sorted(m.id for m in markers) == sorted(ids)
    H, W = page.shape
    K = np.array([[W, 0, W / 2], [0, W, H / 2], [0, 0, 1]], np.float32)
    prob, board = bd.detect(markers, cfg, K, np.zeros(5, np.float32), 0.05)
    assert prob == 1.0 and abs(board.Tvec[2] - W * 0.05 / 175) < 1e-3 * board.Tvec[2]  # Z = f * size / pixels
