"""the reference's YAML wire formats in C++ (include/aruco/serialization.hpp), no GPU needed.

For every fixture the reference ships (tests/golden/yaml/, copied byte for byte by make_golden.py) the C++ reader must
return exactly what cv::FileStorage returns (cv2.FileStorage here), and what the C++ writer emits must be readable by
cv::FileStorage with identical values -- i.e. files travel both ways between this library and the reference."""
import os
import subprocess

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML = os.path.join(ROOT, "tests", "golden", "yaml")
TOOL = os.path.join(ROOT, "tests", "_build", "yaml_tool")


def run_tool(kind, src, dst):
    r = subprocess.run([TOOL, kind, src, dst], capture_output=True, text=True, timeout=60)
    return r.returncode, r.stdout.strip().splitlines(), r.stderr


def cv_markers(node):
    out = []
    for i in range(node.size()):
        m = node.at(i)
        c = m.getNode("corners")
        e = {"id": int(m.getNode("id").real()),
             "corners": [np.float32(c.at(k).at(j).real()) for k in range(c.size()) for j in range(2)]}
        for key in ("Rvec", "Tvec"):
            v = m.getNode(key)
            e[key] = [v.at(k).real() for k in range(3)] if not v.empty() else None
        out.append(e)
    return out


def tool_markers(lines):
    out = []
    for l in lines:
        p = l.split()
        assert p[0] == "marker"
        has = int(p[2])
        n = int(p[9])
        out.append({"id": int(p[1]), "Rvec": [float(x) for x in p[3:6]] if has else None, "Tvec": [float(x) for x in p[6:9]] if has else None,
                    "corners": [np.float32(x) for x in p[10:10 + 2 * n]]})
    return out


def same_markers(a, b):
    assert len(a) == len(b) and len(a) > 0
    for x, y in zip(a, b):
        assert x["id"] == y["id"] and x["Rvec"] == y["Rvec"] and x["Tvec"] == y["Tvec"]  # f64: bit-exact text round trip
        assert len(x["corners"]) == 8 and all(p == q for p, q in zip(x["corners"], y["corners"]))


@pytest.mark.parametrize("name", ["single__expected.yml", "hrm__expected.yml"])
def test_marker_lists_both_ways(built, name, tmp_path):
    src, dst = os.path.join(YAML, name), str(tmp_path / "out.yml")
    rc, lines, err = run_tool("markers", src, dst)
    assert rc == 0, err
    fs_src, fs_dst = cv2.FileStorage(src, cv2.FILE_STORAGE_READ), cv2.FileStorage(dst, cv2.FILE_STORAGE_READ)  # keep alive
    ref = cv_markers(fs_src.getNode("Markers"))
    same_markers(tool_markers(lines), ref)
    same_markers(cv_markers(fs_dst.getNode("Markers")), ref)


@pytest.mark.parametrize("name", ["board__expected.yml", "chessboard__expected.yml"])
def test_boards_both_ways(built, name, tmp_path):
    src, dst = os.path.join(YAML, name), str(tmp_path / "out.yml")
    rc, lines, err = run_tool("board", src, dst)
    assert rc == 0, err
    for path in (src, dst):
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        b = fs.getNode("Board")
        head = lines[0].split()
        assert head[0] == "board" and head[1] == "1"
        assert [float(x) for x in head[2:5]] == [b.getNode("Rvec").at(k).real() for k in range(3)]
        assert [float(x) for x in head[5:8]] == [b.getNode("Tvec").at(k).real() for k in range(3)]
        same_markers(tool_markers(lines[1:]), cv_markers(b.getNode("Markers")))


@pytest.mark.parametrize("name", ["single__intrinsics.yml", "hrm__intrinsics.yml", "board__intrinsics.yml", "chessboard__intrinsics.yml"])
def test_camera_parameters_both_ways(built, name, tmp_path):
    src, dst = os.path.join(YAML, name), str(tmp_path / "out.yml")
    rc, lines, err = run_tool("camera", src, dst)
    assert rc == 0, err
    p = lines[0].split()
    for path in (src, dst):
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        K = fs.getNode("camera_matrix").mat().astype(np.float32).ravel()  # readFromXMLFile converts to f32
        D = fs.getNode("distortion_coefficients").mat().astype(np.float32).ravel()
        assert [int(p[1]), int(p[2])] == [int(fs.getNode("image_width").real()), int(fs.getNode("image_height").real())]
        assert all(np.float32(a) == b for a, b in zip(p[3:12], K))
        assert all(np.float32(a) == b for a, b in zip(p[12:17], list(D[:5]) + [np.float32(0)] * (5 - min(5, D.size))))


@pytest.mark.parametrize("name", ["board__board_pix.yml", "board__board_meters.yml", "chessboard__chessboardinfo_pix.yml"])
def test_board_configuration_both_ways(built, name, tmp_path):
    src, dst = os.path.join(YAML, name), str(tmp_path / "out.yml")
    rc, lines, err = run_tool("boardconf", src, dst)
    assert rc == 0, err
    for path in (src, dst):
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        ms = fs.getNode("aruco_bc_markers")
        head = lines[0].split()
        assert int(head[1]) == ms.size() == int(fs.getNode("aruco_bc_nmarkers").real()) and int(head[2]) == int(fs.getNode("aruco_bc_mInfoType").real())
        for i in range(ms.size()):
            p = lines[1 + i].split()
            c = ms.at(i).getNode("corners")
            assert int(p[1]) == int(ms.at(i).getNode("id").real())
            want = [np.float32(c.at(k).at(d).real()) for k in range(4) for d in range(3)]
            assert all(np.float32(a) == b for a, b in zip(p[2:14], want))


@pytest.mark.parametrize("name", ["hrm__dictionaries__d4x4_100.yml", "hrm__dictionaries__d6x6_100.yml", "mask__dictionary.yml"])
def test_dictionary_both_ways(built, name, tmp_path):
    src, dst = os.path.join(YAML, name), str(tmp_path / "out.yml")
    rc, lines, err = run_tool("dict", src, dst)
    assert rc == 0, err
    for path in (src, dst):
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        n = int(fs.getNode("nmarkers").real())
        head = lines[0].split()
        assert [int(x) for x in head[1:]] == [n, int(fs.getNode("markersize").real()), int(fs.getNode("tau0").real())]
        assert [l.split()[1] for l in lines[1:]] == [fs.getNode("marker_%d" % i).string() for i in range(n)]


def test_reader_errors_mirror_the_reference(built, tmp_path):
    """readFromXMLFile throws on a file without a camera matrix (cameraparameters.cpp:198-199); a board file without
    aruco_bc_nmarkers is 'invalid file type' (serialization.cpp:95-96)."""
    bad = tmp_path / "bad.yml"
    bad.write_text("%YAML:1.0\nimage_width: 640\nimage_height: 480\n")
    rc, _, err = run_tool("camera", str(bad), str(tmp_path / "o.yml"))
    assert rc == 1 and "does not contains valid camera matrix" in err
    rc, _, err = run_tool("boardconf", str(bad), str(tmp_path / "o.yml"))
    assert rc == 1 and "invalid file type" in err
