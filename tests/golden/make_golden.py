#!/usr/bin/env python
"""Builds tests/golden/* from the reference's own fixtures (run HERE, where /root/reference exists).

  python tests/golden/make_golden.py

Outputs (committed; the GPU box has no /root/reference):
  frames.npz      grey (cv2 4.13 BGR2GRAY of the reference PNGs) + the BGR `single` frame
  expected.json   the reference's golden markers testdata/{single,hrm,board,chessboard}/expected.yml
                  (+ board pose), intrinsics (as f32, like CameraParameters), HRM dictionaries,
                  board configurations
Nothing here is reference SOURCE: only its test data (images, golden YAMLs, dictionaries).
"""
import json
import os
import sys

import cv2
import numpy as np

REF = "/root/reference/testdata"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_markers(node):
    out = []
    for i in range(node.size()):
        m = node.at(i)
        e = {"id": int(m.getNode("id").real())}
        c = m.getNode("corners")
        e["corners"] = [[c.at(k).at(0).real(), c.at(k).at(1).real()] for k in range(c.size())]
        for key in ("Rvec", "Tvec"):
            v = m.getNode(key)
            if not v.empty():
                e[key.lower()] = [v.at(k).real() for k in range(3)]
        out.append(e)
    return out


def read_intrinsics(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    K = fs.getNode("camera_matrix").mat().astype(np.float32)
    D = fs.getNode("distortion_coefficients").mat().astype(np.float32).reshape(-1)[:5]
    return {"K": [float(x) for x in K.reshape(-1)], "D": [float(x) for x in D],
            "width": int(fs.getNode("image_width").real()), "height": int(fs.getNode("image_height").real())}


def read_board_cfg(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    ms = fs.getNode("aruco_bc_markers")
    out = {"mInfoType": int(fs.getNode("aruco_bc_mInfoType").real()), "markers": []}
    for i in range(ms.size()):
        m = ms.at(i)
        c = m.getNode("corners")
        out["markers"].append({"id": int(m.getNode("id").real()),
                               "corners": [[c.at(k).at(j).real() for j in range(3)] for k in range(4)]})
    return out


def main():
    frames = {}
    for name, rel in [("single", "single/image-test.png"), ("hrm", "hrm/image-test.png"),
                      ("board", "board/image-test.png"), ("chessboard", "chessboard/chessboard_frame.png"),
                      ("refine_fail", "hrm/refine-fail.png")]:
        bgr = cv2.imread(os.path.join(REF, rel))
        frames[name] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        if name == "single":
            frames["single_bgr"] = bgr
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **frames)

    exp = {"goldens": {}, "intrinsics": {}, "dictionaries": {}, "boards": {}}
    for name in ("single", "hrm"):
        fs = cv2.FileStorage(os.path.join(REF, name, "expected.yml"), cv2.FILE_STORAGE_READ)
        exp["goldens"][name] = {"markers": read_markers(fs.getNode("Markers"))}
    for name in ("board", "chessboard"):
        fs = cv2.FileStorage(os.path.join(REF, name, "expected.yml"), cv2.FILE_STORAGE_READ)
        b = fs.getNode("Board")
        g = {"markers": read_markers(b.getNode("Markers"))}
        for key in ("Rvec", "Tvec"):
            v = b.getNode(key)
            g[key.lower()] = [v.at(k).real() for k in range(3)]
        exp["goldens"][name] = g
    for name in ("single", "hrm", "board", "chessboard"):
        exp["intrinsics"][name] = read_intrinsics(os.path.join(REF, name, "intrinsics.yml"))
    for n in range(4, 9):
        with open(os.path.join(REF, "hrm/dictionaries/d%dx%d_100.yml" % (n, n))) as f:
            exp["dictionaries"]["d%dx%d_100" % (n, n)] = f.read()
    exp["boards"]["board_pix"] = read_board_cfg(os.path.join(REF, "board/board_pix.yml"))
    exp["boards"]["board_meters"] = read_board_cfg(os.path.join(REF, "board/board_meters.yml"))
    exp["boards"]["chessboard_pix"] = read_board_cfg(os.path.join(REF, "chessboard/chessboardinfo_pix.yml"))
    exp["boards"]["chessboard_meters"] = read_board_cfg(os.path.join(REF, "chessboard/chessboardinfo_meters.yml"))
    with open(os.path.join(OUT, "expected.json"), "w") as f:
        json.dump(exp, f, indent=1)
    # the reference's generator goldens (Aruco.CreateMarker, test/core_tests.cpp:32-75, and the printed boards)
    render = {}
    for key, rel in [("marker_471_500", "board/marker-expected.png"), ("locked_marker_471_500", "board/locked-marker-expected.png"),
                     ("board_4x6_150_30", "board/board.png"), ("chessboard_5x7_300", "chessboard/chessboard.png"),
                     ("hrm_board4x4", "hrm/boards/board4x4.png")]:
        render[key] = cv2.imread(os.path.join(REF, rel), cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(os.path.join(OUT, "render.npz"), **render)
    # raw YAML wire-format fixtures for the C++ reader/writer (tests/test_yaml.py): test DATA, byte for byte
    ydir = os.path.join(OUT, "yaml")
    os.makedirs(ydir, exist_ok=True)
    for rel in ("single/intrinsics.yml", "single/expected.yml", "hrm/expected.yml", "hrm/intrinsics.yml", "board/expected.yml",
                "board/board_pix.yml", "board/board_meters.yml", "board/intrinsics.yml", "chessboard/expected.yml",
                "chessboard/chessboardinfo_pix.yml", "chessboard/intrinsics.yml", "hrm/dictionaries/d4x4_100.yml",
                "hrm/dictionaries/d6x6_100.yml", "mask/dictionary.yml", "hrm/boards/board4x4.yml"):
        with open(os.path.join(REF, rel), "rb") as f:
            data = f.read()
        with open(os.path.join(ydir, rel.replace("/", "__")), "wb") as f:
            f.write(data)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    sys.exit(main())
