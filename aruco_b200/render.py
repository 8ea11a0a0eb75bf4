"""Mirror of the reference's marker / board generators -- SURVEY 8(f) "next" row 4:
FiducidalMarkers::createMarkerImage / getMarkerMat / createBoardImage / createBoardImage_ChessBoard /
createBoardImage_Frame (src/arucofidmarkers.cpp:213-436) and MarkerCode::getImg (src/highlyreliablemarkers.cpp:234-256).
The pixels are drawn on the device (k_render.cuh) through the C ABI; there is no CPU path.  Differences from the reference,
both deliberate: the text watermark (`addWaterMark`, cv::putText glyph data) is not reproduced, and the board generators
take the marker ids from the caller instead of drawing them with rand()."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ArucoError
from .board import BoardConfiguration
from .detector import MarkerDetector

_ctx = {}


def _det(device: int) -> MarkerDetector:
    if device not in _ctx:
        _ctx[device] = MarkerDetector(device)
    return _ctx[device]


def createMarkerImage(id: int, size: int, addWaterMark: bool = False, locked: bool = False, device: int = 0) -> np.ndarray:
    """FiducidalMarkers::createMarkerImage (arucofidmarkers.cpp:213)."""
    if addWaterMark:
        raise ArucoError(_lib.AB_E_INVALID, "createMarkerImage: the text watermark is not reproduced (cv::putText font data)")
    d = _det(device)
    side = C.c_int(0)
    d._check(d._lib.ab_create_marker_image(d._h, int(id), int(size), int(locked), None, 0, C.byref(side)))
    out = np.empty((side.value, side.value), np.uint8)
    d._check(d._lib.ab_create_marker_image(d._h, int(id), int(size), int(locked), out.ctypes.data_as(C.c_void_p), side.value, C.byref(side)))
    return out


def getMarkerMat(id: int) -> np.ndarray:
    """FiducidalMarkers::getMarkerMat (arucofidmarkers.cpp:266-284): the 5x5 code as 0/1, cut out of a 7-pixel rendering."""
    return (createMarkerImage(id, 7)[1:6, 1:6] // 255).astype(np.uint8)


def _board(kind: int, gridSize: Tuple[int, int], MarkerSize: int, MarkerDistance: int, ids: Sequence[int], centerData: bool, device: int):
    d = _det(device)
    gw, gh = int(gridSize[0]), int(gridSize[1])
    w, h, n = C.c_int(0), C.c_int(0), C.c_int(0)
    d._check(d._lib.ab_create_board_image(d._h, kind, gw, gh, int(MarkerSize), int(MarkerDistance), int(centerData), None, 0, None, 0,
                                          C.byref(w), C.byref(h), None, None, 0, C.byref(n)))
    ids_in = np.asarray(list(ids), np.int32)
    if ids_in.size < n.value:
        raise ArucoError(_lib.AB_E_INVALID, "%d marker ids needed, %d given" % (n.value, ids_in.size))
    img = np.empty((h.value, w.value), np.uint8)
    ids_out = np.zeros(n.value, np.int32)
    corners = np.zeros((n.value, 4, 3), np.float32)
    d._check(d._lib.ab_create_board_image(d._h, kind, gw, gh, int(MarkerSize), int(MarkerDistance), int(centerData),
                                          ids_in.ctypes.data_as(C.c_void_p), int(ids_in.size), img.ctypes.data_as(C.c_void_p), w.value,
                                          C.byref(w), C.byref(h), ids_out.ctypes.data_as(C.c_void_p), corners.ctypes.data_as(C.c_void_p),
                                          n.value, C.byref(n)))
    return img, BoardConfiguration(ids_out.tolist(), corners, BoardConfiguration.PIX)


def createBoardImage(gridSize, MarkerSize: int, MarkerDistance: int, ids: Sequence[int], device: int = 0):
    """FiducidalMarkers::createBoardImage (arucofidmarkers.cpp:283-329) -> (image, BoardConfiguration in pixels, centred)."""
    return _board(0, gridSize, MarkerSize, MarkerDistance, ids, True, device)


def createBoardImage_ChessBoard(gridSize, MarkerSize: int, ids: Sequence[int], centerData: bool = True, device: int = 0):
    """FiducidalMarkers::createBoardImage_ChessBoard (arucofidmarkers.cpp:337-389)."""
    return _board(1, gridSize, MarkerSize, 0, ids, centerData, device)


def createBoardImage_Frame(gridSize, MarkerSize: int, MarkerDistance: int, ids: Sequence[int], centerData: bool = True, device: int = 0):
    """FiducidalMarkers::createBoardImage_Frame (arucofidmarkers.cpp:397-436)."""
    return _board(2, gridSize, MarkerSize, MarkerDistance, ids, centerData, device)


def hrmMarkerImage(code, pixSize: int, device: int = 0) -> np.ndarray:
    """MarkerCode::getImg (highlyreliablemarkers.cpp:234-256). code: n*n string of '0'/'1' or an (n, n) array."""
    bits = np.array([c == "1" for c in code], np.uint8) if isinstance(code, str) else (np.asarray(code).ravel() != 0).astype(np.uint8)
    n = int(round(np.sqrt(bits.size)))
    if n * n != bits.size:
        raise ArucoError(_lib.AB_E_INVALID, "a square code is needed")
    d = _det(device)
    side = C.c_int(0)
    d._check(d._lib.ab_create_hrm_marker_image(d._h, n, None, int(pixSize), None, 0, C.byref(side)))
    out = np.empty((side.value, side.value), np.uint8)
    d._check(d._lib.ab_create_hrm_marker_image(d._h, n, bits.ctypes.data_as(C.c_void_p), int(pixSize), out.ctypes.data_as(C.c_void_p),
                                               side.value, C.byref(side)))
    return out


def hrmCreateBoardImage(gridSize, codes: Sequence[str], device: int = 0):
    """HighlyReliableMarkers::createBoardImage (highlyreliablemarkers.cpp:498-545), non-chromatic.  codes: the dictionary's
    n*n strings of '0'/'1' (the first gridSize.width * gridSize.height are drawn).  Returns (image, BoardConfiguration in
    pixels, ids = MarkerCode::getId())."""
    gw, gh = int(gridSize[0]), int(gridSize[1])
    bits = np.array([[c == "1" for c in code] for code in codes], np.uint8)
    n = int(round(np.sqrt(bits.shape[1])))
    d = _det(device)
    w, h, nm = C.c_int(0), C.c_int(0), C.c_int(0)
    d._check(d._lib.ab_create_hrm_board_image(d._h, gw, gh, n, None, 0, None, 0, C.byref(w), C.byref(h), None, None, 0, C.byref(nm)))
    img = np.empty((h.value, w.value), np.uint8)
    ids = np.zeros(nm.value, np.int32)
    corners = np.zeros((nm.value, 4, 3), np.float32)
    bits = np.ascontiguousarray(bits)
    d._check(d._lib.ab_create_hrm_board_image(d._h, gw, gh, n, bits.ctypes.data_as(C.c_void_p), int(bits.shape[0]), img.ctypes.data_as(C.c_void_p),
                                              w.value, C.byref(w), C.byref(h), ids.ctypes.data_as(C.c_void_p), corners.ctypes.data_as(C.c_void_p),
                                              nm.value, C.byref(nm)))
    return img, BoardConfiguration(ids.tolist(), corners, BoardConfiguration.PIX)
