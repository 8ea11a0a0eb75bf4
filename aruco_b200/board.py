"""Mirror of the caller of the hot path: aruco::BoardConfiguration / Board / BoardDetector
(src/board.h:56-140, src/boarddetector.{h,cpp}) -- SURVEY 8(f) "next" row 1.  The pose arithmetic (stacked
4*M-point planar solvePnP, reprojection-outlier re-solve, rotateXAxis) runs on the device (ab_detect_board)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import ArucoError, ab_board, ab_board_config, ab_marker
from .detector import Marker, MarkerDetector


class BoardConfiguration:
    PIX, METERS = 0, 1

    def __init__(self, ids=(), objPoints=(), mInfoType=0):
        self.ids = [int(i) for i in ids]
        self.objPoints = np.asarray(objPoints, np.float32).reshape(-1, 4, 3)
        self.mInfoType = int(mInfoType)

    @staticmethod
    def from_dict(d) -> "BoardConfiguration":
        """d: {'mInfoType': 0|1, 'markers': [{'id': .., 'corners': 4x3}, ...]} (the aruco_bc_* YAML layout)."""
        return BoardConfiguration([m["id"] for m in d["markers"]], [m["corners"] for m in d["markers"]], d["mInfoType"])

    def size(self):
        return len(self.ids)

    def getMarkerInfo(self, marker_id):  # src/board.cpp:60-66
        for i, k in enumerate(self.ids):
            if k == marker_id:
                return self.objPoints[i]
        raise ArucoError(_lib.AB_E_INVALID, "Marker with the id given is not found")


class Board(list):
    """vector<Marker> + conf + Rvec/Tvec (src/board.h:103-140)."""

    def __init__(self):
        super().__init__()
        self.conf: Optional[BoardConfiguration] = None
        self.Rvec = None
        self.Tvec = None


class BoardDetector:
    def __init__(self, setYPerpendicular: bool = False, device: int = 0, detector: Optional[MarkerDetector] = None):
        self._setYPerpendicular = bool(setYPerpendicular)
        self.repj_err_thres = -1.0  # boarddetector.cpp:38-41
        self._mdetector = detector or MarkerDetector(device)
        self._bconf: Optional[BoardConfiguration] = None
        self._cam = (None, None)
        self._markerSize = -1.0
        self._vmarkers: List[Marker] = []
        self._boardDetected = Board()

    def setParams(self, bc: BoardConfiguration, camMatrix=None, distCoeff=None, markerSizeMeters: float = -1.0):
        self._bconf, self._cam, self._markerSize = bc, (camMatrix, distCoeff), float(markerSizeMeters)

    def setYPerpendicular(self, enable: bool):
        self._setYPerpendicular = bool(enable)

    def set_repj_err_thres(self, v: float):
        self.repj_err_thres = float(v)

    def getMarkerDetector(self) -> MarkerDetector:
        return self._mdetector

    def getDetectedBoard(self) -> Board:
        return self._boardDetected

    def getDetectedMarkers(self) -> List[Marker]:
        return self._vmarkers

    def detect_image(self, image: np.ndarray) -> float:
        """BoardDetector::detect(const cv::Mat&) (boarddetector.cpp:66-78): markers first (no camera), then the board."""
        self._vmarkers = self._mdetector.detect(image)
        K, D = self._cam
        prob, self._boardDetected = self.detect(self._vmarkers, self._bconf, K, D, self._markerSize)
        return prob

    def detect(self, detectedMarkers: List[Marker], BConf: BoardConfiguration, camMatrix=None, distCoeff=None,
               markerSizeMeters: float = -1.0):
        """boarddetector.cpp:90-204. Returns (prob, Board)."""
        if BConf is None or BConf.size() == 0:
            raise ArucoError(_lib.AB_E_INVALID, "Invalid BoardConfig that is empty")  # CV_Assert (:93)
        det = self._mdetector
        n = len(detectedMarkers)
        buf = (ab_marker * max(n, 1))()
        for i, m in enumerate(detectedMarkers):
            buf[i].id = int(m.id)
            for j, v in enumerate(np.asarray(m.corners, np.float32).reshape(8)):
                buf[i].corners[j] = float(v)
            buf[i].ssize = float(m.ssize)
        ids = np.ascontiguousarray(np.array(BConf.ids, np.int32))
        pts = np.ascontiguousarray(BConf.objPoints.astype(np.float32).reshape(-1))
        cfg = ab_board_config(len(ids), BConf.mInfoType, ids.ctypes.data, pts.ctypes.data)
        Kf, Df = MarkerDetector._cam(camMatrix, distCoeff)
        outm = (ab_marker * max(n, 1))()
        res = ab_board()
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        det._check(det._lib.ab_detect_board(det._h, buf, n, C.byref(cfg), p(Kf), p(Df), float(markerSizeMeters),
                                            float(self.repj_err_thres), int(self._setYPerpendicular), outm, C.byref(res)))
        board = Board()
        board.conf = BConf
        for i in range(res.n_markers):
            board.append(Marker(outm[i].id, np.array(outm[i].corners, np.float32).reshape(4, 2), outm[i].ssize))
        if res.has_pose:
            board.Rvec = np.array(res.rvec, np.float64)
            board.Tvec = np.array(res.tvec, np.float64)
        return float(res.prob), board
