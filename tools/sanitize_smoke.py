#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aruco_b200 import MarkerDetector, HighlyReliableMarkers, FiducidalMarkers, BoardDetector, BoardConfiguration

fr = np.load(os.path.join(ROOT, "tests/golden/frames.npz"))
exp = json.load(open(os.path.join(ROOT, "tests/golden/expected.json")))
intr = exp["intrinsics"]["single"]
K = np.array(intr["K"], np.float32).reshape(3, 3); D = np.array(intr["D"], np.float32)
det = MarkerDetector(0)
n = 0
for cm in (3, 2, 1, 0):
    det.setCornerRefinementMethod(cm)
    n += len(det.detect(fr["single"], K, D, 1.0))
det.enableLockedCornersMethod(True); n += len(det.detect(fr["chessboard"], K, D, 1.0)); det.enableLockedCornersMethod(False)
det.setCornerRefinementMethod(3)
det.enableErosion(True); n += len(det.detect(fr["board"])); det.enableErosion(False)
det.setThresholdMethod(0); det.setThresholdParams(100, 0); n += len(det.detect(fr["board"])); det.setThresholdMethod(1)
det.setThresholdParams(23, 7); n += len(det.detect(fr["board"]))   # generic (non-templated) threshold kernel
det.setThresholdParams(7, 7)
n += len(det.detect(fr["single_bgr"], K, D, 1.0))
n += sum(len(m) for m in det.detect_batch(np.stack([fr["single"], fr["board"], fr["hrm"]]), K, D, 1.0))
rng = np.random.default_rng(0)
n += len(det.detect(rng.integers(0, 256, (101, 333), dtype=np.uint8)))
HighlyReliableMarkers.loadDictionary(exp["dictionaries"]["d4x4_100"])
det.setMakerDetectorFunction(HighlyReliableMarkers.detect); det.setThresholdParams(21, 7); det.setWarpSize(48); det.setMinMaxSize(0.005, 0.5)
n += len(det.detect(fr["hrm"], K, D, 1.0))
det.setMakerDetectorFunction(FiducidalMarkers.detect); det.setThresholdParams(7, 7); det.setWarpSize(56); det.setMinMaxSize(0.04, 0.5)
bd = BoardDetector(detector=det)
ms = det.detect(fr["board"])
prob, board = bd.detect(ms, BoardConfiguration.from_dict(exp["boards"]["board_pix"]), K, D, 1.0)
print("sanitize smoke ok, markers:", n, "board prob", prob)
