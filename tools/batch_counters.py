#!/usr/bin/env python
"""Work counters of one bench-shaped batch (run on the GPU box): starts, contours, pool points, quads, candidates, markers."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aruco_b200 import MarkerDetector, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H = 3840, 2160
frames = np.stack([synth.render_frame(W, H, 100, seed=1000 + i, sigma=2.0)[0] for i in range(n)])
K, D = synth.camera_for(W, H)
det = MarkerDetector()
res = det.detect_batch(frames, K, D, 0.05)
c = det.counters()
print({k: round(v / n, 1) for k, v in c.items()}, "per frame over", n, "frames")
