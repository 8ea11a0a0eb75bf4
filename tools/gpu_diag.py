#!/usr/bin/env python
"""Stage-by-stage parity diagnostic of the CUDA path against the cv2 oracle (run on the GPU box)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cv2_oracle as o
from aruco_b200 import MarkerDetector, HighlyReliableMarkers, FiducidalMarkers, synth

fr = np.load(os.path.join(ROOT, "tests/golden/frames.npz"))
exp = json.load(open(os.path.join(ROOT, "tests/golden/expected.json")))


def compare(name, grey, P, K=None, D=None, size=-1.0, hrm=None, det=None, verbose=True):
    det = det or MarkerDetector()
    det.setThresholdMethod(P.thres_method); det.setThresholdParams(P.p1, P.p2)
    det.setCornerRefinementMethod(P.corner_method); det.setMinMaxSize(P.min_size, P.max_size)
    det.setWarpSize(P.warp_size); det.enableErosion(P.erosion)
    if P.decoder == o.DEC_HRM:
        det.setMakerDetectorFunction(HighlyReliableMarkers.detect)
    else:
        det.setMakerDetectorFunction(FiducidalMarkers.detect)
    t = time.time()
    ms = det.detect(grey, K, D, size)
    tg = time.time() - t
    r = o.detect(grey, P, K, D, size, hrm)
    th = det.getThresholdedImage(0)
    bad_thr = int((th != r["thres"]).sum())
    q, ids, nr = det.getAllCandidates(0)
    oq = np.array([c["quad"] for c in r["candidates"]]).reshape(-1, 4, 2)
    oids = [c["id"] for c in r["candidates"]]
    same_order = q.shape == oq.shape and (q == oq).all()
    set_g = set(map(lambda a: tuple(a.ravel().tolist()), q)); set_o = set(map(lambda a: tuple(a.ravel().tolist()), oq))
    ids_ok = same_order and list(ids) == oids and all(a == b["nrot"] for a, b, i in zip(nr, r["candidates"], oids) if i >= 0)
    canon_bad = 0; contour_bad = 0
    if same_order:
        for i, c in enumerate(r["candidates"]):
            canon_bad += int((det.getCanonical(0, i) != c["canon"]).sum())
            cc = det.getContour(0, i)
            if cc.shape != c["contour"].shape or (cc != c["contour"]).any(): contour_bad += 1
    gm = [(m.id, m.corners) for m in ms]; om = [(m["id"], m["corners"]) for m in r["markers"]]
    mids_ok = [g[0] for g in gm] == [m[0] for m in om]
    dc = max([np.abs(g[1] - m[1]).max() for g, m in zip(gm, om)], default=0.0) if mids_ok else -1
    dr = dt = 0.0
    if mids_ok and size > 0 and K is not None:
        for g, m in zip(ms, r["markers"]):
            dr = max(dr, np.abs(g.Rvec - m["rvec"]).max() / np.abs(m["rvec"]).max())
            dt = max(dt, np.abs(g.Tvec - m["tvec"]).max() / np.abs(m["tvec"]).max())
    ok = bad_thr == 0 and set_g == set_o and ids_ok and mids_ok and dc < 0.01 and dr < 1e-4 and dt < 1e-4
    print("%-14s %s thr_bad=%d cands gpu=%d oracle=%d set_equal=%s same_order=%s ids_ok=%s canon_bad=%d contour_bad=%d "
          "markers gpu=%d oracle=%d ids=%s dcorner=%.2e drvec=%.2e dtvec=%.2e  (%.0f ms) counters=%s" % (
              name, "OK  " if ok else "FAIL", bad_thr, len(q), len(oq), set_g == set_o, same_order, ids_ok, canon_bad, contour_bad,
              len(gm), len(om), mids_ok, dc, dr, dt, tg * 1e3, det.counters()), flush=True)
    return ok


def main():
    allok = True
    for name in ("single", "board", "chessboard"):
        intr = exp["intrinsics"][name]
        K = np.array(intr["K"], np.float32).reshape(3, 3); D = np.array(intr["D"], np.float32)
        allok &= compare(name, fr[name], o.Params(), K, D, 1.0)
        allok &= compare(name + "/nocam", fr[name], o.Params())
        allok &= compare(name + "/subpix", fr[name], o.Params(corner_method=o.SUBPIX), K, D, 1.0)
        allok &= compare(name + "/none", fr[name], o.Params(corner_method=o.NONE), K, D, 1.0)
        allok &= compare(name + "/erode", fr[name], o.Params(erosion=True), K, D, 1.0)
        allok &= compare(name + "/fixed", fr[name], o.Params(thres_method=o.FIXED_THRES, p1=100), K, D, 1.0)
    text = exp["dictionaries"]["d4x4_100"]
    HighlyReliableMarkers.loadDictionary(text)
    Dh = o.HrmDictionary.from_yaml_text(text)
    intr = exp["intrinsics"]["hrm"]
    K = np.array(intr["K"], np.float32).reshape(3, 3); D = np.array(intr["D"], np.float32)
    Ph = o.Params(p1=21, p2=7, warp_size=48, min_size=0.005, decoder=o.DEC_HRM)
    allok &= compare("hrm", fr["hrm"], Ph, K, D, 1.0, Dh)
    allok &= compare("refine_fail", fr["refine_fail"], Ph, K, D, 1.0, Dh)
    for (W, H, n, s, sig) in [(1920, 1080, 50, 1, 2.0), (1920, 1080, 50, 2, 4.0), (3840, 2160, 100, 1, 2.0), (3840, 2160, 100, 0, 2.0)]:
        g, _ = synth.render_frame(W, H, n, s, sig)
        K, D = synth.camera_for(W, H)
        allok &= compare("synth%dx%d s%d" % (W, H, s), g, o.Params(), K, D, 0.05)
        allok &= compare("synth%d subpix" % W, g, o.Params(corner_method=o.SUBPIX), K, D, 0.05)
    print("ALL OK" if allok else "SOME FAILED")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
