"""Study: how many batches in flight pay?  N contexts (each keeps two batches in flight) fed round-robin on one GPU.
usage: python tools/probe/inflight.py [n_contexts ...]"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, bench
from aruco_b200._lib import ab_marker

def main():
    ns = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
    base = bench.Workload("C4", 256, 0, 0, torch)
    works = [base]
    for n in ns:
        while len(works) < n:
            w = bench.Workload.__new__(bench.Workload)
            w.__dict__.update(base.__dict__)
            w._out = None
            from aruco_b200 import FiducidalMarkers, MarkerDetector
            det = MarkerDetector(0)
            w.stream = torch.cuda.Stream(device=base.dev)
            det.set_stream(w.stream.cuda_stream)
            P = bench.oracle_params(base.cfg)
            det.setThresholdMethod(P.thres_method); det.setThresholdParams(P.p1, P.p2)
            det.setCornerRefinementMethod(P.corner_method); det.setMinMaxSize(P.min_size, P.max_size)
            det.setWarpSize(P.warp_size); det.enableErosion(P.erosion)
            det.setMakerDetectorFunction(FiducidalMarkers.detect)
            det.reserve(base.W, base.H, base.B)
            w.det = det
            works.append(w)
        use = works[:n]
        for depth in (1, 2):
            def loop(steps):
                pend = [0] * n
                outs = []
                for w in use:
                    if getattr(w, "_out", None) is None:
                        w._out = (ab_marker * (w.B * w.cap))(); w._cnt = (C.c_int32 * w.B)()
                        w._Kf = np.ascontiguousarray(np.asarray(w.K, np.float32).reshape(9))
                        w._Df = np.ascontiguousarray(np.asarray(w.D, np.float32).reshape(-1)[:5])
                tot = 0
                for s in range(steps):
                    i = s % n
                    w = use[i]; lib = w.det._lib
                    if pend[i] == depth:
                        w.det._check(lib.ab_fetch_results(w.det._h, w._out, w.cap, w._cnt)); pend[i] -= 1; tot += sum(w._cnt)
                    w.det._check(lib.ab_enqueue_batch_device(w.det._h, C.c_void_p(w.frames.data_ptr()), w.W, w.H, w.W, w.W * w.H, w.B,
                                                             w._Kf.ctypes.data_as(C.c_void_p), w._Df.ctypes.data_as(C.c_void_p), w.size))
                    pend[i] += 1
                for i, w in enumerate(use):
                    while pend[i]:
                        w.det._check(w.det._lib.ab_fetch_results(w.det._h, w._out, w.cap, w._cnt)); pend[i] -= 1; tot += sum(w._cnt)
                return tot
            loop(2 * n * depth)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            steps = 24
            tot = loop(steps)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / steps
            print("contexts %d x depth %d = %d in flight: %.3f ms/step, %.0f frames/s, markers/step %d" % (n, depth, n * depth, ms, 256e3 / ms, tot // steps), flush=True)

main()
