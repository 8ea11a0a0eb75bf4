#!/usr/bin/env python
"""Study (run on the GPU box): does keeping two batches in flight (two contexts, two streams) raise the HBM-resident
throughput?  Prints frames/s for 1 and 2 contexts on the bench workload (device-timed)."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aruco_b200 import MarkerDetector, synth

W, H, B = 3840, 2160, int(os.environ.get("STUDY_BATCH", "256"))
dev = torch.device("cuda", 0)
scenes = [synth.render_frame(W, H, 100, seed=1000 + i, sigma=0.0, as_float=True)[0] for i in range(4)]
gen = torch.Generator(device=dev); gen.manual_seed(1)
frames = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
for i in range(B):
    clean = torch.from_numpy(np.asarray(scenes[i % 4], np.float32)).to(dev)
    frames[i] = torch.clamp(torch.round(clean + torch.randn((H, W), generator=gen, device=dev) * 2.0), 0, 255).to(torch.uint8)
K, D = synth.camera_for(W, H)
for nctx in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nctx)]
    dets = []
    for s in streams:
        d = MarkerDetector(0); d.set_stream(s.cuda_stream); d.reserve(W, H, B); dets.append(d)
    def run(steps):
        pending = []
        for it in range(steps):
            d = dets[it % nctx]
            if len(pending) == nctx:
                pending.pop(0).fetch(B, 128, raw=True)
            d.enqueue_device(frames.data_ptr(), W, H, B, K, D, 0.05)
            pending.append(d)
        for d in pending:
            d.fetch(B, 128, raw=True)
    run(4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 12
    run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("contexts", nctx, "frames/s", round(steps * B / dt), "ms/step", round(1e3 * dt / steps, 3), flush=True)
    del dets
