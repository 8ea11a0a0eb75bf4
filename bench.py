#!/usr/bin/env python
"""Benchmark of the hot path MarkerDetector::detect on B200 (contract: see README / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], "C4"): synthetic 3840x2160 grey frames, 100 Fiducidal markers each, batch
of 256 frames per GPU, defaults of the reference (ADPT_THRES 7/7, LINES refinement, per-marker PnP).  One
"step" = one pass of the whole hot path over one batch.  Frames are independent, so N GPUs = N shards, no
collective on the data path (weak scaling: 256 frames per GPU).

  value   whole-job frames/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e     the same through the C ABI call ab_detect_batch with HOST (pinned) frames: H2D + kernels + D2H
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the CPU oracle port (oracle/aruco_oracle.cpp, OpenMP frame-parallel) on a bounded sample

--impl reference times that CPU oracle port with all host threads on the same workload (the reference's own
C++ cannot be built here: no OpenCV C++ -- DESIGN.md section 3).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, N_MARKERS, BATCH, SIGMA = 3840, 2160, 100, 256, 2.0
N_BASE = 8          # distinct rendered scenes; every frame of a batch gets its own noise realisation
MARKER_SIZE = 0.05
KERNELS_PER_BATCH = 14  # threshold, scan_starts, trace, trace_long, emit_long, emit, polygon, frame_filter, homography, sample, otsu, identify, refine_lines, finalize


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner to stdout) must not be
# able to add to it: fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, dev):
        self.rows, self.proc, self.dev = [], None, dev

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so do not ask OpenMP)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_params():
    from oracle.cv2_oracle import Params
    return Params()


def base_scenes(n_base, rank):
    from aruco_b200 import synth
    t = time.time()
    scenes = [synth.render_frame(W, H, N_MARKERS, seed=1000 * rank + i, as_float=True)[0] for i in range(n_base)]
    log("[bench] rendered %d base scenes in %.1fs" % (n_base, time.time() - t))
    return scenes


def run_reference(args):
    """CPU arm: the oracle port, OpenMP frame-parallel over all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from aruco_b200 import synth
    from oracle import native
    native.load()
    threads = host_threads()
    sample = max(2 * threads, 32)  # frames per step: ~2-3 CPU-seconds per step on every thread
    rng = np.random.default_rng(7)
    scenes = base_scenes(min(N_BASE, 4), 0)
    frames = np.stack([np.clip(np.rint(scenes[i % len(scenes)] + rng.normal(0, SIGMA, (H, W)).astype(np.float32)), 0, 255).astype(np.uint8)
                       for i in range(sample)])
    K, D = synth.camera_for(W, H)
    P = oracle_params()
    for _ in range(args.warmup):
        native.detect_batch(frames[:threads], P, K, D, MARKER_SIZE, threads=threads)
    t0 = time.perf_counter()
    nm = 0
    for _ in range(args.steps):
        res = native.detect_batch(frames, P, K, D, MARKER_SIZE, threads=threads)
        nm += sum(len(r) for r in res)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {"impl": "reference", "metric": "frames_per_s_4k_100markers", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "mpix_per_s": fps * W * H / 1e6,
            "config": {"workload": "C4: synthetic 3840x2160 grey, 100 Fiducidal markers/frame, ADPT_THRES 7/7 + LINES + PnP",
                       "frames_per_step": sample, "noise_sigma": SIGMA},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d frames/step x %d steps, oracle/aruco_oracle.cpp OpenMP frame-parallel (reference C++ "
                                       "unbuildable: no OpenCV C++)" % (sample, args.steps)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "markers_per_frame": nm / (args.steps * sample), "gpu_launches": 0}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from aruco_b200 import MarkerDetector, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    # ---- synthetic batch, resident in HBM: N_BASE rendered scenes + per-frame Gaussian noise (sigma 2) --------
    scenes = base_scenes(N_BASE, rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    frames = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    for i in range(B):
        clean = torch.from_numpy(scenes[i % N_BASE]).to(dev)
        noise = torch.randn((H, W), generator=gen, device=dev, dtype=torch.float32) * SIGMA
        frames[i] = torch.clamp(torch.round(clean + noise), 0, 255).to(torch.uint8)
    del clean, noise
    K, D = synth.camera_for(W, H)
    stream = torch.cuda.current_stream()

    det = MarkerDetector(local)
    det.set_stream(stream.cuda_stream)
    det.reserve(W, H, B)
    cap = 128

    def step():
        det.enqueue_device(frames.data_ptr(), W, H, B, K, D, MARKER_SIZE)
        return det.fetch(B, cap, raw=True)

    # ---- parity gate before timing: two frames against the CPU oracle -------------------------------------
    buf, counts = step()
    parity = {"checked_frames": 0, "ids_exact": None}
    if rank == 0:
        from oracle import native
        P = oracle_params()
        ok = True
        for f in (0, B - 1):
            host = frames[f].cpu().numpy()
            ref = native.detect(host, P, K, D, MARKER_SIZE, debug=False)["markers"]
            got = [buf[f * cap + i].id for i in range(counts[f])]
            ok &= got == [m["id"] for m in ref]
            for i, m in enumerate(ref):
                if not ok:
                    break
                ok &= float(np.abs(np.array(buf[f * cap + i].corners).reshape(4, 2) - m["corners"]).max()) < 0.01
        parity = {"checked_frames": 2, "ids_exact": bool(ok)}
        log("[bench] parity vs oracle on 2 frames: %s; markers/frame=%.1f counters=%s" % (ok, float(np.mean(counts)), det.counters()))
        if not ok:
            raise SystemExit("parity check against the oracle failed -- refusing to report a number")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------------------
    # Two batches are kept in flight (two contexts on two streams, results of batch n fetched while batch n+1 runs): the
    # latency-bound kernels at the end of a batch (k_finalize keeps 7 % of the warps busy) overlap the bandwidth-bound
    # start of the next one.  tools/pipeline_study.py: 46.8 k -> 51.0 k frames/s.  Every step still enqueues one batch
    # of B frames and fetches its markers.
    stream_b = torch.cuda.Stream(device=dev)
    det_b = MarkerDetector(local)
    det_b.set_stream(stream_b.cuda_stream)
    det_b.reserve(W, H, B)
    dets = [det, det_b]

    def run_steps(n):
        pending, total = [], 0
        for it in range(n):
            d = dets[it % 2]
            if len(pending) == 2:
                total += sum(pending.pop(0).fetch(B, cap, raw=True)[1])
            d.enqueue_device(frames.data_ptr(), W, H, B, K, D, MARKER_SIZE)
            pending.append(d)
        for d in pending:
            total += sum(d.fetch(B, cap, raw=True)[1])
        return total

    run_steps(max(args.warmup, 4))
    kernel_ms = {k: 0.0 for k in det.KERNELS}
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1a, e1b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    stream_b.wait_event(e0)  # nothing of the timed region starts before e0 on either stream
    n_markers = run_steps(args.steps)
    e1a.record(stream)
    e1b.record(stream_b)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max(e0.elapsed_time(e1a), e0.elapsed_time(e1b))
    del det_b, dets
    # per-kernel durations (for the roofline): a separate, untimed pass -- the library pipelines sub-batches over
    # several streams in the timed loop, per-kernel CUDA events need everything on one stream
    det.enable_timing(True)
    step()
    for _ in range(3):
        step()
        for k, v in det.kernel_ms().items():
            kernel_ms[k] += v / 3
    det.enable_timing(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    fps = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end through the C ABI with host frames (pinned): H2D + kernels + D2H every step ---------------
    e2e = None
    if not args.skip_e2e:
        host = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
        host.copy_(frames)
        det2 = MarkerDetector(local)
        det2.set_stream(stream.cuda_stream)
        det2.reserve(W, H, 32)  # 32-frame chunks: H2D of chunk c+1 overlaps the kernels of chunk c
        from aruco_b200._lib import ab_marker
        import ctypes as C
        out = (ab_marker * (B * cap))()
        cnts = (C.c_int32 * B)()
        Kf = np.ascontiguousarray(K.reshape(9))
        Df = np.ascontiguousarray(D)

        def step_e2e():
            rc = det2._lib.ab_detect_batch(det2._h, C.c_void_p(host.data_ptr()), W, H, W, W * H, B, Kf.ctypes.data_as(C.c_void_p),
                                           Df.ctypes.data_as(C.c_void_p), MARKER_SIZE, out, cap, cnts)
            det2._check(rc)

        for _ in range(2):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms2 = max(e0.elapsed_time(e1), wall_ms)  # copies run on the library's copy stream: take the host clock too
        t = torch.tensor([ms2], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_fps = world * B * args.steps / (float(t.item()) / 1e3)
        same = list(cnts) == list(counts)
        e2e = {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
               "d2h_bytes_per_step": B * cap * C.sizeof(ab_marker) + 4 * B, "chunk_frames": 32,
               "same_counts_as_device_path": bool(same)}
        del det2, host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    dom = max(kernel_ms, key=kernel_ms.get)
    # algorithmic bytes per launch: threshold reads the grey frame and
    # writes the u8 binarised frame (the 1/8 B/px packed copy is not counted); the scan and the walkers read the packed
    # image; the other kernels work per candidate (< 1 % of a frame) and are given the whole-path figure 3*W*H.
    packed = W * H / 8.0 * B
    alg_bytes = {"threshold": 2.0 * W * H * B, "scan_starts": packed, "trace": packed, "trace_long": packed, "emit": packed}.get(dom, 3.0 * W * H * B)
    achieved = alg_bytes / (kernel_ms[dom] / 1e3) / 1e9
    n_cands_batch = int(det.counters()["candidates"])
    warp_gbs = n_cands_batch * 2.0 * 56 * 56 / (kernel_ms["sample"] / 1e3) / 1e9
    thr_gbs = 2.0 * W * H * B / (kernel_ms["threshold"] / 1e3) / 1e9
    # DRAM bytes per launch of the threshold kernel from the committed ncu --set full capture of this command
    # (profiles/r1p_ncu_threshold_pair_raw.csv: dram__bytes_read.sum 2.238 GB + dram__bytes_write.sum 2.361 GB at 256 x 4K);
    # only quoted for the workload it was captured on
    traffic = 4.599e9 if (dom == "threshold" and B == BATCH) else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (ncu, profiles/r1p_ncu_threshold_pair_raw.csv)",
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src, "kernel_ms": kernel_ms,
                "threshold_kernel": {"achieved": thr_gbs, "frac": thr_gbs / peak},
                # the warp stage (k_homography + k_sample): N_cand * (S^2 gathered + S^2 written) algorithmic bytes
                "warp_kernel": {"achieved": warp_gbs, "frac": warp_gbs / peak, "candidates_per_batch": n_cands_batch},
                "whole_path": {"algorithmic_bytes_per_frame": 3 * W * H, "achieved": fps / world * 3 * W * H / 1e9,
                               "frac": fps / world * 3 * W * H / 1e9 / peak}}

    # ---- CPU baseline: the oracle port on a bounded sample of the same workload ------------------------------
    cpu = None
    if not args.skip_cpu:
        from oracle import native
        threads = host_threads()
        sample = min(B, 256)  # the whole batch: ~20 CPU-seconds of work spread over the host threads
        hostf = frames[:sample].cpu().numpy()
        native.detect_batch(hostf[:threads], oracle_params(), K, D, MARKER_SIZE, threads=threads)
        t0 = time.perf_counter()
        native.detect_batch(hostf, oracle_params(), K, D, MARKER_SIZE, threads=threads)
        dt = time.perf_counter() - t0
        cpu = {"value": sample / dt, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "%d frames of the same batch, oracle/aruco_oracle.cpp, OpenMP frame-parallel on %d threads" % (sample, threads)}

    line = {"metric": "frames_per_s_4k_100markers", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "mpix_per_s": fps * W * H / 1e6,
            "config": {"workload": "C4: synthetic 3840x2160 grey, 100 Fiducidal markers/frame, ADPT_THRES 7/7 + LINES + PnP",
                       "frames_per_gpu_per_step": B, "noise_sigma": SIGMA, "distinct_scenes": N_BASE,
                       "l2": "inputs larger than L2 (batch = %.2f GB per GPU vs 126 MB L2)" % (B * W * H / 1e9),
                       "parallelism": "frame shards, one per GPU, no collective", "pipelining": "2 batches in flight per GPU (two contexts / streams); every step enqueues one batch and fetches its markers"},
            "markers_per_frame": n_markers / (args.steps * B), "parity": parity, "clocks": clocks, "e2e": e2e,
            "gpu_launches": KERNELS_PER_BATCH * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "target_frames_per_s": 2000}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
