"""Host-side logic that needs no GPU: synthetic generator, dictionary parsing, sharding (gloo, world 2)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_synth_is_deterministic_and_truthful():
    from aruco_b200 import synth
    a, ta = synth.render_frame(640, 480, 12, seed=3, sigma=1.0, marker_px=56)
    b, tb = synth.render_frame(640, 480, 12, seed=3, sigma=1.0, marker_px=56)
    assert (a == b).all() and ta["ids"] == tb["ids"] and a.dtype == np.uint8 and a.shape == (480, 640)
    c, _ = synth.render_frame(640, 480, 12, seed=4, sigma=1.0, marker_px=56)
    assert (a != c).any()
    assert len(set(ta["ids"])) == 12 and all(0 <= i < 1024 for i in ta["ids"])
    # marker layout of createMarkerImage (arucofidmarkers.cpp:220-229): id 0 -> every row is word 10000
    bits = synth.fiducidal_bits(0)
    assert bits[0].sum() == 0 and (bits[1:6, 1] == 1).all() and bits[1:6, 2:6].sum() == 0


def test_synth_frame_is_detected_by_the_oracle(built):
    from aruco_b200 import synth
    from oracle import native
    from oracle.cv2_oracle import Params
    g, truth = synth.render_frame(1920, 1080, 50, seed=1, sigma=2.0)
    K, D = synth.camera_for(1920, 1080)
    r = native.detect(g, Params(), K, D, 0.05, debug=False)
    ids = [m["id"] for m in r["markers"]]
    assert set(ids) <= set(truth["ids"]) and len(ids) >= 45
    # detected corners sit on the rendered marker corners (any cyclic order)
    tc = {i: c for i, c in zip(truth["ids"], truth["corners"])}
    for m in r["markers"]:
        d = np.linalg.norm(m["corners"][:, None, :] - tc[m["id"]][None, :, :], axis=2).min(axis=1)
        assert d.max() < 2.5


def test_hrm_dictionary_parsing(expected):
    from aruco_b200 import HighlyReliableMarkers
    for n in range(4, 9):
        assert HighlyReliableMarkers.loadDictionary(expected["dictionaries"]["d%dx%d_100" % (n, n)])
        dn, bits, tau0, rate = HighlyReliableMarkers._dict
        assert dn == n and bits.shape == (100, n * n) and rate == 1.0 and tau0 in (4, 7, 12, 17, 23)


def test_shard_ranges_cover_every_frame_once():
    from aruco_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 255, 256, 257):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, c = shard_range(n, world, r)
                seen += list(range(s, s + c))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from aruco_b200 import synth
    from aruco_b200.sharding import gather_in_frame_order, shard_range
    from oracle import native
    from oracle.cv2_oracle import Params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard_range(n_frames, world, rank)
    local = []
    for f in range(start, start + count):
        g, _ = synth.render_frame(640, 480, 6, seed=f, sigma=1.0, marker_px=70)
        local.append(sorted(m["id"] for m in native.detect(g, Params(), debug=False)["markers"]))
    full = gather_in_frame_order(local, n_frames, world, rank)
    dist.barrier()
    if rank == 0:
        q.put(full)
    dist.destroy_process_group()


def test_two_rank_sharding_gives_the_unsharded_result(built):
    """N>1 path on CPU: world_size 2 over gloo; each rank detects its shard (with the CPU oracle standing in for
    the device, this is host logic only) and the gathered per-frame results equal the single-process run."""
    import torch.multiprocessing as mp
    from aruco_b200 import synth
    from oracle import native
    from oracle.cv2_oracle import Params
    n_frames, world, port = 5, 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = []
    for f in range(n_frames):
        g, _ = synth.render_frame(640, 480, 6, seed=f, sigma=1.0, marker_px=70)
        ref.append(sorted(m["id"] for m in native.detect(g, Params(), debug=False)["markers"]))
    assert full == ref and sum(len(r) for r in ref) >= 20


def test_bench_reference_arm_runs_without_a_gpu(built):
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) on the single-frame config: one JSON line
    with the contract's keys, the same metric / config dict the GPU arm would emit, no GPU launches."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["markers_per_frame"] == 6.0 and d["config"]["workload"].startswith("C1")
