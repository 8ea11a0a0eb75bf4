#!/usr/bin/env python
"""SURVEY 8(e) check: a fixed batch sharded over N GPUs returns byte-identical result arrays to the unsharded run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shard_check.py

Every rank renders the same seeded frames, detects its contiguous shard on its own GPU (no collective on the data
path) and the per-frame raw marker bytes are gathered on the host in frame order; rank 0 also runs the whole batch
on its GPU and compares byte for byte.  Prints one JSON line.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def frame_bytes(buf, counts, cap, f):
    n = int(counts[f])
    return bytes(C.string_at(C.addressof(buf) + f * cap * C.sizeof(buf._type_), n * C.sizeof(buf._type_)))


def main():
    import torch
    import torch.distributed as dist
    from aruco_b200 import MarkerDetector, synth
    from aruco_b200.sharding import gather_in_frame_order, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")  # host-side gather only
    W, H, n_frames, cap = 1920, 1080, int(os.environ.get("SHARD_CHECK_FRAMES", "24")), 128
    frames = np.stack([synth.render_frame(W, H, 50, seed=500 + i, sigma=2.0)[0] for i in range(n_frames)])
    K, D = synth.camera_for(W, H)
    det = MarkerDetector(local)

    def run(sub):
        dev = torch.from_numpy(sub).cuda()
        det.reserve(W, H, len(sub))
        det.enqueue_device(dev.data_ptr(), W, H, len(sub), K, D, 0.05)
        buf, counts = det.fetch(len(sub), cap, raw=True)
        torch.cuda.synchronize()
        return [frame_bytes(buf, counts, cap, f) for f in range(len(sub))]

    start, count = shard_range(n_frames, world, rank)
    mine = run(frames[start:start + count])
    full = gather_in_frame_order(mine, n_frames, world, rank)
    if rank == 0:
        whole = run(frames)
        same = [a == b for a, b in zip(full, whole)]
        print(json.dumps({"check": "sharded == unsharded (raw ab_marker bytes per frame)", "n_gpus": world, "frames": n_frames,
                          "markers": sum(len(b) for b in whole) // 96, "identical_frames": int(sum(same)), "ok": bool(all(same))}))
        rc = 0 if all(same) else 1
    else:
        rc = 0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
