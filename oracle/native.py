"""ctypes loader of the C++ CPU oracle (oracle/aruco_oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class orc_params(C.Structure):
    _fields_ = [("thres_method", C.c_int32), ("p1", C.c_double), ("p2", C.c_double), ("corner_method", C.c_int32),
                ("min_size", C.c_float), ("max_size", C.c_float), ("warp_size", C.c_int32), ("border_dist", C.c_float),
                ("locked_corners", C.c_int32), ("erosion", C.c_int32), ("decoder", C.c_int32),
                ("set_y_perpendicular", C.c_int32), ("p1_range", C.c_int32)]


class orc_marker(C.Structure):
    _fields_ = [("id", C.c_int32), ("has_pose", C.c_int32), ("corners", C.c_float * 8), ("ssize", C.c_float),
                ("pad_", C.c_float), ("rvec", C.c_double * 3), ("tvec", C.c_double * 3)]


class orc_dict(C.Structure):
    _fields_ = [("n", C.c_int32), ("count", C.c_int32), ("tau0", C.c_int32), ("rate", C.c_float), ("bits", C.c_void_p)]


class orc_debug(C.Structure):
    _fields_ = [("thres", C.c_void_p), ("n_contours", C.c_int32), ("n_candidates", C.c_int32),
                ("cap_candidates", C.c_int32), ("quads", C.c_void_p), ("ids", C.c_void_p), ("nrot", C.c_void_p),
                ("canon", C.c_void_p)]


_lib = None


def build():
    src = os.path.join(_HERE, "aruco_oracle.cpp")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB_PATH, src])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_approx_poly.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
        _lib.orc_threshold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        _lib.orc_solve_pnp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(orc_params), C.c_void_p, C.c_void_p, C.c_float,
                                    C.POINTER(orc_dict), C.c_void_p, C.c_int, C.POINTER(orc_debug)]
        _lib.orc_detect_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(orc_params), C.c_void_p, C.c_void_p,
                                          C.c_float, C.POINTER(orc_dict), C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_params(P) -> orc_params:
    """P: oracle.cv2_oracle.Params-like object (same field names)."""
    return orc_params(P.thres_method, float(P.p1), float(P.p2), P.corner_method, P.min_size, P.max_size, P.warp_size,
                      P.border_dist, int(P.locked_corners), int(P.erosion), P.decoder, int(P.set_y_perpendicular),
                      int(getattr(P, "p1_range", 0)))


def make_dict(codes, n, tau0, rate=1.0):
    bits = np.ascontiguousarray(np.array([[c == "1" for c in s] for s in codes], np.uint8))
    d = orc_dict(n, len(codes), tau0, rate, bits.ctypes.data)
    d._keep = bits
    return d


def dict_from_yaml_text(text, rate=1.0):
    kv = {}
    for line in text.splitlines():
        if ":" in line and not line.startswith("%"):
            k, v = line.split(":", 1)
            kv[k.strip()] = v.strip().strip('"')
    nm, n, tau0 = int(kv["nmarkers"]), int(kv["markersize"]), int(kv["tau0"])
    return make_dict([kv["marker_%d" % i] for i in range(nm)], n, tau0, rate)


def _markers(buf, n):
    out = []
    for i in range(n):
        m = buf[i]
        e = {"id": m.id, "corners": np.array(m.corners, np.float32).reshape(4, 2)}
        if m.has_pose:
            e["rvec"] = np.array(m.rvec)
            e["tvec"] = np.array(m.tvec)
        out.append(e)
    return out


def detect(grey, P, K=None, D=None, marker_size=-1.0, hrm=None, cap=512, debug=True):
    """One frame through the C++ oracle. Returns dict(markers, thres, n_contours, quads, ids, nrot, canon)."""
    lib = load()
    grey = np.ascontiguousarray(grey)
    H, W = grey.shape
    Kf = None if K is None else np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    Df = None if D is None else np.ascontiguousarray(np.asarray(D, np.float32).reshape(-1)[:5])
    pp = make_params(P)
    buf = (orc_marker * cap)()
    S = P.warp_size
    res = {}
    dbg = None
    if debug:
        thres = np.zeros((H, W), np.uint8)
        quads = np.zeros((cap, 4, 2), np.float32)
        ids = np.zeros(cap, np.int32)
        nrot = np.zeros(cap, np.int32)
        canon = np.zeros((cap, S, S), np.uint8)
        dbg = orc_debug(thres.ctypes.data, 0, 0, cap, quads.ctypes.data, ids.ctypes.data, nrot.ctypes.data, canon.ctypes.data)
    n = lib.orc_detect(_p(grey), W, H, C.byref(pp), _p(Kf), _p(Df), float(marker_size), C.byref(hrm) if hrm is not None else None,
                       buf, cap, C.byref(dbg) if dbg is not None else None)
    if n < 0:
        raise RuntimeError("oracle: orc_detect returned %d" % n)
    res["markers"] = _markers(buf, n)
    if debug:
        nc = dbg.n_candidates
        res.update(thres=thres, n_contours=dbg.n_contours, quads=quads[:nc], ids=ids[:nc], nrot=nrot[:nc], canon=canon[:nc])
    return res


def detect_batch(frames, P, K=None, D=None, marker_size=-1.0, hrm=None, cap=256, threads=0):
    lib = load()
    frames = np.ascontiguousarray(frames)
    n, H, W = frames.shape
    Kf = None if K is None else np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    Df = None if D is None else np.ascontiguousarray(np.asarray(D, np.float32).reshape(-1)[:5])
    pp = make_params(P)
    buf = (orc_marker * (cap * n))()
    counts = np.zeros(n, np.int32)
    rc = lib.orc_detect_batch(_p(frames), W, H, n, C.byref(pp), _p(Kf), _p(Df), float(marker_size),
                              C.byref(hrm) if hrm is not None else None, buf, cap, _p(counts), threads)
    if rc != 0:
        raise RuntimeError("oracle: orc_detect_batch failed")
    return [_markers((orc_marker * cap).from_buffer(buf, f * cap * C.sizeof(orc_marker)), int(counts[f])) for f in range(n)]


def max_threads():
    return load().orc_max_threads()
