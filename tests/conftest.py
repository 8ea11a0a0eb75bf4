import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def frames():
    return np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))


@pytest.fixture(scope="session")
def expected():
    with open(os.path.join(ROOT, "tests", "golden", "expected.json")) as f:
        return json.load(f)


def source_hash():
    """sha256 over the product sources, as the Makefile computes SRCHASH."""
    import glob
    import hashlib
    files = sorted(glob.glob(os.path.join(ROOT, "aruco_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "aruco_b200", "csrc", "*.cuh"))
                   + [os.path.join(ROOT, "include", "aruco_b200.h")], key=lambda p: os.path.relpath(p, ROOT))
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def library_source_hash():
    """The hash embedded in the shipped binary's ab_version(), read in a child process (a stale library must not stay
    loaded in this one)."""
    import subprocess
    lib = os.path.join(ROOT, "aruco_b200", "lib", "libaruco_b200.so")
    if not os.path.exists(lib):
        return None
    code = "import ctypes,sys; l=ctypes.CDLL(sys.argv[1]); l.ab_version.restype=ctypes.c_char_p; print(l.ab_version().decode())"
    try:
        out = subprocess.run([sys.executable, "-c", code, lib], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    return out.strip().rsplit("src:", 1)[-1] if "src:" in out else None


@pytest.fixture(scope="session")
def built():
    """Everything compiled (product library for sm_100a + test infrastructure) FROM THE SOURCES IN THE TREE: the shipped .so
    files are git-ignored build products, so the library's embedded source hash is checked and a stale one is rebuilt."""
    import subprocess
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "aruco_b200", "lib", "libaruco_b200.so"), os.path.join(ROOT, "oracle", "_build", "liboracle.so"),
            os.path.join(ROOT, "tests", "_build", "libhostcheck.so")]
    if library_source_hash() != source_hash():
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", ROOT, "-B", "aruco_b200/lib/libaruco_b200.so"], env=env)
        assert library_source_hash() == source_hash(), "the rebuilt library does not carry the hash of the sources"
    if not all(os.path.exists(p) for p in need):
        g.build()
    else:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", ROOT, "all"], env=env)  # oracle / hostcheck / facade follow their sources by mtime
    return True


def intrinsics(expected, name):
    intr = expected["intrinsics"][name]
    return np.array(intr["K"], np.float32).reshape(3, 3), np.array(intr["D"], np.float32)


def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


needs_cv2 = pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")

# north-star tolerances (BASELINE.json): corners 0.01 px, Rvec/Tvec 1e-4 relative
CORNER_TOL = 0.01
POSE_RTOL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
