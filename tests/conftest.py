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


@pytest.fixture(scope="session")
def built():
    """Everything compiled (product library for sm_100a + test infrastructure)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "aruco_b200", "lib", "libaruco_b200.so"), os.path.join(ROOT, "oracle", "_build", "liboracle.so"),
            os.path.join(ROOT, "tests", "_build", "libhostcheck.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
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
