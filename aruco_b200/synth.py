"""Seeded synthetic frames for the benchmark/parity configs C3/C4/C5.

numpy only (no cv2): white canvas, gx x gy grid of markers, one random homography over the whole
canvas (4 % inset, +-3 % corner jitter), bilinear sampling with white fill, 3x3 Gaussian blur sigma 0.8,
contrast 0.8*I+20, additive Gaussian noise.  Marker bitmaps follow the reference generators' bit layout:
FiducidalMarkers::createMarkerImage (src/arucofidmarkers.cpp:214-229) and MarkerCode::getImg
(src/highlyreliablemarkers.cpp:234-256).  Input generator only -- not on the detection path.
"""
from __future__ import annotations

import numpy as np

_FID_WORDS = (0x10, 0x17, 0x09, 0x0E)


def fiducidal_bits(marker_id: int) -> np.ndarray:
    """7x7 cell matrix (1 = white) of marker `marker_id` (arucofidmarkers.cpp:220-229)."""
    assert 0 <= marker_id < 1024
    m = np.zeros((7, 7), np.uint8)
    for y in range(5):
        val = _FID_WORDS[(marker_id >> (2 * (4 - y))) & 3]
        for x in range(5):
            m[y + 1, x + 1] = (val >> (4 - x)) & 1
    return m


def hrm_bits(code: str, n: int) -> np.ndarray:
    """(n+2)x(n+2) cell matrix of an HRM code string (highlyreliablemarkers.cpp:234-256)."""
    m = np.zeros((n + 2, n + 2), np.uint8)
    m[1:-1, 1:-1] = np.array([c == "1" for c in code], np.uint8).reshape(n, n)
    return m


def _homography(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]
        b[i], b[i + 4] = u, v
    h = np.linalg.solve(A, b)
    return np.append(h, 1.0).reshape(3, 3)


def grid_for(n_markers: int):
    return {50: (10, 5), 100: (10, 10)}.get(n_markers, (int(np.ceil(np.sqrt(n_markers))),) * 2)


def render_frame(W: int, H: int, n_markers: int, seed: int, sigma: float = 2.0, marker_px: int | None = None,
                 hrm_codes=None, hrm_n: int = 0, flips: int = 0, clean: bool = False, as_float: bool = False):
    """Returns (grey u8 HxW, truth) with truth = {'ids': [...], 'corners': (n,4,2) f64 image coords}."""
    rng = np.random.default_rng(seed)
    gx, gy = grid_for(n_markers)
    if marker_px is None:
        # 4K: 7 cells x 27 px = 189 px sides, so that after the canvas homography (4 % inset, +-3 % jitter: the contour
        # shrinks to as little as 0.85x) every contour stays above the detector's minimum length 0.04*3840*4 = 614 px;
        # with the 175 px of round 1 about 5 % of the markers fell below it and were (correctly) rejected
        marker_px = 189 if W >= 3000 else (140 if W >= 1900 else int(0.09 * W))
    canvas = np.full((H, W), 255.0, np.float32)
    ins = 0.04
    x0, y0 = ins * W, ins * H
    cw, ch = (W - 2 * x0) / gx, (H - 2 * y0) / gy
    assert marker_px + 8 <= min(cw, ch), "markers do not fit the grid"
    if hrm_codes is None:
        ids = rng.choice(1024, size=gx * gy, replace=False)
        ncell = 7
    else:
        ids = rng.permutation(len(hrm_codes))[: gx * gy]
        ncell = hrm_n + 2
    cell = marker_px // ncell
    side = cell * ncell
    truth_ids, truth_c = [], []
    k = 0
    for j in range(gy):
        for i in range(gx):
            if k >= n_markers or k >= len(ids):
                break
            mid = int(ids[k])
            k += 1
            bits = fiducidal_bits(mid) if hrm_codes is None else hrm_bits(hrm_codes[mid], hrm_n).copy()
            if flips and hrm_codes is not None:
                inner = bits[1:-1, 1:-1]
                pos = rng.choice(hrm_n * hrm_n, size=flips, replace=False)
                inner.reshape(-1)[pos] ^= 1
            px = int(round(x0 + i * cw + (cw - side) / 2))
            py = int(round(y0 + j * ch + (ch - side) / 2))
            canvas[py:py + side, px:px + side] = np.kron(bits, np.ones((cell, cell), np.uint8)).astype(np.float32) * 255
            truth_ids.append(mid)
            truth_c.append([[px, py], [px + side - 1, py], [px + side - 1, py + side - 1], [px, py + side - 1]])
    src = np.array([[0, 0], [W - 1, 0], [W - 1, H - 1], [0, H - 1]], np.float64)
    base = np.array([[x0, y0], [W - 1 - x0, y0], [W - 1 - x0, H - 1 - y0], [x0, H - 1 - y0]], np.float64)
    jit = rng.uniform(-0.03, 0.03, size=(4, 2)) * np.array([W, H])
    dst = base + jit
    Hm = _homography(src, dst)
    Hi = np.linalg.inv(Hm)
    out = np.empty((H, W), np.float32)
    xs = np.arange(W, dtype=np.float64)
    band = 256
    for ys in range(0, H, band):
        ye = min(H, ys + band)
        yy = np.arange(ys, ye, dtype=np.float64)[:, None]
        den = Hi[2, 0] * xs[None, :] + Hi[2, 1] * yy + Hi[2, 2]
        sx = (Hi[0, 0] * xs[None, :] + Hi[0, 1] * yy + Hi[0, 2]) / den
        sy = (Hi[1, 0] * xs[None, :] + Hi[1, 1] * yy + Hi[1, 2]) / den
        fx, fy = np.floor(sx), np.floor(sy)
        ax, ay = (sx - fx).astype(np.float32), (sy - fy).astype(np.float32)
        ix, iy = fx.astype(np.int64), fy.astype(np.int64)
        inside = (ix >= 0) & (iy >= 0) & (ix < W - 1) & (iy < H - 1)
        ixc, iyc = np.clip(ix, 0, W - 2), np.clip(iy, 0, H - 2)
        v = (canvas[iyc, ixc] * (1 - ax) * (1 - ay) + canvas[iyc, ixc + 1] * ax * (1 - ay)
             + canvas[iyc + 1, ixc] * (1 - ax) * ay + canvas[iyc + 1, ixc + 1] * ax * ay)
        out[ys:ye] = np.where(inside, v, np.float32(255))
    g = np.exp(-1.0 / (2 * 0.8 * 0.8))
    k3 = np.array([g, 1.0, g], np.float32) / np.float32(1 + 2 * g)
    p = np.pad(out, 1, mode="edge")
    out = p[:, :-2] * k3[0] + p[:, 1:-1] * k3[1] + p[:, 2:] * k3[2]
    out = out[:-2] * k3[0] + out[1:-1] * k3[1] + out[2:] * k3[2]
    out = out * np.float32(0.8) + np.float32(20)
    if as_float:  # noise-free f32 image: the caller adds its own noise (bench: on the GPU, per frame)
        tc = np.array(truth_c, np.float64)
        ph = np.concatenate([tc, np.ones(tc.shape[:2] + (1,))], axis=2) @ Hm.T
        return out, {"ids": truth_ids, "corners": ph[..., :2] / ph[..., 2:3], "canvas_corners": tc}
    if not clean and sigma > 0:
        out = out + rng.normal(0.0, sigma, size=out.shape).astype(np.float32)
    grey = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    tc = np.array(truth_c, np.float64)
    ph = np.concatenate([tc, np.ones(tc.shape[:2] + (1,))], axis=2) @ Hm.T
    tc_img = ph[..., :2] / ph[..., 2:3]
    return grey, {"ids": truth_ids, "corners": tc_img, "canvas_corners": tc}


def camera_for(W: int, H: int):
    """Synthetic pinhole camera of SURVEY 8(d): fx=fy=W, cx=W/2, cy=H/2, D=0."""
    K = np.array([[W, 0, W / 2], [0, W, H / 2], [0, 0, 1]], np.float32)
    D = np.zeros(5, np.float32)
    return K, D
