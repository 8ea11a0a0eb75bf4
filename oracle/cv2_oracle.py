"""cv2-driven restatement of aruco::MarkerDetector::detect  --  TEST INFRASTRUCTURE ONLY.

This file is a checker, not a product path: only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it.  Nothing under aruco_b200/ does.

The reference (paroj/aruco, ArUco 1.3.0 fork) is thin glue over OpenCV; all arithmetic lives in the
un-vendored, un-pinned third-party dependency OpenCV (CMakeLists.txt:50 `FIND_PACKAGE(OpenCV REQUIRED)`,
README.md:88 ">= 2.4.9").  The only OpenCV in this image is the Python wheel opencv-python-headless
4.13.0, so this oracle calls *the same OpenCV primitives the reference calls* (adaptiveThreshold,
findContours, approxPolyDP, isContourConvex, getPerspectiveTransform, warpPerspective, threshold(OTSU),
undistortPoints, projectPoints, cornerSubPix, solvePnP, ...) and restates the reference's own glue
line by line.  Parity pin: reproduces the reference's four golden files testdata/{single,hrm,board,
chessboard}/expected.yml (tests/test_oracle_golden.py).

Deterministic resolutions of the reference's undefined behaviour:
  B.1 nRotations initialised to 0; B.2 the min-side>10 filter always passes; B.8 stable sort.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional

import numpy as np

try:  # cv2 is the third-party dependency the reference delegates to
    import cv2
except Exception as _e:  # pragma: no cover
    cv2 = None
    _cv2_err = _e

# enums -- src/markerdetector.h:125,186
FIXED_THRES, ADPT_THRES, CANNY = 0, 1, 2
NONE, HARRIS, SUBPIX, LINES = 0, 1, 2, 3
DEC_FID, DEC_HRM = 0, 1


@dataclasses.dataclass
class Params:
    """Mirror of MarkerDetector's private state; defaults = ctor, src/markerdetector.cpp:235-249."""
    thres_method: int = ADPT_THRES
    p1: float = 7.0
    p2: float = 7.0
    corner_method: int = LINES
    min_size: float = 0.04
    max_size: float = 0.5
    warp_size: int = 56
    border_dist: float = 0.025
    locked_corners: bool = False
    erosion: bool = False          # API-compat extension: cv2.erode(thres, None)
    decoder: int = DEC_FID
    set_y_perpendicular: bool = False
    p1_range: int = 0              # _thresParam1_range (setThresholdParamRange, markerdetector.h:152)


class HrmDictionary:
    """highlyreliablemarkers.cpp:149-180 (MarkerCode::set), :312-328 (loadDictionary), :387-496 (tree)."""

    def __init__(self, codes: List[str], n: int, tau0: int, rate: float = 1.0):
        self.n = n
        self.tau0 = tau0
        self.codes = [np.array([c == "1" for c in s], dtype=np.uint8).reshape(n, n) for s in codes]
        self.correction = int(np.float32(rate) * np.float32((tau0 - 1) // 2))  # :318 (integer division)
        self.rot_bits = [code_rotations(c)[0] for c in self.codes]
        self.ids0 = [code_rotations(c)[1][0] for c in self.codes]
        self._build_tree()

    @staticmethod
    def from_yaml_text(text: str, rate: float = 1.0) -> "HrmDictionary":
        kv = {}
        for line in text.splitlines():
            if ":" in line and not line.startswith("%"):
                k, v = line.split(":", 1)
                kv[k.strip()] = v.strip().strip('"')
        nm, n, tau0 = int(kv["nmarkers"]), int(kv["markersize"]), int(kv["tau0"])
        return HrmDictionary([kv["marker_%d" % i] for i in range(nm)], n, tau0, rate)

    def _build_tree(self):
        # restates BalancedBinaryTree::loadDictionary verbatim (:387-476)
        order = sorted((int(i), k) for k, i in enumerate(self.ids0))
        self.order = order
        sz = len(order)
        levels = 0
        while 2.0 ** levels <= sz:
            levels += 1
        visited = [False] * sz
        root = sz // 2
        visited[root] = True
        self.root = root
        intervals = [(0, root), (root, sz)]
        tree = [[0, 0] for _ in range(sz)]
        tree[root][0] = (0 + root) // 2 if not visited[(0 + root) // 2] else -1
        tree[root][1] = (root + sz) // 2 if not visited[(root + sz) // 2] else -1
        for _ in range(1, levels):
            nint = len(intervals)
            for _j in range(nint):
                lo, hi = intervals.pop()
                center = (hi + lo) // 2
                if not visited[center]:
                    visited[center] = True
                else:
                    continue
                lc, hc = (lo + center) // 2, (center + hi) // 2
                if not visited[lc]:
                    intervals.insert(0, (lo, center))
                    tree[center][0] = lc
                else:
                    tree[center][0] = -1
                if not visited[hc]:
                    intervals.insert(0, (center, hi))
                    tree[center][1] = hc
                else:
                    tree[center][1] = -1
        self.tree = tree

    def find_id(self, ident: int):
        pos = self.root
        while pos != -1:
            pid = self.order[pos][0]
            if pid == ident:
                return self.order[pos][1]
            pos = self.tree[pos][1] if pid < ident else self.tree[pos][0]
        return None


def code_rotations(code: np.ndarray):
    """MarkerCode::set (:149-180): 4 rotated bit strings + 4 folded 32-bit ids (x86 shift semantics, B.4)."""
    n = code.shape[0]
    bits = np.zeros((4, n * n), dtype=np.uint8)
    ids = [0, 0, 0, 0]
    for y in range(n):
        for x in range(n):
            for i in range(4):
                _x, _y = x, y
                if i == 1:
                    _y, _x = x, n - y - 1
                elif i == 2:
                    _y, _x = n - y - 1, n - x - 1
                elif i == 3:
                    _y, _x = n - x - 1, y
                pos = _y * n + _x
                v = int(code[y, x] != 0)
                bits[i, pos] = v
                if v:
                    ids[i] |= (2 << (pos & 31)) & 0xFFFFFFFF
    return bits, ids


# ------------------------------------------------------------------------------------------------
def threshold(grey: np.ndarray, method: int, p1: float, p2: float) -> np.ndarray:
    """MarkerDetector::thresHold, src/markerdetector.cpp:643-677."""
    assert grey.dtype == np.uint8 and grey.ndim == 2
    if method == FIXED_THRES:
        return cv2.threshold(grey, p1, 255, cv2.THRESH_BINARY_INV)[1]
    if method == ADPT_THRES:
        if p1 < 3:
            p1 = 3
        elif int(p1) % 2 != 1:
            p1 = int(p1 + 1)
        return cv2.adaptiveThreshold(grey, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, int(p1), p2)
    if method == CANNY:
        return cv2.Canny(grey, 10, 220)
    raise ValueError(method)


def perimeter(c: np.ndarray) -> np.float32:
    """src/utils.h:37-44 -- f64 norm of f32 differences accumulated in f32."""
    s = np.float32(0)
    for i in range(4):
        d = c[i].astype(np.float32) - c[(i + 1) % 4].astype(np.float32)
        s = np.float32(s + np.float32(np.sqrt(float(d[0]) * float(d[0]) + float(d[1]) * float(d[1]))))
    return s


def size_limits(w: int, h: int, min_size: float, max_size: float):
    """src/markerdetector.cpp:500-501 (f32 product, truncated)."""
    m = max(w, h)
    mn = int(np.float32(np.float32(np.float32(min_size) * np.float32(m)) * np.float32(4)))
    mx = int(np.float32(np.float32(np.float32(max_size) * np.float32(m)) * np.float32(4)))
    return mn, mx


def detect_rectangles(thres, min_size: float, max_size: float):
    """MarkerDetector::detectRectangles, src/markerdetector.cpp:496-635.  `thres`: one binary image or the list
    of threshold images of setThresholdParamRange (candidates are joined in image order, :561-563).
    Returns list of dicts {corners (4,2) f32, contour (n,2) i32 (already reversed if swapped), idx}."""
    images = thres if isinstance(thres, (list, tuple)) else [thres]
    h, w = images[0].shape
    mn, mx = size_limits(w, h, min_size, max_size)
    cands = []
    n_contours_total = 0
    for img in images:
        contours, _ = cv2.findContours(img.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        n_contours_total += len(contours)
        for i, c in enumerate(contours):
            n = c.shape[0]
            if n <= mn or n >= mx:
                continue
            approx = cv2.approxPolyDP(c, float(n) * 0.05, True)
            if approx.shape[0] != 4:
                continue
            if not cv2.isContourConvex(approx):
                continue
            # :542-552 min-side filter is an out-of-bounds no-op (Appendix B.2): always pass
            cands.append({"corners": approx.reshape(4, 2).astype(np.float32), "contour": c.reshape(-1, 2), "idx": i})
    contours = [None] * n_contours_total
    swapped = []
    for cd in cands:
        c = cd["corners"]
        d1, d2 = c[1] - c[0], c[2] - c[0]
        o = np.float32(d1[0] * d2[1]) - np.float32(d1[1] * d2[0])
        if o < 0:
            c[[1, 3]] = c[[3, 1]]
            swapped.append(True)
        else:
            swapped.append(False)
    nC = len(cands)
    remove = [False] * nC
    per = [perimeter(cd["corners"]) for cd in cands]
    for i in range(nC):
        for j in range(i + 1, nC):
            d = cands[i]["corners"].astype(np.float64) - cands[j]["corners"].astype(np.float64)
            dist = np.sqrt((d * d).sum(axis=1))
            if (dist < 6).all():
                if per[i] > per[j]:
                    remove[j] = True
                else:
                    remove[i] = True
    out = []
    for i, cd in enumerate(cands):
        if remove[i]:
            continue
        if swapped[i]:
            cd["contour"] = cd["contour"][::-1].copy()
        out.append(cd)
    return out, len(contours)


def warp(grey: np.ndarray, corners: np.ndarray, S: int) -> np.ndarray:
    """MarkerDetector::warp, src/markerdetector.cpp:684-697."""
    dst = np.array([[0, 0], [S - 1, 0], [S - 1, S - 1], [0, S - 1]], dtype=np.float32)
    M = cv2.getPerspectiveTransform(corners.astype(np.float32), dst)
    return cv2.warpPerspective(grey, M, (S, S), flags=cv2.INTER_NEAREST)


def get_marker_code(bw: np.ndarray, n: int, cell: int) -> np.ndarray:
    """aruco::getMarkerCode, src/arucofidmarkers.cpp:189-204."""
    out = np.zeros((n, n), dtype=np.uint8)
    for y in range(n):
        for x in range(n):
            sq = bw[(y + 1) * cell:(y + 2) * cell, (x + 1) * cell:(x + 2) * cell]
            if int(np.count_nonzero(sq)) > (cell * cell) // 2:
                out[y, x] = 1
    return out


def check_borders(bw: np.ndarray, n: int, cell: int) -> bool:
    """aruco::checkBorders, src/arucofidmarkers.cpp:168-184."""
    for y in range(n):
        inc = 1 if (y == 0 or y == n - 1) else n - 1
        for x in range(0, n, inc):
            sq = bw[y * cell:(y + 1) * cell, x * cell:(x + 1) * cell]
            if int(np.count_nonzero(sq)) > (cell * cell) // 2:
                return False
    return True


_FID_WORDS = np.array([[1, 0, 0, 0, 0], [1, 0, 1, 1, 1], [0, 1, 0, 0, 1], [0, 1, 1, 1, 0]], dtype=np.uint8)


def _hamm_dist_marker(bits: np.ndarray) -> int:
    """src/arucofidmarkers.cpp:74-98."""
    d = 0
    for y in range(5):
        d += int(min(int((bits[y] != w).sum()) for w in _FID_WORDS))
    return d


def fiducidal_detect(canon: np.ndarray):
    """FiducidalMarkers::detect + analyzeMarkerImage, src/arucofidmarkers.cpp:438-452, 100-137.
    Returns (id, nRotations, otsu-binarised canonical, 5x5 bits or None)."""
    _, bw = cv2.threshold(canon, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    sw = bw.shape[0] // 7
    if not check_borders(bw, 7, sw):
        return -1, 0, bw, None
    bits = get_marker_code(bw, 5, sw)
    rots = [bits]
    min_dist = _hamm_dist_marker(bits)
    nrot = 0  # B.1
    for i in range(1, 4):
        rots.append(np.rot90(rots[-1], k=-1).copy())  # out(i,j) = in(cols-j-1, i)  (:63-72)
        d = _hamm_dist_marker(rots[i])
        if d < min_dist:
            min_dist, nrot = d, i
    if min_dist != 0:
        return -1, nrot, bw, bits
    b = rots[nrot]
    mid = 0
    for y in range(5):
        mid |= ((int(b[y, 1]) << 1) | int(b[y, 3])) << (2 * (4 - y))
    return mid, nrot, bw, bits


def hrm_detect(canon: np.ndarray, D: HrmDictionary):
    """HighlyReliableMarkers::detect, src/highlyreliablemarkers.cpp:332-383 (returns dictionary index)."""
    _, bw = cv2.threshold(canon, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    cell = bw.shape[0] // (D.n + 2)
    code = get_marker_code(bw, D.n, cell)
    bits, ids = code_rotations(code)
    for i in range(4):
        pos = D.find_id(ids[i])
        if pos is not None:
            return pos, i, bw, code
    res, min_marker, min_rot = D.n * D.n, 0, 0
    for k, rb in enumerate(D.rot_bits):
        r2, mr = D.n * D.n, 0
        for i in range(4):
            hd = int((rb[0] != bits[i]).sum())
            if hd < r2:
                r2, mr = hd, i
        if r2 < res:
            res, min_marker, min_rot = r2, k, mr
    if res <= D.correction:
        return min_marker, min_rot, bw, code
    return -1, 0, bw, code


# cv::solve(DECOMP_SVD) in CV_32F is OpenCV's own one-sided Jacobi SVD (modules/core/src/lapack.cpp) -- unless the
# build carries a LAPACK HAL, which takes over at >= 25 rows (hal_internal.cpp: HAL_SVD_SMALL_MATRIX_THRESH) with
# sgesdd, whose f32 rounding depends on the BLAS kernels picked for the host CPU.  This wheel has one (OpenBLAS).
#   "jacobi": cv2.solve below 25 rows (= OpenCV's Jacobi, the real thing), the restated Jacobi from 25 rows on
#             -- the arithmetic OpenCV's source defines, identical on every machine; tests/test_oracle_cross.py
#             checks the restatement bit for bit against cv2.solve for every row count below 25.
#   "cv2":    cv2.solve for every size (this wheel: LAPACK from 25 rows on).
LINES_SOLVER = "jacobi"
_HAL_SVD_ROWS = 25


def _seq_sum(v) -> float:
    return float(np.cumsum(np.asarray(v, np.float64))[-1])  # sequential f64 accumulation


def jacobi_svd_solve_f32(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """cv::solve(A[m x 2], B[m], DECOMP_SVD) for CV_32F without a LAPACK HAL: JacobiSVDImpl_<float> on A^T
    (eps = 2*FLT_EPSILON, f64 dot products, f32 rotations) + SVBkSb (f32 products summed in f64)."""
    f32 = np.float32
    m = A.shape[0]
    At = [A[:, 0].astype(f32).copy(), A[:, 1].astype(f32).copy()]
    Vt = np.eye(2, dtype=f32)
    W = [_seq_sum(At[0].astype(np.float64) ** 2), _seq_sum(At[1].astype(np.float64) ** 2)]
    eps = float(np.finfo(f32).eps) * 2
    for _ in range(max(m, 30)):
        p = _seq_sum(At[0].astype(np.float64) * At[1].astype(np.float64))
        if abs(p) <= eps * math.sqrt(W[0] * W[1]):
            break
        p *= 2
        beta = W[0] - W[1]
        gamma = math.hypot(p, beta)
        if beta < 0:
            sn = f32(math.sqrt((gamma - beta) * 0.5 / gamma))
            cs = f32(p / (gamma * float(sn) * 2))
        else:
            cs = f32(math.sqrt((gamma + beta) / (gamma * 2)))
            sn = f32(p / (gamma * float(cs) * 2))
        t0 = (cs * At[0]).astype(f32) + (sn * At[1]).astype(f32)
        t1 = ((-sn) * At[0]).astype(f32) + (cs * At[1]).astype(f32)
        At = [t0.astype(f32), t1.astype(f32)]
        W = [_seq_sum(t0.astype(np.float64) ** 2), _seq_sum(t1.astype(np.float64) ** 2)]
        v0 = (cs * Vt[0]).astype(f32) + (sn * Vt[1]).astype(f32)
        v1 = ((-sn) * Vt[0]).astype(f32) + (cs * Vt[1]).astype(f32)
        Vt = np.stack([v0, v1]).astype(f32)
    Wd = [math.sqrt(_seq_sum(At[0].astype(np.float64) ** 2)), math.sqrt(_seq_sum(At[1].astype(np.float64) ** 2))]
    order = [1, 0] if Wd[0] < Wd[1] else [0, 1]
    w = [f32(Wd[0]), f32(Wd[1])]
    tiny = float(np.finfo(f32).tiny)
    U = [(At[i] * f32(1 / Wd[i] if Wd[i] > tiny else 0.0)).astype(f32) for i in range(2)]
    thr = (float(w[0]) + float(w[1])) * float(f32(2 * np.finfo(np.float64).eps))
    X = np.zeros(2, f32)
    Bf = B.reshape(-1).astype(f32)
    for i in order:
        wi = float(w[i])
        if abs(wi) <= thr:
            continue
        sv = _seq_sum((U[i] * Bf).astype(f32)) * (1 / wi)
        for j in range(2):
            X[j] = f32(float(X[j]) + sv * float(Vt[i, j]))
    return X


def _solve_svd_f32(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    if LINES_SOLVER == "jacobi" and A.shape[0] >= _HAL_SVD_ROWS:
        return jacobi_svd_solve_f32(A, B)
    return cv2.solve(A, B, flags=cv2.DECOMP_SVD)[1].reshape(2)


def _interpolate2dline(pts: np.ndarray):
    """src/markerdetector.cpp:83-130 (f32 SVD least squares)."""
    pts = pts.astype(np.float32)
    minx, maxx = pts[:, 0].min(), pts[:, 0].max()
    miny, maxy = pts[:, 1].min(), pts[:, 1].max()
    A = np.zeros((len(pts), 2), np.float32)
    A[:, 1] = 1
    if np.float32(maxx - minx) > np.float32(maxy - miny):
        A[:, 0] = pts[:, 0]
        X = _solve_svd_f32(A, pts[:, 1:2].copy())
        return np.array([X[0], -1.0, X[1]], np.float32)
    A[:, 0] = pts[:, 1]
    X = _solve_svd_f32(A, pts[:, 0:1].copy())
    return np.array([-1.0, X[0], X[1]], np.float32)


def _cross_point(l1, l2):
    """src/markerdetector.cpp:132-139."""
    A = np.array([[l1[0], l1[1]], [l2[0], l2[1]]], np.float32)
    B = np.array([[-l1[2]], [-l2[2]]], np.float32)
    _, X = cv2.solve(A, B, flags=cv2.DECOMP_SVD)
    return X.reshape(2)


def refine_candidate_lines(corners: np.ndarray, contour: np.ndarray, K, D) -> np.ndarray:
    """MarkerDetector::refineCandidateLines, src/markerdetector.cpp:931-997."""
    n = len(contour)
    ci = [0, 0, 0, 0]
    rc = np.rint(corners).astype(np.int32)  # Point(Point2f) rounds
    for k in range(4):
        m = np.nonzero((contour[:, 0] == rc[k, 0]) & (contour[:, 1] == rc[k, 1]))[0]
        ci[k] = int(m[-1]) if len(m) else 0  # last match wins (B.7)
    if ci[1] > ci[0] and (ci[2] > ci[1] or ci[2] < ci[0]):
        inverse = False
    elif ci[2] > ci[1] and ci[2] < ci[0]:
        inverse = False
    else:
        inverse = True
    inc = -1 if inverse else 1
    c2f = contour.astype(np.float32)
    use_cam = K is not None and D is not None
    if use_cam:
        c2f = cv2.undistortPoints(c2f.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    lines = []
    for l in range(4):
        j = ci[l]
        pts = []
        while j != ci[(l + 1) % 4]:
            pts.append(c2f[j])
            j = (j + inc) % n
        if len(pts) == 1:
            pts.append(c2f[ci[(l + 1) % 4]])
        lines.append(_interpolate2dline(np.array(pts, np.float32)))
    cross = np.array([_cross_point(lines[i], lines[(i - 1) % 4]) for i in range(4)], np.float32)
    if use_cam:
        Kf = K.astype(np.float32)
        p3 = np.ones((4, 3), np.float32)
        p3[:, 0] = (cross[:, 0] - Kf[0, 2]) / Kf[0, 0]
        p3[:, 1] = (cross[:, 1] - Kf[1, 2]) / Kf[1, 1]
        z = np.zeros((3, 1), np.float32)
        cross = cv2.projectPoints(p3.reshape(-1, 1, 3), z, z, K, D)[0].reshape(4, 2).astype(np.float32)
    return cross


def find_corner_maxima(corners: np.ndarray, grey: np.ndarray, wsize: int) -> np.ndarray:
    """findCornerMaxima (locked corners), src/markerdetector.cpp:157-199."""
    out = corners.copy()
    H, W = grey.shape
    for i in range(len(corners)):
        cx, cy = corners[i]
        x0, y0 = max(0, int(cx - wsize)), max(0, int(cy - wsize))
        x1, y1 = min(W, int(cx + wsize)), min(H, int(cy + wsize))
        reg = grey[y0:y1, x0:x1]
        harr = cv2.cornerHarris(reg, 3, 3, 0.04)
        hint = cv2.integral(harr, sdepth=cv2.CV_64F)
        b = 4
        hs = harr.copy()
        for y in range(b, harr.shape[0] - b):
            for x in range(b, harr.shape[1] - b):
                hs[y, x] = np.float32(hint[y + b, x + b] - hint[y + b, x] - hint[y, x + b] + hint[y, x])
        harr = hs
        best = (-1.0, -1.0)
        ccx, ccy = float(reg.shape[1] // 2), float(reg.shape[0] // 2)
        den = np.float32(reg.shape[1] // 2 + reg.shape[0] // 2)
        maxv = 0.0
        for yy in range(harr.shape[0]):
            for xx in range(harr.shape[1]):
                d = np.float32(np.float32(abs(ccx - xx) + abs(ccy - yy)) / den)
                wgt = np.float32(1.0 - float(d))
                v = float(np.float32(wgt * harr[yy, xx]))
                if v > maxv:
                    maxv = v
                    best = (float(xx), float(yy))
        out[i] = (best[0] + x0, best[1] + y0)
    return out


def harris_refine(grey: np.ndarray, corners: np.ndarray) -> np.ndarray:
    """SubPixelCorner::RefineCorner, src/subpixelcorner.cpp:70-189 (1 iteration, D==0: Appendix B.3)."""
    win, ap = 15, 3
    H, W = grey.shape
    coeff = 1.0 / (win * win)
    mx = np.array([np.float32(np.exp(-i * i * coeff)) for i in range(-(win // 2), win // 2 + 1)], np.float32)
    mask = np.outer(mx, mx).astype(np.float32)
    out = corners.copy()
    for k in range(len(corners)):
        est = corners[k].astype(np.float32).copy()
        if est[0] < 0 or est[1] < 0 or est[1] > H or est[1] > W:
            continue
        cur = est.copy()
        local = cv2.getRectSubPix(grey, (win + 2 * (ap // 2), win + 2 * (ap // 2)), (float(cur[0]), float(cur[1])))
        Dx = cv2.Sobel(local, cv2.CV_32F, 1, 0, ksize=ap, scale=1, delta=0)
        Dy = cv2.Sobel(local, cv2.CV_32F, 0, 1, ksize=ap, scale=1, delta=0)
        A = B = C = Dd = E = F = 0.0
        for i in range(ap // 2, win + 1):
            ly = i - win // 2 - ap // 2
            for j in range(ap // 2, win + 1):
                lx = j - win // 2 - ap // 2
                val = float(mask[ly + win // 2, lx + win // 2])
                dxx = float(np.float32(Dx[i, j] * Dx[i, j])) * val
                dyy = float(np.float32(Dy[i, j] * Dy[i, j])) * val
                dxy = float(np.float32(Dx[i, j] * Dy[i, j])) * val
                A += dxx
                B += dxy
                E += dyy
                C += dxx * lx + dxy * ly
                F += dxy * lx + dyy * ly
        det = A * E - B * B
        if abs(det) > np.finfo(np.float64).eps ** 2:
            det = 1.0 / det
            est[0] = np.float32(float(cur[0]) + ((C * E) - (B * F)) * det)
            est[1] = np.float32(float(cur[1]) + ((A * F) - (C * Dd)) * det)
        if abs(float(corners[k, 0]) - float(est[0])) > win or abs(float(corners[k, 1]) - float(est[1])) > win:
            est = corners[k].copy()
        out[k] = est
    return out


def rotate_x_axis(rvec: np.ndarray) -> np.ndarray:
    """aruco::rotateXAxis, src/utils.cpp:16-30 (f32 rotation matrices)."""
    R = cv2.Rodrigues(rvec.reshape(3, 1))[0].astype(np.float32)
    RX = np.eye(3, dtype=np.float32)
    a = np.float32(np.pi / 2)
    RX[1, 1] = np.cos(a)
    RX[1, 2] = -np.sin(a)
    RX[2, 1] = np.sin(a)
    RX[2, 2] = np.cos(a)
    R = (R @ RX).astype(np.float32)
    return cv2.Rodrigues(R)[0].reshape(3).astype(np.float64)


def object_points(size: float) -> np.ndarray:
    """getObjectPoints, src/marker.cpp:91-108."""
    h = np.float32(np.float32(size) / 2.0)
    return np.array([[-h, -h, 0], [-h, h, 0], [h, h, 0], [h, -h, 0]], np.float32)


def valid_region(W: int, H: int, border: float):
    """src/markerdetector.cpp:433-434: Point*float -> saturate_cast<int> (round-half-even); Rect(p1,p2)."""
    b = np.float32(border)
    ob = np.float32(np.float32(1.0) - b)
    x0, y0 = int(np.rint(np.float32(W * b))), int(np.rint(np.float32(H * b)))
    x1, y1 = int(np.rint(np.float32(W * ob))), int(np.rint(np.float32(H * ob)))
    return min(x0, x1), min(y0, y1), max(x0, x1), max(y0, y1)


def detect(image: np.ndarray, P: Params = None, K=None, D=None, marker_size: float = -1.0,
           hrm: Optional[HrmDictionary] = None, keep: bool = True) -> dict:
    """MarkerDetector::detect, src/markerdetector.cpp:302-478.  Returns every intermediate."""
    if cv2 is None:  # pragma: no cover
        raise RuntimeError("cv2 unavailable: %r" % (_cv2_err,))
    P = P or Params()
    if image.ndim == 3:
        grey = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    else:
        grey = image
    H, W = grey.shape
    if K is not None:
        K = np.asarray(K, np.float32).reshape(3, 3)       # CameraParameters holds f32 (cameraparameters.cpp:204)
    if D is not None:
        D = np.asarray(D, np.float32).reshape(1, -1)
    n_t = 2 * P.p1_range + 1
    if n_t == 1:
        images = [threshold(grey, P.thres_method, P.p1, P.p2)]
    else:  # :328-332, param1 = p1 - range + range*i
        images = [threshold(grey, P.thres_method, P.p1 - P.p1_range + P.p1_range * i, P.p2) for i in range(n_t)]
    if P.erosion:
        images = [cv2.erode(t, None) for t in images]
    thres = images[n_t // 2]
    cands, n_contours = detect_rectangles(images, P.min_size, P.max_size)
    S = P.warp_size
    res = {"grey": grey, "thres": thres, "n_contours": n_contours, "candidates": [], "markers": []}
    decoded = []
    for cd in cands:
        canon = warp(grey, cd["corners"], S)
        if P.decoder == DEC_FID:
            mid, nrot, bw, bits = fiducidal_detect(canon)
        else:
            mid, nrot, bw, bits = hrm_detect(canon, hrm)
        entry = {"quad": cd["corners"].copy(), "id": mid, "nrot": nrot, "idx": cd["idx"]}
        if keep:
            entry.update(canon=canon, contour=cd["contour"], bits=bits)
        res["candidates"].append(entry)
        if mid != -1:
            c = cd["corners"].copy()
            if P.corner_method == LINES:
                c = refine_candidate_lines(c, cd["contour"], K, D)
            c = np.roll(c, -((4 - nrot) % 4), axis=0)  # std::rotate(begin, begin+4-nRot, end)  (:364-366)
            decoded.append({"id": mid, "corners": c.astype(np.float32)})
    if decoded and P.corner_method in (HARRIS, SUBPIX):
        C = np.concatenate([m["corners"] for m in decoded]).astype(np.float32)
        if P.locked_corners:
            C = find_corner_maxima(C, grey, int(P.p1))
        if P.corner_method == HARRIS:
            C = harris_refine(grey, C)
        else:
            w = int(P.p1)
            C = cv2.cornerSubPix(grey, C.reshape(-1, 1, 2).copy(), (w, w), (-1, -1),
                                 (cv2.TERM_CRITERIA_MAX_ITER | cv2.TERM_CRITERIA_EPS, 8, 0.005)).reshape(-1, 2)
        for i, m in enumerate(decoded):
            m["corners"] = C[4 * i:4 * i + 4].copy()
    decoded.sort(key=lambda m: m["id"])  # stable (B.8)
    n = len(decoded)
    remove = [False] * n
    for i in range(n - 1):
        if decoded[i]["id"] == decoded[i + 1]["id"] and not remove[i + 1]:
            if perimeter(decoded[i]["corners"]) > perimeter(decoded[i + 1]["corners"]):
                remove[i + 1] = True
            else:
                remove[i] = True
    x0, y0, x1, y1 = valid_region(W, H, P.border_dist)
    for i, m in enumerate(decoded):
        for c in m["corners"]:
            if not np.isfinite(c).all():
                remove[i] = True
                break
            xi, yi = int(np.rint(c[0])), int(np.rint(c[1]))
            if not (x0 <= xi < x1 and y0 <= yi < y1):
                remove[i] = True
                break
    markers = [m for i, m in enumerate(decoded) if not remove[i]]
    if K is not None and marker_size > 0:
        obj = object_points(marker_size)
        for m in markers:
            ok, rvec, tvec = cv2.solvePnP(obj, m["corners"].reshape(4, 1, 2).astype(np.float32), K, D)
            rvec = rvec.reshape(3).astype(np.float64)
            if P.set_y_perpendicular:
                rvec = rotate_x_axis(rvec)
            m["rvec"], m["tvec"], m["ssize"] = rvec, tvec.reshape(3).astype(np.float64), float(marker_size)
    res["markers"] = markers
    return res


def board_detect(markers, board_cfg: dict, K=None, D=None, marker_size: float = -1.0, repj_err_thres: float = -1.0,
                 set_y_perpendicular: bool = False):
    """BoardDetector::detect (src/boarddetector.cpp:90-204) on real OpenCV.  markers: list of {'id','corners'};
    board_cfg: {'mInfoType', 'markers': [{'id','corners' 4x3}]}.  Returns dict(prob, markers, rvec, tvec)."""
    ids = [m["id"] for m in board_cfg["markers"]]
    pts = {m["id"]: np.array(m["corners"], np.float32) for m in board_cfg["markers"]}
    first = np.array(board_cfg["markers"][0]["corners"], np.float32)
    d01 = float(np.sqrt(((first[0].astype(np.float64) - first[1].astype(np.float64)) ** 2).sum()))
    pix = board_cfg["mInfoType"] == 0
    sel = [m for m in markers if m["id"] in pts]
    out = {"prob": 0.0, "markers": sel, "rvec": None, "tvec": None}
    if not sel or K is None:
        return out
    if not ((marker_size > 0 and pix) or not pix):
        return out
    mpp = float(marker_size) / d01 if pix else 1.0
    obj, img = [], []
    for m in sel:
        for p in range(4):
            img.append(np.asarray(m["corners"], np.float32)[p])
            obj.append((pts[m["id"]][p].astype(np.float64) * mpp).astype(np.float32))
    obj, img = np.array(obj, np.float32), np.array(img, np.float32)
    K = np.asarray(K, np.float32).reshape(3, 3)
    Dm = np.zeros((1, 4), np.float32) if D is None else np.asarray(D, np.float32).reshape(1, -1)
    _, rvec, tvec = cv2.solvePnP(obj, img.reshape(-1, 1, 2), K, Dm)
    if repj_err_thres > 0:
        rep = cv2.projectPoints(obj, rvec, tvec, K, Dm)[0].reshape(-1, 2).astype(np.float32)
        err = np.sqrt(((rep - img).astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
        keep = err < np.float32(repj_err_thres)
        _, rvec, tvec = cv2.solvePnP(obj[keep], img[keep].reshape(-1, 1, 2), K, Dm)
    rvec = rvec.reshape(3).astype(np.float64)
    if set_y_perpendicular:
        rvec = rotate_x_axis(rvec)
    out.update(prob=len(sel) / len(ids), rvec=rvec, tvec=tvec.reshape(3).astype(np.float64))
    return out


# ---------------------------------------------------------------------------------------------------------
# Marker / board generators: numpy restatement, pinned by the reference's own PNGs
# (testdata/board/{marker,locked-marker}-expected.png, board.png, chessboard/chessboard.png; tests/golden/render.npz)
# ---------------------------------------------------------------------------------------------------------
def create_marker_image(marker_id, size, locked=False):
    """FiducidalMarkers::createMarkerImage without the watermark (src/arucofidmarkers.cpp:213-263)."""
    assert 0 <= marker_id < 1024
    m = np.zeros((size, size), np.uint8)
    sw = size // 7
    words = [0x10, 0x17, 0x09, 0x0E]
    for y in range(5):
        val = words[(marker_id >> (2 * (4 - y))) & 3]
        for x in range(5):
            if (val >> (4 - x)) & 1:
                m[(y + 1) * sw:(y + 2) * sw, (x + 1) * sw:(x + 2) * sw] = 255
    if locked:
        sq = int(np.float32(size) * np.float32(0.25))
        big = np.full((size + 2 * sq, size + 2 * sq), 255, np.uint8)
        big[:sq, :sq] = 0
        big[big.shape[0] - sq:, :sq] = 0
        big[big.shape[0] - sq:, big.shape[1] - sq:] = 0
        big[:sq, big.shape[1] - sq:] = 0
        big[sq:sq + size, sq:sq + size] = m
        m = big
    return m


def create_board_image(kind, grid_w, grid_h, marker_size, marker_distance, ids, center=True):
    """kind 0: createBoardImage (:283-329, always centred), 1: _ChessBoard (:337-389), 2: _Frame (:397-436).
    Returns (image, ids used, corners [n,4,3] f32)."""
    dist = 0 if kind == 1 else marker_distance
    step = marker_size + dist
    size_y = grid_h * marker_size + (grid_h - 1) * dist
    size_x = grid_w * marker_size + (grid_w - 1) * dist
    cx, cy = size_x // 2, size_y // 2
    img = np.full((size_y, size_x), 255, np.uint8)
    used, corners, k = [], [], 0
    for y in range(grid_h):
        to_write = (y % 2) != 0
        for x in range(grid_w):
            to_write = not to_write
            use = True if kind == 0 else (to_write if kind == 1 else (y in (0, grid_h - 1) or x in (0, grid_w - 1)))
            if not use:
                continue
            img[y * step:y * step + marker_size, x * step:x * step + marker_size] = create_marker_image(ids[k], marker_size)
            x0, y0, s = float(x * step), float(y * step), float(marker_size)
            c = np.array([[x0, y0, 0], [x0 + s, y0, 0], [x0 + s, y0 + s, 0], [x0, y0 + s, 0]], np.float32)
            if kind == 0 or center:
                c -= np.array([cx, cy, 0], np.float32)
            used.append(int(ids[k]))
            corners.append(c)
            k += 1
    return img, used, np.array(corners, np.float32).reshape(-1, 4, 3)


def hrm_marker_image(bits, n, pix_size):
    """MarkerCode::getImg (src/highlyreliablemarkers.cpp:234-256); bits: n*n 0/1 row-major."""
    nrows = n + 2
    if pix_size % nrows != 0:
        pix_size = pix_size + nrows - pix_size % nrows
    cell = pix_size // nrows
    img = np.zeros((pix_size, pix_size), np.uint8)
    for i in range(n):
        for j in range(n):
            if bits[i * n + j]:
                img[(i + 1) * cell:(i + 2) * cell, (j + 1) * cell:(j + 2) * cell] = 255
    return img


def hrm_create_board_image(grid_w, grid_h, codes, n):
    """HighlyReliableMarkers::createBoardImage, non-chromatic (src/highlyreliablemarkers.cpp:498-545).
    codes: list of n*n '0'/'1' strings.  Returns (image, ids = MarkerCode::getId(), corners [m,4,3] f32)."""
    ms = (n + 2) * 20
    md = ms // 5
    size_y = grid_h * ms + (grid_h - 1) * md
    size_x = grid_w * ms + (grid_w - 1) * md
    cx, cy = np.float32(size_x / 2.), np.float32(size_y / 2.)
    img = np.full((size_y, size_x), 255, np.uint8)
    ids, corners, k = [], [], 0
    for y in range(grid_h):
        for x in range(grid_w):
            bits = [c == "1" for c in codes[k]]
            img[y * (ms + md):y * (ms + md) + ms, x * (ms + md):x * (ms + md) + ms] = hrm_marker_image(bits, n, ms)
            ids.append(int(code_rotations(np.array(bits, np.uint8).reshape(n, n))[1][0]))  # MarkerCode::getId(): `2 << pos` fold
            x0, y0, s = np.float32(x * (ms + md)) - cx, np.float32(y * (ms + md)) - cy, np.float32(ms)
            corners.append([[x0, -y0, 0], [x0 + s, -y0, 0], [x0 + s, -(y0 + s), 0], [x0, -(y0 + s), 0]])
            k += 1
    return img, ids, np.array(corners, np.float32)
