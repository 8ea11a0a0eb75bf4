// Stage 4: corner refinement of decoded markers.
//   k_refine_lines : MarkerDetector::refineCandidateLines (src/markerdetector.cpp:931-997) -- one CTA per
//                    decoded candidate, one warp per marker side: least-squares line through the (optionally
//                    undistorted) contour pixels of the side, corners = intersections of adjacent lines,
//                    re-distorted (distortPoints, :141-153).
//   k_refine_subpix: cv::cornerSubPix(win=(p1,p1), zero=(-1,-1), MAX_ITER 8 | EPS 0.005) (:402-405) -- one
//                    warp per corner.
#pragma once
#include "ab_device.cuh"

namespace ab {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fminf(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}

constexpr int LINES_MAX_SWEEPS = 8;  // Jacobi sweeps kept for replay (two columns converge in 1-3)
constexpr int LINES_SIDE_CAP = 512;  // points of a side staged in shared memory (longer sides are re-read from the contour)

// refineCandidateLines for one candidate by one CTA of 128 threads (warp l = side l).  `pts` = the contour in the
// order the reference holds it (`rev`: read it backwards, :622-625), `c` = the 4 corners (integer valued).
// Every side is fitted exactly as interpolate2Dline does (:83-130): cv::solve(DECOMP_SVD) on the m x 2 CV_32F system
// = OpenCV's one-sided Jacobi SVD of the columns (coordinate, 1).  The rotated columns are not stored: a sweep
// re-derives them from the points by replaying the earlier rotations (1-3 of them), so one pass over the side yields
// the squared norms and dot product the next sweep needs.  f64 sums are tree-reduced (the reference adds
// sequentially; they only enter through values rounded to f32).  The first pass stages the side's points (after
// cv::undistortPoints when UNDIST, :957-959) in shared memory; the later passes read them from there.
template <bool UNDIST>
__device__ __forceinline__ void refine_lines_cta(const uint32_t* pts, int n, bool rev, const float* c, const Camera& cam,
                                                 float2 (*s_side)[LINES_SIDE_CAP], int* s_ci, float (*s_line)[3], float2 (*s_rot)[LINES_MAX_SWEEPS],
                                                 float* refined, unsigned int* err) {
    const int t = threadIdx.x, lane = t & 31, l = t >> 5;
    if (t < 4) s_ci[t] = -1;
    __syncthreads();
    uint32_t ck[4];
#pragma unroll
    for (int k = 0; k < 4; k++) ck[k] = (uint32_t)(int)c[2 * k] | ((uint32_t)(int)c[2 * k + 1] << 16);
    // four contour words in flight per thread (the pool was written kernels ago: these are L2 round trips)
    for (int j0 = t; j0 < n; j0 += 4 * blockDim.x) {
        uint32_t p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = j0 + u * blockDim.x;
            p[u] = j < n ? (rev ? pts[n - 1 - j] : pts[j]) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (p[u] == ck[k]) atomicMax(&s_ci[k], j0 + u * (int)blockDim.x);  // last match wins (:935-941)
    }
    __syncthreads();
    const int c0 = s_ci[0], c1 = s_ci[1], c2 = s_ci[2], c3 = s_ci[3];
    if (c0 < 0 || c1 < 0 || c2 < 0 || c3 < 0) return;  // a corner that is not on the contour (cannot happen inside detect)
    bool inverse;
    if (c1 > c0 && (c2 > c1 || c2 < c0)) inverse = false;
    else if (c2 > c1 && c2 < c0) inverse = false;
    else inverse = true;
    const int inc = inverse ? -1 : 1;
    {
        const int start = s_ci[l], end = s_ci[(l + 1) & 3];
        int m = inverse ? (start - end) : (end - start);
        if (m < 0) m += n;
        const int m2 = (m == 1) ? 2 : m;  // 1-point side: add the next corner (:972-976)
        const bool staged = m2 <= LINES_SIDE_CAP;
        float2* side = s_side[l];
        auto fetch = [&](int i, float& x, float& y) {  // point i of the side from the contour
            int j = start + inc * i;                   // start in [0, n), i < m2 <= n + 1
            if (m == 1 && i == 1) j = end;
            if (j >= n) j -= n;
            if (j < 0) j += n;
            uint32_t p = rev ? pts[n - 1 - j] : pts[j];
            x = (float)(p & 0xFFFFu);
            y = (float)(p >> 16);
            if (UNDIST) undistort_point_px(cam, x, y, &x, &y);  // contour points are integer pixels
        };
        auto point = [&](int i, float& x, float& y) {
            if (staged) {
                float2 q = side[i];
                x = q.x;
                y = q.y;
            } else {
                fetch(i, x, y);
            }
        };
        // pass 0: bounding box decides the parametrisation; column sums of both candidates
        float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
        double Sx = 0, Sy = 0, Sxx = 0, Syy = 0;
#pragma unroll 4
        for (int i = lane; i < m2; i += 32) {
            float x, y;
            fetch(i, x, y);
            if (staged) side[i] = make_float2(x, y);
            mnx = fminf(mnx, x);
            mxx = fmaxf(mxx, x);
            mny = fminf(mny, y);
            mxy = fmaxf(mxy, y);
            Sx += (double)x;
            Sy += (double)y;
            Sxx += (double)x * x;
            Syy += (double)y * y;
        }
        __syncwarp();
        mnx = warp_min_f(mnx);
        mxx = warp_max_f(mxx);
        mny = warp_min_f(mny);
        mxy = warp_max_f(mxy);
        const bool along_x = mxx - mnx > mxy - mny;  // y = a x + c (:103-115), else x = b y + c (:116-128)
        double W2[2] = {warp_sum_d(along_x ? Sxx : Syy), (double)m2};
        double p = warp_sum_d(along_x ? Sx : Sy);
        // the rotations done so far (cos, sin) live in shared memory: replaying them is a loop over `ns` entries (ncu r2s: an
        // unrolled, predicated replay of all LINES_MAX_SWEEPS slots made this kernel issue bound at 14 k warp-instructions
        // per marker)
        float2* rot = s_rot[l];
        float Vt[2][2] = {{1.f, 0.f}, {0.f, 1.f}};
        int ns = 0;
        const int max_iter = max(m2, 30);
        for (int iter = 0; iter < max_iter; iter++) {
            if (fabs(p) <= (double)(FLT_EPSILON * 2) * sqrt(W2[0] * W2[1])) break;
            if (ns == LINES_MAX_SWEEPS) {
                if (lane == 0) atomicOr(err, ERR_LINE_FIT);
                break;
            }
            float cc, ss;
            jacobi_rotation_f32(W2[0], W2[1], p, &cc, &ss);
            if (lane == 0) rot[ns] = make_float2(cc, ss);
            __syncwarp();
            ns++;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float v0 = cc * Vt[0][k] + ss * Vt[1][k], v1 = -ss * Vt[0][k] + cc * Vt[1][k];
                Vt[0][k] = v0;
                Vt[1][k] = v1;
            }
            double a = 0, bb = 0, pn = 0;
            for (int i = lane; i < m2; i += 32) {
                float x, y;
                point(i, x, y);
                float t0 = along_x ? x : y, t1 = 1.f;
#pragma unroll 1
                for (int k = 0; k < ns; k++) {
                    const float2 cs2 = rot[k];
                    float n0 = cs2.x * t0 + cs2.y * t1, n1 = -cs2.y * t0 + cs2.x * t1;
                    t0 = n0;
                    t1 = n1;
                }
                a += (double)t0 * t0;
                bb += (double)t1 * t1;
                pn += (double)t0 * t1;
            }
            W2[0] = warp_sum_d(a);
            W2[1] = warp_sum_d(bb);
            p = warp_sum_d(pn);
        }
        float w[2], sc[2], X[2];
        const int o0 = jacobi_scales_f32(W2, w, sc);
        double ub[2] = {0, 0};
        for (int i = lane; i < m2; i += 32) {
            float x, y;
            point(i, x, y);
            float t0 = along_x ? x : y, t1 = 1.f;
            const float rhs = along_x ? y : x;
#pragma unroll 1
            for (int k = 0; k < ns; k++) {
                const float2 cs2 = rot[k];
                float n0 = cs2.x * t0 + cs2.y * t1, n1 = -cs2.y * t0 + cs2.x * t1;
                t0 = n0;
                t1 = n1;
            }
            ub[0] += (double)((t0 * sc[0]) * rhs);
            ub[1] += (double)((t1 * sc[1]) * rhs);
        }
        ub[0] = warp_sum_d(ub[0]);
        ub[1] = warp_sum_d(ub[1]);
        jacobi_backsubst_f32(o0, w, ub, Vt, X);
        if (lane == 0) {
            s_line[l][0] = along_x ? X[0] : -1.f;
            s_line[l][1] = along_x ? -1.f : X[0];
            s_line[l][2] = X[1];
        }
    }
    __syncthreads();
    if (t < 4) {
        // getCrossPoint(lines[i], lines[i-1]) (:132-139, :985-987)
        float x, y;
        cross_point_f32(s_line[t], s_line[(t + 3) & 3], &x, &y);
        if (cam.has_K && cam.has_D) {
            // distortPoints (:141-153): normalise in f32, project with R=t=0 in f64
            float xn = (x - cam.cxf) / cam.fxf, yn = (y - cam.cyf) / cam.fyf;
            double u, v;
            distort_norm_to_px(cam, (double)xn, (double)yn, &u, &v);
            x = (float)u;
            y = (float)v;
        }
        refined[2 * t] = x;
        refined[2 * t + 1] = y;
    }
}

// grid = (LINES_CTAS_PER_FRAME, frames): a CTA walks the candidate list of its frame with that stride, so every CTA of the
// grid has decoded candidates to work on (a CTA per candidate SLOT launched 512 CTAs per frame for ~100 markers)
constexpr int LINES_CTAS_PER_FRAME = 64;
template <bool UNDIST>
__global__ void __launch_bounds__(128) k_refine_lines(Batch b) {
    __shared__ float2 s_side[4][LINES_SIDE_CAP];
    __shared__ float2 s_rot[4][LINES_MAX_SWEEPS];
    __shared__ int s_ci[4];
    __shared__ float s_line[4][3];
    const int f = blockIdx.y, nc = min((int)b.n_cands[f], b.cap_c);
    for (int ci = blockIdx.x; ci < nc; ci += gridDim.x) {
        CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
        if (cand->id < 0) continue;
        const ContourRec rec = b.contours[cand->contour];
        // the reference reverses the contour of swapped candidates (:622-625)
        refine_lines_cta<UNDIST>(b.pool + rec.off, (int)rec.n, cand->swapped != 0, cand->c, b.cam, s_side, s_ci, s_line, s_rot, cand->refined,
                                 &b.cnt->err);
        __syncthreads();  // s_ci / s_line / s_side are reused by the next candidate
    }
}

// public worker MarkerDetector::refineCandidateLines (h:280): one candidate, contour given by the caller
template <bool UNDIST>
__global__ void __launch_bounds__(128) k_refine_lines_single(const uint32_t* pts, int n, const float* corners, Camera cam, float* out,
                                                             unsigned int* err) {
    __shared__ float2 s_side[4][LINES_SIDE_CAP];
    __shared__ float2 s_rot[4][LINES_MAX_SWEEPS];
    __shared__ int s_ci[4];
    __shared__ float s_line[4][3];
    refine_lines_cta<UNDIST>(pts, n, false, corners, cam, s_side, s_ci, s_line, s_rot, out, err);
}

// bilinear getRectSubPix sample with replicated border, f32 arithmetic in OpenCV's order
__device__ __forceinline__ float subpix_sample(const uint8_t* img, size_t row, int W, int H, int ix, int iy,
                                               float a11, float a12, float a21, float a22) {
    int x0 = min(max(ix, 0), W - 1), x1 = min(max(ix + 1, 0), W - 1);
    int y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
    float p00 = img[(size_t)y0 * row + x0], p01 = img[(size_t)y0 * row + x1];
    float p10 = img[(size_t)y1 * row + x0], p11 = img[(size_t)y1 * row + x1];
    return p00 * a11 + p01 * a12 + p10 * a21 + p11 * a22;
}

// grid.x covers 4*cap_c corners of frame blockIdx.y in groups of (blockDim.x/32) warps
__global__ void __launch_bounds__(128) k_refine_subpix(Batch b) {
    extern __shared__ float s_patch[];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int corner = blockIdx.x * (blockDim.x >> 5) + wib;
    const int ci = corner >> 2, k = corner & 3;
    if (ci >= (int)b.n_cands[f]) return;
    CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
    if (cand->id < 0) return;
    const int w = b.subpix_win, win = 2 * w + 1, pw = win + 2;
    float* patch = s_patch + (size_t)wib * pw * pw;
    const uint8_t* img = b.grey + (size_t)f * b.grey_frame;
    const float cTx = cand->refined[2 * k], cTy = cand->refined[2 * k + 1];
    float cIx = cTx, cIy = cTy;
    if (!(cIx >= 0 && cIx < b.W && cIy >= 0 && cIy < b.H)) return;  // the reference would CV_Assert
    const double eps2 = 0.005 * 0.005;
    const float fw = (float)w;
    for (int iter = 0; iter < 8; iter++) {
        float cx = cIx - (float)(pw - 1) * 0.5f, cy = cIy - (float)(pw - 1) * 0.5f;
        int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
        float a = cx - (float)ipx, bfr = cy - (float)ipy;
        float a11 = (1.f - a) * (1.f - bfr), a12 = a * (1.f - bfr), a21 = (1.f - a) * bfr, a22 = a * bfr;
        for (int i = lane; i < pw * pw; i += 32) {
            int py = i / pw, px = i - py * pw;
            patch[i] = subpix_sample(img, b.grey_row, b.W, b.H, ipx + px, ipy + py, a11, a12, a21, a22);
        }
        __syncwarp();
        double A = 0, Bq = 0, C = 0, bb1 = 0, bb2 = 0;
        for (int i = lane; i < win * win; i += 32) {
            int yy = i / win, xx = i - yy * win;
            float my = (float)(yy - w) / fw, mx = (float)(xx - w) / fw;
            float mk = expf(-my * my) * expf(-mx * mx);
            const float* sp = patch + (yy + 1) * pw + (xx + 1);
            double m = mk;
            double tgx = (double)(sp[1] - sp[-1]);
            double tgy = (double)(sp[pw] - sp[-pw]);
            double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
            double px = xx - w, py = yy - w;
            A += gxx;
            Bq += gxy;
            C += gyy;
            bb1 += gxx * px + gxy * py;
            bb2 += gxy * px + gyy * py;
        }
        A = warp_sum_d(A);
        Bq = warp_sum_d(Bq);
        C = warp_sum_d(C);
        bb1 = warp_sum_d(bb1);
        bb2 = warp_sum_d(bb2);
        __syncwarp();
        double det = A * C - Bq * Bq;
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        double scale = 1.0 / det;
        float nx = (float)((double)cIx + C * scale * bb1 - Bq * scale * bb2);
        float ny = (float)((double)cIy - Bq * scale * bb1 + A * scale * bb2);
        double err = (double)((nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy));
        cIx = nx;
        cIy = ny;
        if (cIx < 0 || cIx >= b.W || cIy < 0 || cIy >= b.H) break;
        if (!(err > eps2)) break;
    }
    if (fabsf(cIx - cTx) > fw || fabsf(cIy - cTy) > fw) {
        cIx = cTx;
        cIy = cTy;
    }
    if (lane == 0) {
        cand->refined[2 * k] = cIx;
        cand->refined[2 * k + 1] = cIy;
    }
}

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// findCornerMaxima ("locked corners", src/markerdetector.cpp:157-199): one CTA per corner.
// cornerHarris(blockSize 3, ksize 3, k 0.04) over the +-wsize region around the corner (the Sobel of the C++
// ROI reads the parent image, the 3x3 box sum reflects at the ROI border), 4x4 block sums over the interior,
// arg-max of the centre-weighted response (first strict maximum in raster order).
__global__ void __launch_bounds__(64) k_corner_maxima(Batch b) {
    extern __shared__ float s_buf[];
    __shared__ unsigned long long s_best;
    const int f = blockIdx.y, corner = blockIdx.x, t = threadIdx.x;
    const int ci = corner >> 2, k = corner & 3;
    if (ci >= (int)b.n_cands[f]) return;
    CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
    if (cand->id < 0) return;
    const int wsize = b.subpix_win;
    const float cxf = cand->refined[2 * k], cyf = cand->refined[2 * k + 1];
    const int x0 = max(0, (int)(cxf - wsize)), y0 = max(0, (int)(cyf - wsize));
    const int x1 = min(b.W, (int)(cxf + wsize)), y1 = min(b.H, (int)(cyf + wsize));
    const int rw = x1 - x0, rh = y1 - y0;
    if (rw <= 0 || rh <= 0) {
        if (t == 0) {
            cand->refined[2 * k] = -1.f + x0;
            cand->refined[2 * k + 1] = -1.f + y0;
        }
        return;
    }
    const int n = rw * rh;
    float *ca = s_buf, *cb = ca + n, *cc = cb + n, *harr = cc + n, *hs = harr + n;
    const uint8_t* img = b.grey + (size_t)f * b.grey_frame;
    const float scale = (float)(1.0 / (4.0 * 3.0 * 255.0)), c2 = (float)(2.0 * (1.0 / (4.0 * 3.0 * 255.0)));
    if (t == 0) s_best = 0ull;
    for (int i = t; i < n; i += blockDim.x) {
        int y = i / rw, x = i - y * rw, gx = x0 + x, gy = y0 + y;
        int xm = reflect101(gx - 1, b.W), xc = gx, xp = reflect101(gx + 1, b.W);
        const uint8_t* rm = img + (size_t)reflect101(gy - 1, b.H) * b.grey_row;
        const uint8_t* r0 = img + (size_t)gy * b.grey_row;
        const uint8_t* rp = img + (size_t)reflect101(gy + 1, b.H) * b.grey_row;
        float d_m = (float)rm[xp] - (float)rm[xm], d_0 = (float)r0[xp] - (float)r0[xm], d_p = (float)rp[xp] - (float)rp[xm];
        float dx = c2 * d_0 + scale * (d_m + d_p);
        float s_m = c2 * (float)rm[xc] + scale * ((float)rm[xm] + (float)rm[xp]);
        float s_p = c2 * (float)rp[xc] + scale * ((float)rp[xm] + (float)rp[xp]);
        float dy = s_p - s_m;
        ca[i] = dx * dx;
        cb[i] = dx * dy;
        cc[i] = dy * dy;
    }
    __syncthreads();
    for (int i = t; i < n; i += blockDim.x) {
        int y = i / rw, x = i - y * rw;
        double sa = 0, sb = 0, sc = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int j = reflect101(y + dy, rh) * rw + reflect101(x + dx, rw);
                sa += ca[j];
                sb += cb[j];
                sc += cc[j];
            }
        float a = (float)sa, bq = (float)sb, c = (float)sc;
        harr[i] = (float)((double)(a * c - bq * bq) - 0.04 * (double)(a + c) * (double)(a + c));
    }
    __syncthreads();
    unsigned long long best = 0ull;  // (float bits of the positive response << 32) | ~raster index
    const float ccx = (float)(rw / 2), ccy = (float)(rh / 2), den = (float)(rw / 2 + rh / 2);
    for (int i = t; i < n; i += blockDim.x) {
        int y = i / rw, x = i - y * rw;
        float h = harr[i];
        if (y >= 4 && y < rh - 4 && x >= 4 && x < rw - 4) {
            double sum = 0;
            for (int dy = 0; dy < 4; dy++)
                for (int dx = 0; dx < 4; dx++) sum += harr[(y + dy) * rw + x + dx];
            h = (float)sum;
        }
        hs[i] = h;
        float d = (float)(fabs((double)(ccx - (float)x)) + fabs((double)(ccy - (float)y))) / den;
        float w = (float)(1. - (double)d);
        float v = w * h;
        if (v > 0.f) {
            unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
            if (key > best) best = key;
        }
    }
    if (best) atomicMax(&s_best, best);
    __syncthreads();
    if (t == 0) {
        float bx = -1.f, by = -1.f;
        if (s_best) {
            int i = (int)(0xFFFFFFFFu - (unsigned)(s_best & 0xFFFFFFFFull));
            by = (float)(i / rw);
            bx = (float)(i - (i / rw) * rw);
        }
        cand->refined[2 * k] = bx + (float)x0;
        cand->refined[2 * k + 1] = by + (float)y0;
    }
}

// HARRIS mode: SubPixelCorner::RefineCorner (src/subpixelcorner.cpp:70-189), one warp per corner.  Reproduces
// the reference's quirks: exactly one iteration, `D` (never assigned) in the y update, 8-bit
// getRectSubPix patch (16-bit fixed-point bilinear weights) before the 3x3 Sobel.
__global__ void __launch_bounds__(128) k_refine_harris(Batch b) {
    __shared__ int s_patch[4][17 * 17];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int corner = blockIdx.x * (blockDim.x >> 5) + wib;
    const int ci = corner >> 2, k = corner & 3;
    if (ci >= (int)b.n_cands[f]) return;
    CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
    if (cand->id < 0) return;
    const int win = 15, S = 17;
    const float ex = cand->refined[2 * k], ey = cand->refined[2 * k + 1];
    if (ex < 0 || ey < 0 || ey > b.H || ey > b.W) return;  // (:87-89, the reference compares y with the width too)
    int* patch = s_patch[wib];
    const uint8_t* img = b.grey + (size_t)f * b.grey_frame;
    {
        float cx = ex - (float)(S - 1) * 0.5f, cy = ey - (float)(S - 1) * 0.5f;
        int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
        float a = cx - (float)ipx, bb = cy - (float)ipy;
        int a11 = __float2int_rn((1.f - a) * (1.f - bb) * 65536.f), a12 = __float2int_rn(a * (1.f - bb) * 65536.f);
        int a21 = __float2int_rn((1.f - a) * bb * 65536.f), a22 = __float2int_rn(a * bb * 65536.f);
        for (int i = lane; i < S * S; i += 32) {
            int py = i / S, px = i - py * S;
            int xx0 = min(max(ipx + px, 0), b.W - 1), xx1 = min(max(ipx + px + 1, 0), b.W - 1);
            int yy0 = min(max(ipy + py, 0), b.H - 1), yy1 = min(max(ipy + py + 1, 0), b.H - 1);
            int v = (int)img[(size_t)yy0 * b.grey_row + xx0] * a11 + (int)img[(size_t)yy0 * b.grey_row + xx1] * a12 +
                    (int)img[(size_t)yy1 * b.grey_row + xx0] * a21 + (int)img[(size_t)yy1 * b.grey_row + xx1] * a22;
            patch[i] = (v + (1 << 15)) >> 16;
        }
    }
    __syncwarp();
    const double coeff = 1. / (win * win);
    double A = 0, B = 0, C = 0, E = 0, F = 0;
    for (int i = lane; i < win * win; i += 32) {
        int yy = i / win, xx = i - yy * win;  // window position 0..14  (patch index +1)
        int ly = yy - win / 2, lx = xx - win / 2;
        const int* q = patch + (yy + 1) * S + (xx + 1);
        float dx = (float)((q[-S + 1] - q[-S - 1]) + 2 * (q[1] - q[-1]) + (q[S + 1] - q[S - 1]));
        float dy = (float)((q[S - 1] - q[-S - 1]) + 2 * (q[S] - q[-S]) + (q[S + 1] - q[-S + 1]));
        float mxv = (float)exp(-(double)(lx * lx) * coeff), myv = (float)exp(-(double)(ly * ly) * coeff);
        double val = (double)(mxv * myv);
        double dxx = (double)(dx * dx) * val, dyy = (double)(dy * dy) * val, dxy = (double)(dx * dy) * val;
        A += dxx;
        B += dxy;
        E += dyy;
        C += dxx * lx + dxy * ly;
        F += dxy * lx + dyy * ly;
    }
    A = warp_sum_d(A);
    B = warp_sum_d(B);
    C = warp_sum_d(C);
    E = warp_sum_d(E);
    F = warp_sum_d(F);
    if (lane == 0) {
        const double D = 0;  // never assigned in the reference
        double det = A * E - B * B;
        float nx = ex, ny = ey;
        if (fabs(det) > DBL_EPSILON * DBL_EPSILON) {
            det = 1.0 / det;
            nx = (float)((double)ex + ((C * E) - (B * F)) * det);
            ny = (float)((double)ey + ((A * F) - (C * D)) * det);
        }
        if (fabs((double)ex - (double)nx) > win || fabs((double)ey - (double)ny) > win) {
            nx = ex;
            ny = ey;
        }
        cand->refined[2 * k] = nx;
        cand->refined[2 * k + 1] = ny;
    }
}

}  // namespace ab
