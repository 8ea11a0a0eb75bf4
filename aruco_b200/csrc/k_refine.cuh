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

__global__ void __launch_bounds__(128) k_refine_lines(Batch b) {
    __shared__ int s_ci[4];
    __shared__ float s_line[4][3];
    const int f = blockIdx.y, ci = blockIdx.x, t = threadIdx.x, lane = t & 31, l = t >> 5;
    if (ci >= (int)b.n_cands[f]) return;
    CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
    if (cand->id < 0) return;
    const ContourRec rec = b.contours[cand->contour];
    const int n = (int)rec.n;
    const uint32_t* pts = b.pool + rec.off;
    const bool rev = cand->swapped != 0;  // the reference reverses the contour of swapped candidates (:622-625)
    if (t < 4) s_ci[t] = -1;
    __syncthreads();
    uint32_t ck[4];
#pragma unroll
    for (int k = 0; k < 4; k++) ck[k] = (uint32_t)(int)cand->c[2 * k] | ((uint32_t)(int)cand->c[2 * k + 1] << 16);
    for (int j = t; j < n; j += blockDim.x) {
        uint32_t p = rev ? pts[n - 1 - j] : pts[j];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (p == ck[k]) atomicMax(&s_ci[k], j);  // last match wins (:935-941)
    }
    __syncthreads();
    const int c0 = s_ci[0], c1 = s_ci[1], c2 = s_ci[2], c3 = s_ci[3];
    if (c0 < 0 || c1 < 0 || c2 < 0 || c3 < 0) return;  // cannot happen for approxPolyDP vertices
    bool inverse;
    if (c1 > c0 && (c2 > c1 || c2 < c0)) inverse = false;
    else if (c2 > c1 && c2 < c0) inverse = false;
    else inverse = true;
    const int inc = inverse ? -1 : 1;
    const bool undist = b.cam.has_K && b.cam.has_D;
    {
        const int start = s_ci[l], end = s_ci[(l + 1) & 3];
        int m = inverse ? (start - end) : (end - start);
        if (m < 0) m += n;
        int m2 = (m == 1) ? 2 : m;  // 1-point side: add the next corner (:972-976)
        float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
        double Sx = 0, Sy = 0, Sxx = 0, Syy = 0, Sxy = 0;
        for (int i = lane; i < m2; i += 32) {
            int j = start + inc * i;
            if (m == 1 && i == 1) j = end;
            j %= n;
            if (j < 0) j += n;
            uint32_t p = rev ? pts[n - 1 - j] : pts[j];
            float x = (float)(p & 0xFFFFu), y = (float)(p >> 16);
            if (undist) undistort_point_px(b.cam, x, y, &x, &y);
            mnx = fminf(mnx, x);
            mxx = fmaxf(mxx, x);
            mny = fminf(mny, y);
            mxy = fmaxf(mxy, y);
            double dx = x, dy = y;
            Sx += dx;
            Sy += dy;
            Sxx += dx * dx;
            Syy += dy * dy;
            Sxy += dx * dy;
        }
        mnx = warp_min_f(mnx);
        mxx = warp_max_f(mxx);
        mny = warp_min_f(mny);
        mxy = warp_max_f(mxy);
        Sx = warp_sum_d(Sx);
        Sy = warp_sum_d(Sy);
        Sxx = warp_sum_d(Sxx);
        Syy = warp_sum_d(Syy);
        Sxy = warp_sum_d(Sxy);
        if (lane == 0) {
            double N = (double)m2;
            if (mxx - mnx > mxy - mny) {  // y = a x + c  (interpolate2Dline, :103-115)
                double a = (N * Sxy - Sx * Sy) / (N * Sxx - Sx * Sx);
                double c = (Sy - a * Sx) / N;
                s_line[l][0] = (float)a;
                s_line[l][1] = -1.f;
                s_line[l][2] = (float)c;
            } else {  // x = b y + c
                double bb = (N * Sxy - Sx * Sy) / (N * Syy - Sy * Sy);
                double c = (Sx - bb * Sy) / N;
                s_line[l][0] = -1.f;
                s_line[l][1] = (float)bb;
                s_line[l][2] = (float)c;
            }
        }
    }
    __syncthreads();
    if (t < 4) {
        // getCrossPoint(lines[i], lines[i-1]) (:132-139, :985-987)
        const float* l1 = s_line[t];
        const float* l2 = s_line[(t + 3) & 3];
        double a11 = l1[0], a12 = l1[1], a21 = l2[0], a22 = l2[1], b1 = -(double)l1[2], b2 = -(double)l2[2];
        double det = a11 * a22 - a12 * a21;
        float x = (float)((b1 * a22 - a12 * b2) / det);
        float y = (float)((a11 * b2 - b1 * a21) / det);
        if (undist) {
            // distortPoints (:141-153): normalise in f32, project with R=t=0 in f64
            float xn = (x - b.cam.cxf) / b.cam.fxf, yn = (y - b.cam.cyf) / b.cam.fyf;
            double u, v;
            distort_norm_to_px(b.cam, (double)xn, (double)yn, &u, &v);
            x = (float)u;
            y = (float)v;
        }
        cand->refined[2 * t] = x;
        cand->refined[2 * t + 1] = y;
    }
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

}  // namespace ab
