// CPU ORACLE -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.
//
// Dependency-free, sequential C++ restatement of the reference's hot path aruco::MarkerDetector::detect
// (paroj/aruco src/markerdetector.cpp:302-478) including the OpenCV primitives it delegates to.  The
// arithmetic of the path lives in the third-party dependency OpenCV, which is neither vendored nor pinned
// by the reference (CMakeLists.txt:50 `FIND_PACKAGE(OpenCV REQUIRED)`, README.md:88 ">= 2.4.9"); the
// primitives are restated from their published algorithms with the semantics of OpenCV 4.13.0 (the only
// OpenCV in this image, python wheel) -- SURVEY.md Appendix A.
//
// Parity pin: tests/test_oracle_golden.py checks this oracle against the reference's own golden files
// testdata/{single,hrm,board,chessboard}/expected.yml (committed as tests/golden/expected.json) and, where
// cv2 is importable, primitive by primitive and end to end against oracle/cv2_oracle.py (real OpenCV).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
// OpenMP is used exactly where the reference uses it (src/markerdetector.cpp:456,587 via src/ar_omp.h) plus
// one frame-parallel loop in orc_detect_batch for the CPU baseline.
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Pt { int x, y; };
struct Pt2f { float x, y; };

// ------------------------------------------------------------------------------------------------
// thresHold (src/markerdetector.cpp:643-677)
// ------------------------------------------------------------------------------------------------
// cv::adaptiveThreshold(MEAN_C, BINARY_INV): box mean over k x k (replicate border) rounded to nearest,
// dst = (src - mean <= -floor(C)) ? 255 : 0
void adaptive_threshold(const uint8_t* src, int W, int H, int k, double C, uint8_t* dst) {
    const int r = k / 2, k2 = k * k;
    const int idelta = (int)floor(C);
    std::vector<int> integ((size_t)(W + k) * (H + k), 0);  // integral of the replicate-padded image
    const int PW = W + 2 * r, PH = H + 2 * r, IW = PW + 1;
    for (int y = 0; y < PH; y++) {
        int sy = std::min(std::max(y - r, 0), H - 1);
        int rowsum = 0;
        for (int x = 0; x < PW; x++) {
            int sx = std::min(std::max(x - r, 0), W - 1);
            rowsum += src[(size_t)sy * W + sx];
            integ[(size_t)(y + 1) * IW + x + 1] = integ[(size_t)y * IW + x + 1] + rowsum;
        }
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int S = integ[(size_t)(y + k) * IW + x + k] - integ[(size_t)y * IW + x + k] - integ[(size_t)(y + k) * IW + x] +
                    integ[(size_t)y * IW + x];
            int mean = (2 * S + k2) / (2 * k2);
            dst[(size_t)y * W + x] = ((int)src[(size_t)y * W + x] - mean <= -idelta) ? 255 : 0;
        }
}

// cv::Canny(src, dst, low, high), aperture 3, L1 gradient (src/markerdetector.cpp:669): Sobel with replicated border,
// non-maximum suppression with OpenCV's fixed-point tangent tests, hysteresis with an explicit stack.
void canny(const uint8_t* src, int W, int H, int low, int high, uint8_t* dst) {
    std::vector<int> dxv((size_t)W * H), dyv((size_t)W * H), mag((size_t)(W + 2) * (H + 2), 0);
    auto P = [&](int x, int y) { return (int)src[(size_t)std::min(std::max(y, 0), H - 1) * W + std::min(std::max(x, 0), W - 1)]; };
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int dx = (P(x + 1, y - 1) - P(x - 1, y - 1)) + 2 * (P(x + 1, y) - P(x - 1, y)) + (P(x + 1, y + 1) - P(x - 1, y + 1));
            int dy = (P(x - 1, y + 1) - P(x - 1, y - 1)) + 2 * (P(x, y + 1) - P(x, y - 1)) + (P(x + 1, y + 1) - P(x + 1, y - 1));
            dxv[(size_t)y * W + x] = dx;
            dyv[(size_t)y * W + x] = dy;
            mag[(size_t)(y + 1) * (W + 2) + x + 1] = abs(dx) + abs(dy);
        }
    std::vector<uint8_t> map((size_t)(W + 2) * (H + 2), 0);  // 0 none, 1 candidate, 2 edge
    std::vector<int> stack;
    const int MS = W + 2;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int* mp = &mag[(size_t)(y + 1) * MS + x + 1];
            int m = mp[0];
            if (m <= low) continue;
            int xs = dxv[(size_t)y * W + x], ys = dyv[(size_t)y * W + x];
            long long ax = abs(xs), ay = (long long)abs(ys) << 15, tg22x = ax * 13573;
            bool keep;
            if (ay < tg22x) keep = m > mp[-1] && m >= mp[1];
            else {
                long long tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = m > mp[-MS] && m >= mp[MS];
                else {
                    int s = (xs ^ ys) < 0 ? -1 : 1;
                    keep = m > mp[-MS - s] && m > mp[MS + s];
                }
            }
            if (!keep) continue;
            int idx = (y + 1) * MS + x + 1;
            if (m > high) {
                map[idx] = 2;
                stack.push_back(idx);
            } else
                map[idx] = 1;
        }
    const int nb[8] = {-MS - 1, -MS, -MS + 1, -1, 1, MS - 1, MS, MS + 1};
    while (!stack.empty()) {
        int i = stack.back();
        stack.pop_back();
        for (int k = 0; k < 8; k++)
            if (map[i + nb[k]] == 1) {
                map[i + nb[k]] = 2;
                stack.push_back(i + nb[k]);
            }
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) dst[(size_t)y * W + x] = map[(size_t)(y + 1) * MS + x + 1] == 2 ? 255 : 0;
}

void erode3x3(const uint8_t* src, int W, int H, uint8_t* dst) {  // cv::erode(src, dst, Mat()): border = +inf
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t m = 255;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    m = std::min(m, src[(size_t)yy * W + xx]);
                }
            dst[(size_t)y * W + x] = m;
        }
}

// ------------------------------------------------------------------------------------------------
// cv::findContours(RETR_LIST, CHAIN_APPROX_NONE): Suzuki-Abe border following as OpenCV implements it
// (serial raster scan with +-NBD marks), SURVEY A.2.  Contours are returned in OpenCV's order.
// ------------------------------------------------------------------------------------------------
void find_contours(const uint8_t* bin, int W, int H, std::vector<std::vector<Pt>>& out) {
    const int PW = W + 2;
    std::vector<int8_t> img((size_t)PW * (H + 2), 0);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) img[(size_t)(y + 1) * PW + x + 1] = bin[(size_t)y * W + x] ? 1 : 0;
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int off[16];
    for (int i = 0; i < 16; i++) off[i] = dy[i & 7] * PW + dx[i & 7];
    const int NBD = 2;
    std::vector<std::vector<Pt>> found;
    for (int y = 1; y <= H; y++) {
        for (int x = 1; x <= W + 1; x++) {
            int p = img[(size_t)y * PW + x], prev = img[(size_t)y * PW + x - 1];
            if (p == prev) continue;
            bool is_hole;
            int sx;
            if (prev == 0 && p == 1) {
                is_hole = false;
                sx = x;
            } else if (p == 0 && prev >= 1) {
                is_hole = true;
                sx = x - 1;
            } else
                continue;
            std::vector<Pt> c;
            int i0 = y * PW + sx;
            int s_end = is_hole ? 0 : 4, s = s_end, i1;
            do {
                s = (s - 1) & 7;
                i1 = i0 + off[s];
            } while (img[i1] == 0 && s != s_end);
            if (img[i1] == 0) {  // isolated pixel (s came back to s_end)
                img[i0] = -NBD;
                c.push_back(Pt{i0 % PW - 1, i0 / PW - 1});
            } else {
                int i3 = i0;
                for (;;) {
                    int s_e = s, i4;
                    do {
                        s++;
                        i4 = i3 + off[s & 15 & 7];
                    } while (img[i4] == 0);
                    s &= 7;
                    if ((unsigned)(s - 1) < (unsigned)s_e) img[i3] = -NBD;
                    else if (img[i3] == 1) img[i3] = NBD;
                    c.push_back(Pt{i3 % PW - 1, i3 / PW - 1});
                    if (i4 == i0 && i3 == i1) break;
                    i3 = i4;
                    s = (s + 4) & 7;
                }
            }
            found.push_back(std::move(c));
        }
    }
    out.assign(found.rbegin(), found.rend());
}

// ------------------------------------------------------------------------------------------------
// cv::approxPolyDP(closed) -- OpenCV 4.13 (segment distance), SURVEY A.3
// ------------------------------------------------------------------------------------------------
double seg_d2(Pt p, Pt s, Pt e) {
    double dx = e.x - s.x, dy = e.y - s.y, qx = p.x - s.x, qy = p.y - s.y;
    double dd = dx * dx + dy * dy, t = qx * dx + qy * dy;
    if (t < 0) return qx * qx + qy * qy;
    if (t > dd) {
        double fx = p.x - e.x, fy = p.y - e.y;
        return fx * fx + fy * fy;
    }
    double c = qx * dy - qy * dx;
    return c * c / dd;
}

void approx_poly_dp(const std::vector<Pt>& src, double eps, std::vector<Pt>& dst) {
    dst.clear();
    const int count = (int)src.size();
    if (count == 0) return;
    const double E = eps * eps;
    std::vector<std::pair<int, int>> stack;
    int rs = 0, pos = 0;
    bool le = false;
    Pt start_pt{0, 0};
    for (int it = 0; it < 3; it++) {
        double max_dist = 0;
        pos = (pos + rs) % count;
        start_pt = src[pos];
        if (++pos >= count) pos = 0;
        for (int j = 1; j < count; j++) {
            Pt pt = src[pos];
            if (++pos >= count) pos = 0;
            double ddx = pt.x - start_pt.x, ddy = pt.y - start_pt.y, dist = ddx * ddx + ddy * ddy;
            if (dist > max_dist) {
                max_dist = dist;
                rs = j;
            }
        }
        le = max_dist <= E;
    }
    if (!le) {
        int A = pos % count, B = (rs + A) % count;
        stack.push_back({B, A});
        stack.push_back({A, B});
    } else
        dst.push_back(start_pt);
    while (!stack.empty()) {
        int s = stack.back().first, e = stack.back().second;
        stack.pop_back();
        Pt end_pt = src[e];
        pos = s;
        start_pt = src[pos];
        if (++pos >= count) pos = 0;
        int split = 0;
        if (pos != e) {
            double max_dist = 0;
            while (pos != e) {
                Pt pt = src[pos];
                if (++pos >= count) pos = 0;
                double d = seg_d2(pt, start_pt, end_pt);
                if (d > max_dist) {
                    max_dist = d;
                    split = (pos + count - 1) % count;
                }
            }
            le = max_dist <= E;
        } else {
            le = true;
            start_pt = src[s];
        }
        if (le) dst.push_back(start_pt);
        else {
            stack.push_back({split, e});
            stack.push_back({s, split});
        }
    }
    // final clean-up
    int cnt = (int)dst.size(), new_count = cnt;
    pos = cnt - 1;
    start_pt = dst[pos];
    if (++pos >= cnt) pos = 0;
    int wpos = pos;
    Pt pt = dst[pos];
    if (++pos >= cnt) pos = 0;
    for (int i = 0; i < cnt && new_count > 2; i++) {
        Pt end_pt = dst[pos];
        if (++pos >= cnt) pos = 0;
        double dx = end_pt.x - start_pt.x, dy = end_pt.y - start_pt.y;
        double dist = fabs((pt.x - start_pt.x) * dy - (pt.y - start_pt.y) * dx);
        double sip = (double)((pt.x - start_pt.x) * (end_pt.x - pt.x) + (pt.y - start_pt.y) * (end_pt.y - pt.y));
        if (dist * dist <= 0.5 * E * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
            new_count--;
            dst[wpos] = start_pt = end_pt;
            if (++wpos >= cnt) wpos = 0;
            pt = dst[pos];
            if (++pos >= cnt) pos = 0;
            i++;
            continue;
        }
        dst[wpos] = start_pt = pt;
        if (++wpos >= cnt) wpos = 0;
        pt = end_pt;
    }
    dst.resize(new_count);
}

bool is_contour_convex(const std::vector<Pt>& p) {
    int n = (int)p.size();
    Pt prev = p[(n - 2 + n) % n], cur = p[n - 1];
    int dx0 = cur.x - prev.x, dy0 = cur.y - prev.y, orientation = 0;
    for (int i = 0; i < n; i++) {
        prev = cur;
        cur = p[i];
        int dx = cur.x - prev.x, dy = cur.y - prev.y;
        long long dxdy0 = (long long)dx * dy0, dydx0 = (long long)dy * dx0;
        orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
        if (orientation == 3) return false;
        dx0 = dx;
        dy0 = dy;
    }
    return true;
}

float perimeter(const Pt2f* a) {  // src/utils.h:37-44
    float sum = 0;
    for (int i = 0; i < 4; i++) {
        int i2 = (i + 1) % 4;
        float dx = a[i].x - a[i2].x, dy = a[i].y - a[i2].y;
        sum += sqrt((double)dx * dx + (double)dy * dy);
    }
    return sum;
}

struct Candidate {
    Pt2f c[4];
    std::vector<Pt> contour;
    int idx;
    int id, nrot;
};

// detectRectangles (src/markerdetector.cpp:496-635)
void detect_rectangles(const std::vector<const uint8_t*>& thres_images, int W, int H, float min_size, float max_size,
                       std::vector<Candidate>& out, int* n_contours) {
    int minSize = min_size * std::max(W, H) * 4;
    int maxSize = max_size * std::max(W, H) * 4;
    std::vector<Candidate> cands;
    std::vector<Pt> approx;
    if (n_contours) *n_contours = 0;
    for (size_t ti = 0; ti < thres_images.size(); ti++) {  // one pass per threshold image (:506-560), joined in order
    std::vector<std::vector<Pt>> contours;
    find_contours(thres_images[ti], W, H, contours);
    if (n_contours) *n_contours += (int)contours.size();
    for (size_t i = 0; i < contours.size(); i++) {
        if ((int)contours[i].size() <= minSize || (int)contours[i].size() >= maxSize) continue;
        approx_poly_dp(contours[i], double(contours[i].size()) * 0.05, approx);
        if (approx.size() != 4) continue;
        if (!is_contour_convex(approx)) continue;
        // :542-552 min-side test reads approxCurve[i] out of bounds -> never rejects
        Candidate cd;
        for (int k = 0; k < 4; k++) cd.c[k] = Pt2f{(float)approx[k].x, (float)approx[k].y};
        cd.idx = (int)i;
        cd.contour = contours[i];
        cd.id = -1;
        cd.nrot = 0;
        cands.push_back(std::move(cd));
    }
    }
    std::vector<char> swapped(cands.size(), 0);
    for (size_t i = 0; i < cands.size(); i++) {
        Pt2f* c = cands[i].c;
        float d1x = c[1].x - c[0].x, d1y = c[1].y - c[0].y, d2x = c[2].x - c[0].x, d2y = c[2].y - c[0].y;
        float o = (d1x * d2y) - (d1y * d2x);
        if (o < 0.0) {
            std::swap(c[1], c[3]);
            swapped[i] = 1;
        }
    }
    int n = (int)cands.size();
    std::vector<char> too_near_first, dummy;
    std::vector<char> remove(n, 0);
    std::vector<std::pair<int, int>> pairs;
#pragma omp parallel
    {
        std::vector<std::pair<int, int>> local;
#pragma omp for nowait
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                bool near = true;
                for (int c = 0; c < 4 && near; c++) {
                    float ddx = cands[i].c[c].x - cands[j].c[c].x, ddy = cands[i].c[c].y - cands[j].c[c].y;
                    float d = sqrt((double)ddx * ddx + (double)ddy * ddy);
                    near = d < 6;
                }
                if (near) local.push_back({i, j});
            }
#pragma omp critical
        pairs.insert(pairs.end(), local.begin(), local.end());
    }
    for (auto& pr : pairs) {
        if (perimeter(cands[pr.first].c) > perimeter(cands[pr.second].c)) remove[pr.second] = 1;
        else remove[pr.first] = 1;
    }
    for (int i = 0; i < n; i++) {
        if (remove[i]) continue;
        if (swapped[i]) std::reverse(cands[i].contour.begin(), cands[i].contour.end());
        out.push_back(std::move(cands[i]));
    }
}

// ------------------------------------------------------------------------------------------------
// warp (src/markerdetector.cpp:684-697): getPerspectiveTransform (LU) + warpPerspective(INTER_NEAREST)
// ------------------------------------------------------------------------------------------------
bool gauss8(double A[8][8], double* b) {
    for (int i = 0; i < 8; i++) {
        int k = i;
        for (int j = i + 1; j < 8; j++)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < DBL_EPSILON * 100) return false;
        if (k != i) {
            for (int j = i; j < 8; j++) std::swap(A[i][j], A[k][j]);
            std::swap(b[i], b[k]);
        }
        double d = -1 / A[i][i];
        for (int j = i + 1; j < 8; j++) {
            double alpha = A[j][i] * d;
            for (int c = i + 1; c < 8; c++) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = 7; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < 8; c++) s -= A[i][c] * b[c];
        b[i] = s / A[i][i];
    }
    return true;
}

bool homography4(const double* sx, const double* sy, const double* dx, const double* dy, double* M) {
    double a[8][8], b[8];
    for (int i = 0; i < 4; i++) {
        for (int c = 0; c < 8; c++) a[i][c] = a[i + 4][c] = 0;
        a[i][0] = a[i + 4][3] = sx[i];
        a[i][1] = a[i + 4][4] = sy[i];
        a[i][2] = a[i + 4][5] = 1;
        a[i][6] = -sx[i] * dx[i];
        a[i][7] = -sy[i] * dx[i];
        a[i + 4][6] = -sx[i] * dy[i];
        a[i + 4][7] = -sy[i] * dy[i];
        b[i] = dx[i];
        b[i + 4] = dy[i];
    }
    if (!gauss8(a, b)) return false;
    for (int i = 0; i < 8; i++) M[i] = b[i];
    M[8] = 1;
    return true;
}

bool warp_marker(const uint8_t* grey, int W, int H, const Pt2f* q, int S, uint8_t* out) {
    double sx[4], sy[4], dx[4] = {0, (double)S - 1, (double)S - 1, 0}, dy[4] = {0, 0, (double)S - 1, (double)S - 1}, M[9];
    for (int i = 0; i < 4; i++) {
        sx[i] = q[i].x;
        sy[i] = q[i].y;
    }
    memset(out, 0, (size_t)S * S);
    if (!homography4(sx, sy, dx, dy, M)) return false;
    const double* m = M;
    double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (det == 0) return false;
    double d = 1. / det, I[9];
    I[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    I[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    I[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    I[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    I[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    I[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    I[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    I[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    I[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    int bh0 = std::min(16, S), bw0 = std::min(1024 / bh0, S);
    for (int y = 0; y < S; y++)
        for (int bx = 0; bx < S; bx += bw0) {
            double X0 = I[0] * bx + I[1] * y + I[2], Y0 = I[3] * bx + I[4] * y + I[5], W0 = I[6] * bx + I[7] * y + I[8];
            for (int x1 = 0; x1 < bw0 && bx + x1 < S; x1++) {
                double Wv = W0 + I[6] * x1;
                Wv = Wv ? 1. / Wv : 0;
                double fX = std::max(-2147483648.0, std::min(2147483647.0, (X0 + I[0] * x1) * Wv));
                double fY = std::max(-2147483648.0, std::min(2147483647.0, (Y0 + I[3] * x1) * Wv));
                int X = (int)rint(fX), Y = (int)rint(fY);
                if (X >= 0 && Y >= 0 && X < W && Y < H) out[y * S + bx + x1] = grey[(size_t)Y * W + X];
            }
        }
    return true;
}

int otsu(const uint8_t* img, int N) {
    int h[256] = {0};
    for (int i = 0; i < N; i++) h[img[i]]++;
    double mu = 0, scale = 1. / N;
    for (int i = 0; i < 256; i++) mu += i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
    for (int i = 0; i < 256; i++) {
        double p_i = h[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        double q2 = 1. - q1;
        if (std::min(q1, q2) < FLT_EPSILON || std::max(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        double mu2 = (mu - q1 * mu1) / q2;
        double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) {
            max_sigma = sigma;
            max_val = i;
        }
    }
    return (int)max_val;
}

int count_cell(const uint8_t* bw, int S, int x0, int y0, int cell) {
    int nz = 0;
    for (int y = 0; y < cell; y++)
        for (int x = 0; x < cell; x++) nz += bw[(y0 + y) * S + x0 + x] != 0;
    return nz;
}

// FiducidalMarkers::detect (src/arucofidmarkers.cpp:438-452, 100-137)
int fiducidal_detect(const uint8_t* canon, int S, int* nrot) {
    std::vector<uint8_t> bw((size_t)S * S);
    int t = otsu(canon, S * S);
    for (int i = 0; i < S * S; i++) bw[i] = canon[i] > t ? 255 : 0;
    *nrot = 0;
    int sw = S / 7;
    for (int y = 0; y < 7; y++) {
        int inc = (y == 0 || y == 6) ? 1 : 6;
        for (int x = 0; x < 7; x += inc)
            if (count_cell(bw.data(), S, x * sw, y * sw, sw) > (sw * sw) / 2) return -1;
    }
    int rot[4][5][5];
    for (int y = 0; y < 5; y++)
        for (int x = 0; x < 5; x++) rot[0][y][x] = count_cell(bw.data(), S, (x + 1) * sw, (y + 1) * sw, sw) > (sw * sw) / 2;
    static const int ids[4][5] = {{1, 0, 0, 0, 0}, {1, 0, 1, 1, 1}, {0, 1, 0, 0, 1}, {0, 1, 1, 1, 0}};
    auto hamm = [&](int r) {
        int dist = 0;
        for (int y = 0; y < 5; y++) {
            int minSum = 100000;
            for (int p = 0; p < 4; p++) {
                int sum = 0;
                for (int x = 0; x < 5; x++) sum += rot[r][y][x] != ids[p][x];
                minSum = std::min(minSum, sum);
            }
            dist += minSum;
        }
        return dist;
    };
    int minDist = hamm(0);
    for (int r = 1; r < 4; r++) {
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) rot[r][i][j] = rot[r - 1][5 - j - 1][i];
        int dist = hamm(r);
        if (dist < minDist) {
            minDist = dist;
            *nrot = r;
        }
    }
    if (minDist != 0) return -1;
    int id = 0;
    for (int y = 0; y < 5; y++) id |= ((rot[*nrot][y][1] << 1) | rot[*nrot][y][3]) << (2 * (4 - y));
    return id;
}

// HighlyReliableMarkers (src/highlyreliablemarkers.cpp:149-180, 277-289, 312-383, 387-496)
struct HrmDict {
    int n = 0, count = 0, correction = 0, root = 0;
    std::vector<std::vector<uint8_t>> rot0;  // rotation-0 bit strings
    std::vector<std::pair<uint32_t, uint32_t>> order;
    std::vector<std::pair<int, int>> tree;
};

void code_rotations(const uint8_t* code, int n, std::vector<uint8_t> bits[4], uint32_t ids[4]) {
    for (int r = 0; r < 4; r++) {
        bits[r].assign((size_t)n * n, 0);
        ids[r] = 0;
    }
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            for (int r = 0; r < 4; r++) {
                int _x = x, _y = y;
                if (r == 1) { _y = x; _x = n - y - 1; }
                else if (r == 2) { _y = n - y - 1; _x = n - x - 1; }
                else if (r == 3) { _y = n - x - 1; _x = y; }
                int pos = _y * n + _x;
                bool val = code[y * n + x] != 0;
                bits[r][pos] = val;
                if (val) ids[r] |= (uint32_t)(((uint64_t)2 << (pos & 31)) & 0xFFFFFFFFu);  // x86 shl semantics (B.4)
            }
}

void build_dict(HrmDict& D, const uint8_t* bits, int n, int count, int tau0, float rate) {
    D.n = n;
    D.count = count;
    D.correction = rate * ((tau0 - 1) / 2);
    D.rot0.clear();
    D.order.clear();
    for (int i = 0; i < count; i++) {
        std::vector<uint8_t> rb[4];
        uint32_t ids[4];
        code_rotations(bits + (size_t)i * n * n, n, rb, ids);
        D.rot0.push_back(rb[0]);
        D.order.push_back({ids[0], (uint32_t)i});
    }
    std::sort(D.order.begin(), D.order.end());
    unsigned sz = count, levels = 0;
    while (pow(float(2), float(levels)) <= sz) levels++;
    std::vector<bool> visited(sz, false);
    unsigned rootIdx = sz / 2;
    visited[rootIdx] = true;
    D.root = rootIdx;
    std::vector<std::pair<unsigned, unsigned>> intervals;
    intervals.push_back({0, rootIdx});
    intervals.push_back({rootIdx, sz});
    D.tree.assign(sz, {0, 0});
    D.tree[rootIdx].first = !visited[(0 + rootIdx) / 2] ? (int)((0 + rootIdx) / 2) : -1;
    D.tree[rootIdx].second = !visited[(rootIdx + sz) / 2] ? (int)((rootIdx + sz) / 2) : -1;
    for (unsigned i = 1; i < levels; i++) {
        unsigned nint = intervals.size();
        for (unsigned j = 0; j < nint; j++) {
            unsigned lo = intervals.back().first, hi = intervals.back().second;
            intervals.pop_back();
            unsigned center = (hi + lo) / 2;
            if (!visited[center]) visited[center] = true;
            else continue;
            unsigned lc = (lo + center) / 2, hc = (center + hi) / 2;
            if (!visited[lc]) {
                intervals.insert(intervals.begin(), {lo, center});
                D.tree[center].first = lc;
            } else D.tree[center].first = -1;
            if (!visited[hc]) {
                intervals.insert(intervals.begin(), {center, hi});
                D.tree[center].second = hc;
            } else D.tree[center].second = -1;
        }
    }
}

int hrm_detect(const HrmDict& D, const uint8_t* canon, int S, int* nrot) {
    std::vector<uint8_t> bw((size_t)S * S);
    int t = otsu(canon, S * S);
    for (int i = 0; i < S * S; i++) bw[i] = canon[i] > t ? 255 : 0;
    *nrot = 0;
    int n = D.n, cell = S / (n + 2);
    std::vector<uint8_t> code((size_t)n * n);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) code[y * n + x] = count_cell(bw.data(), S, (x + 1) * cell, (y + 1) * cell, cell) > (cell * cell) / 2;
    std::vector<uint8_t> bits[4];
    uint32_t ids[4];
    code_rotations(code.data(), n, bits, ids);
    for (int r = 0; r < 4; r++) {
        int pos = D.root;
        while (pos != -1) {
            uint32_t pid = D.order[pos].first;
            if (pid == ids[r]) {
                *nrot = r;
                return (int)D.order[pos].second;
            }
            pos = pid < ids[r] ? D.tree[pos].second : D.tree[pos].first;
        }
    }
    unsigned res = n * n, minMarker = 0, minRot = 0;
    for (int i = 0; i < D.count; i++) {
        unsigned r2 = n * n, mr = 0;
        for (unsigned r = 0; r < 4; r++) {
            unsigned hd = 0;
            for (int k = 0; k < n * n; k++) hd += D.rot0[i][k] != bits[r][k];
            if (hd < r2) {
                mr = r;
                r2 = hd;
            }
        }
        if (r2 < res) {
            minMarker = i;
            minRot = mr;
            res = r2;
        }
    }
    if (res <= (unsigned)D.correction) {
        *nrot = minRot;
        return (int)minMarker;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// camera model
// ------------------------------------------------------------------------------------------------
struct Cam {
    bool hasK = false, hasD = false;
    float Kf[9];
    double fx, fy, cx, cy, k1 = 0, k2 = 0, p1 = 0, p2 = 0, k3 = 0;
};

void undistort_px(const Cam& c, float u, float v, float* ou, float* ov, int iters = 5) {
    double x0 = (u - c.cx) / c.fx, y0 = (v - c.cy) / c.fy, x = x0, y = y0;
    for (int j = 0; j < iters; j++) {
        double r2 = x * x + y * y;
        double icdist = 1. / (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
        double dX = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x), dY = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ou = (float)(x * c.fx + c.cx);
    *ov = (float)(y * c.fy + c.cy);
}

void project_norm(const Cam& c, double x, double y, double* u, double* v) {
    double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2, a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    double cd = 1 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
    *u = (x * cd + c.p1 * a1 + c.p2 * a2) * c.fx + c.cx;
    *v = (y * cd + c.p1 * a3 + c.p2 * a1) * c.fy + c.cy;
}

// cv::solve(A, B, X, DECOMP_SVD) for an m x 2 CV_32F system, as interpolate2Dline / getCrossPoint call it
// (src/markerdetector.cpp:112,124,138).  OpenCV (un-vendored dependency; modules/core/src/lapack.cpp, 4.x)
// transposes A, runs its one-sided Jacobi SVD in f32 with f64 dot products (JacobiSVDImpl_<float>:
// eps = 2*FLT_EPSILON, at most max(m,30) sweeps, singular values sorted descending, rows scaled by 1/w)
// and back-substitutes with SVBkSb (f32 products accumulated in f64, threshold (float)(2*DBL_EPSILON)*sum(w)).
// cv2 4.13 is bit-identical to this for m < 25 (tests/test_oracle_cross.py); wheels built with a LAPACK HAL
// switch to sgesdd at m >= 25, whose rounding depends on the BLAS kernels of the host CPU.
void svd_solve_f32_m2(const float* c0, const float* c1, int m, const float* b, float X[2]) {
    std::vector<float> a0(c0, c0 + m), a1(c1, c1 + m);
    float* At[2] = {a0.data(), a1.data()};
    double W[2];
    float Vt[2][2] = {{1.f, 0.f}, {0.f, 1.f}};
    for (int i = 0; i < 2; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { float t = At[i][k]; sd += (double)t * t; }
        W[i] = sd;
    }
    const float eps = FLT_EPSILON * 2;
    const int max_iter = std::max(m, 30);
    for (int iter = 0; iter < max_iter; iter++) {
        float* Ai = At[0];
        float* Aj = At[1];
        double a = W[0], p = 0, bb = W[1];
        for (int k = 0; k < m; k++) p += (double)Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * bb)) break;
        p *= 2;
        double beta = a - bb, gamma = hypot(p, beta);
        float c, s;
        if (beta < 0) {
            double delta = (gamma - beta) * 0.5;
            s = (float)sqrt(delta / gamma);
            c = (float)(p / (gamma * s * 2));
        } else {
            c = (float)sqrt((gamma + beta) / (gamma * 2));
            s = (float)(p / (gamma * c * 2));
        }
        a = bb = 0;
        for (int k = 0; k < m; k++) {
            float t0 = c * Ai[k] + s * Aj[k];
            float t1 = -s * Ai[k] + c * Aj[k];
            Ai[k] = t0; Aj[k] = t1;
            a += (double)t0 * t0; bb += (double)t1 * t1;
        }
        W[0] = a; W[1] = bb;
        for (int k = 0; k < 2; k++) {
            float t0 = c * Vt[0][k] + s * Vt[1][k];
            float t1 = -s * Vt[0][k] + c * Vt[1][k];
            Vt[0][k] = t0; Vt[1][k] = t1;
        }
    }
    for (int i = 0; i < 2; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { float t = At[i][k]; sd += (double)t * t; }
        W[i] = sqrt(sd);
    }
    int o0 = 0, o1 = 1;
    if (W[0] < W[1]) { std::swap(W[0], W[1]); o0 = 1; o1 = 0; }
    const int ord[2] = {o0, o1};
    float w[2] = {(float)W[0], (float)W[1]};
    for (int i = 0; i < 2; i++) {
        float sc = (float)(W[i] > (double)FLT_MIN ? 1 / W[i] : 0.);
        float* r = At[ord[i]];
        for (int k = 0; k < m; k++) r[k] *= sc;
    }
    double threshold = ((double)w[0] + (double)w[1]) * (double)(float)(DBL_EPSILON * 2);
    X[0] = X[1] = 0.f;
    for (int i = 0; i < 2; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        const float* u = At[ord[i]];
        double sacc = 0;
        for (int j = 0; j < m; j++) sacc += u[j] * b[j];  // f32 product, f64 accumulator (SVBkSbImpl_)
        sacc *= wi;
        for (int j = 0; j < 2; j++) X[j] = (float)(X[j] + sacc * Vt[ord[i]][j]);
    }
}

// interpolate2Dline (src/markerdetector.cpp:83-130)
void interpolate_2d_line(const std::vector<Pt2f>& pts, float line[3]) {
    float minX, maxX, minY, maxY;
    minX = maxX = pts[0].x;
    minY = maxY = pts[0].y;
    for (size_t i = 1; i < pts.size(); i++) {
        minX = std::min(minX, pts[i].x); maxX = std::max(maxX, pts[i].x);
        minY = std::min(minY, pts[i].y); maxY = std::max(maxY, pts[i].y);
    }
    const int m = (int)pts.size();
    std::vector<float> a(m), one(m, 1.f), rhs(m);
    float X[2];
    if (maxX - minX > maxY - minY) {  // Ax + C = y
        for (int i = 0; i < m; i++) { a[i] = pts[i].x; rhs[i] = pts[i].y; }
        svd_solve_f32_m2(a.data(), one.data(), m, rhs.data(), X);
        line[0] = X[0]; line[1] = -1.f; line[2] = X[1];
    } else {  // By + C = x
        for (int i = 0; i < m; i++) { a[i] = pts[i].y; rhs[i] = pts[i].x; }
        svd_solve_f32_m2(a.data(), one.data(), m, rhs.data(), X);
        line[0] = -1.f; line[1] = X[0]; line[2] = X[1];
    }
}

// getCrossPoint (src/markerdetector.cpp:132-139): Matx22f::solve(Vec2f, DECOMP_SVD)
void cross_point(const float* l1, const float* l2, float* x, float* y) {
    const float c0[2] = {l1[0], l2[0]}, c1[2] = {l1[1], l2[1]}, rhs[2] = {-l1[2], -l2[2]};
    float X[2];
    svd_solve_f32_m2(c0, c1, 2, rhs, X);
    *x = X[0];
    *y = X[1];
}

// refineCandidateLines (src/markerdetector.cpp:931-997)
void refine_lines(Candidate& cd, const Cam& cam) {
    const std::vector<Pt>& ct = cd.contour;
    const int n = (int)ct.size();
    int ci[4] = {0, 0, 0, 0};
    for (int j = 0; j < n; j++)
        for (int k = 0; k < 4; k++)
            if (ct[j].x == (int)rintf(cd.c[k].x) && ct[j].y == (int)rintf(cd.c[k].y)) ci[k] = j;
    bool inverse;
    if ((ci[1] > ci[0]) && (ci[2] > ci[1] || ci[2] < ci[0])) inverse = false;
    else if (ci[2] > ci[1] && ci[2] < ci[0]) inverse = false;
    else inverse = true;
    int inc = inverse ? -1 : 1;
    bool und = cam.hasK && cam.hasD;
    std::vector<Pt2f> c2f(n);
    for (int j = 0; j < n; j++) {
        c2f[j] = Pt2f{(float)ct[j].x, (float)ct[j].y};
        if (und) undistort_px(cam, c2f[j].x, c2f[j].y, &c2f[j].x, &c2f[j].y);
    }
    float lines[4][3];
    for (int l = 0; l < 4; l++) {
        std::vector<Pt2f> pts;
        int j = ci[l];
        while (j != ci[(l + 1) % 4]) {
            pts.push_back(c2f[j]);
            j = ((j + inc) % n + n) % n;
        }
        if (pts.size() == 1) pts.push_back(c2f[ci[(l + 1) % 4]]);
        interpolate_2d_line(pts, lines[l]);
    }
    for (int i = 0; i < 4; i++) {
        const float* l1 = lines[i];
        const float* l2 = lines[(i + 3) % 4];
        float x, y;
        cross_point(l1, l2, &x, &y);
        if (und) {
            float xn = (x - cam.Kf[2]) / cam.Kf[0], yn = (y - cam.Kf[5]) / cam.Kf[4];
            double u, v;
            project_norm(cam, xn, yn, &u, &v);
            x = (float)u;
            y = (float)v;
        }
        cd.c[i] = Pt2f{x, y};
    }
}

// cv::cornerSubPix
void corner_subpix(const uint8_t* img, int W, int H, Pt2f* pt, int w) {
    const int win = 2 * w + 1, pw = win + 2;
    std::vector<float> mask((size_t)win * win), buf((size_t)pw * pw);
    for (int i = 0; i < win; i++) {
        float y = (float)(i - w) / w, vy = expf(-y * y);
        for (int j = 0; j < win; j++) {
            float x = (float)(j - w) / w;
            mask[i * win + j] = (float)(vy * expf(-x * x));
        }
    }
    Pt2f cT = *pt, cI = cT;
    if (!(cT.x >= 0 && cT.x < W && cT.y >= 0 && cT.y < H)) return;
    int iter = 0;
    double err = 0;
    const double eps = 0.005 * 0.005;
    do {
        float cx = cI.x - (pw - 1) * 0.5f, cy = cI.y - (pw - 1) * 0.5f;
        int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
        float a = cx - ipx, b = cy - ipy;
        float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
        for (int i = 0; i < pw; i++)
            for (int j = 0; j < pw; j++) {
                int x0 = std::min(std::max(ipx + j, 0), W - 1), x1 = std::min(std::max(ipx + j + 1, 0), W - 1);
                int y0 = std::min(std::max(ipy + i, 0), H - 1), y1 = std::min(std::max(ipy + i + 1, 0), H - 1);
                buf[i * pw + j] = img[(size_t)y0 * W + x0] * a11 + img[(size_t)y0 * W + x1] * a12 + img[(size_t)y1 * W + x0] * a21 +
                                  img[(size_t)y1 * W + x1] * a22;
            }
        double A = 0, B = 0, C = 0, bb1 = 0, bb2 = 0;
        for (int i = 0; i < win; i++) {
            const float* sp = &buf[(i + 1) * pw + 1];
            double py = i - w;
            for (int j = 0; j < win; j++) {
                double m = mask[i * win + j];
                double tgx = sp[j + 1] - sp[j - 1], tgy = sp[j + pw] - sp[j - pw];
                double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m, px = j - w;
                A += gxx; B += gxy; C += gyy;
                bb1 += gxx * px + gxy * py;
                bb2 += gxy * px + gyy * py;
            }
        }
        double det = A * C - B * B;
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        double scale = 1.0 / det;
        Pt2f cI2;
        cI2.x = (float)(cI.x + C * scale * bb1 - B * scale * bb2);
        cI2.y = (float)(cI.y - B * scale * bb1 + A * scale * bb2);
        err = (cI2.x - cI.x) * (cI2.x - cI.x) + (cI2.y - cI.y) * (cI2.y - cI.y);
        cI = cI2;
        if (cI.x < 0 || cI.x >= W || cI.y < 0 || cI.y >= H) break;
    } while (++iter < 8 && err > eps);
    if (fabs(cI.x - cT.x) > w || fabs(cI.y - cT.y) > w) cI = cT;
    *pt = cI;
}

// cv::getRectSubPix for 8u -> 8u (generic C++ path: 16-bit fixed-point bilinear weights, replicated border)
void rect_subpix_u8(const uint8_t* img, int W, int H, float cx, float cy, int S, int* out) {
    cx -= (S - 1) * 0.5f;
    cy -= (S - 1) * 0.5f;
    int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
    float a = cx - ipx, b = cy - ipy;
    int a11 = (int)lrintf((1.f - a) * (1.f - b) * 65536.f), a12 = (int)lrintf(a * (1.f - b) * 65536.f);
    int a21 = (int)lrintf((1.f - a) * b * 65536.f), a22 = (int)lrintf(a * b * 65536.f);
    for (int i = 0; i < S; i++)
        for (int j = 0; j < S; j++) {
            int x0 = std::min(std::max(ipx + j, 0), W - 1), x1 = std::min(std::max(ipx + j + 1, 0), W - 1);
            int y0 = std::min(std::max(ipy + i, 0), H - 1), y1 = std::min(std::max(ipy + i + 1, 0), H - 1);
            int v = img[(size_t)y0 * W + x0] * a11 + img[(size_t)y0 * W + x1] * a12 + img[(size_t)y1 * W + x0] * a21 +
                    img[(size_t)y1 * W + x1] * a22;
            out[i * S + j] = (v + (1 << 15)) >> 16;
        }
}

// SubPixelCorner::RefineCorner (src/subpixelcorner.cpp:70-189) with its quirks: one iteration,
// D == 0 in the y update, u8 patch before Sobel.
void harris_refine(const uint8_t* img, int W, int H, Pt2f* pt) {
    const int win = 15, S = 17;
    float mx[15];
    double coeff = 1. / (win * win);
    for (int i = -win / 2, k = 0; i <= win / 2; i++, k++) mx[k] = (float)exp(-i * i * coeff);
    Pt2f est = *pt;
    if (est.x < 0 || est.y < 0 || est.y > H || est.y > W) return;
    int patch[17 * 17];
    rect_subpix_u8(img, W, H, est.x, est.y, S, patch);
    double A = 0, B = 0, C = 0, D = 0, E = 0, F = 0;
    for (int i = 1; i <= win; i++) {
        int ly = i - win / 2 - 1;
        for (int j = 1; j <= win; j++) {
            int lx = j - win / 2 - 1;
            const int* q = patch + i * S + j;
            float dx = (float)((q[-S + 1] - q[-S - 1]) + 2 * (q[1] - q[-1]) + (q[S + 1] - q[S - 1]));
            float dy = (float)((q[S - 1] - q[-S - 1]) + 2 * (q[S] - q[-S]) + (q[S + 1] - q[-S + 1]));
            double val = (float)(mx[lx + win / 2] * mx[ly + win / 2]);
            double dxx = (float)(dx * dx) * val, dyy = (float)(dy * dy) * val, dxy = (float)(dx * dy) * val;
            A += dxx; B += dxy; E += dyy;
            C += dxx * lx + dxy * ly;
            F += dxy * lx + dyy * ly;
        }
    }
    double det = A * E - B * B;
    Pt2f cur = est;
    if (fabs(det) > DBL_EPSILON * DBL_EPSILON) {
        det = 1.0 / det;
        est.x = (float)(cur.x + ((C * E) - (B * F)) * det);
        est.y = (float)(cur.y + ((A * F) - (C * D)) * det);
    }
    if (fabs(pt->x - est.x) > win || fabs(pt->y - est.y) > win) est = *pt;
    *pt = est;
}

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

// findCornerMaxima (src/markerdetector.cpp:157-199): cornerHarris(3,3,0.04) on the +-wsize region (the Sobel
// of a C++ ROI reads the parent image outside the ROI; the 3x3 box sum reflects at the ROI border), 4x4 block
// sums over the interior, maximum of the centre-weighted response.
void find_corner_maxima(const uint8_t* grey, int W, int H, Pt2f* pt, int wsize) {
    int x0 = std::max(0, (int)(pt->x - wsize)), y0 = std::max(0, (int)(pt->y - wsize));
    int x1 = std::min(W, (int)(pt->x + wsize)), y1 = std::min(H, (int)(pt->y + wsize));
    int rw = x1 - x0, rh = y1 - y0;
    if (rw <= 0 || rh <= 0) { *pt = Pt2f{-1.f + x0, -1.f + y0}; return; }
    const float scale = (float)(1.0 / (4.0 * 3.0 * 255.0)), c2 = (float)(2.0 * (1.0 / (4.0 * 3.0 * 255.0)));
    auto px = [&](int x, int y) { return (float)grey[(size_t)reflect101(y, H) * W + reflect101(x, W)]; };
    std::vector<float> ca((size_t)rw * rh), cb((size_t)rw * rh), cc((size_t)rw * rh), harr((size_t)rw * rh);
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            int gx = x0 + x, gy = y0 + y;
            float r_m = px(gx + 1, gy - 1) - px(gx - 1, gy - 1), r_0 = px(gx + 1, gy) - px(gx - 1, gy), r_p = px(gx + 1, gy + 1) - px(gx - 1, gy + 1);
            float dx = c2 * r_0 + scale * (r_m + r_p);
            float s_m = c2 * px(gx, gy - 1) + scale * (px(gx - 1, gy - 1) + px(gx + 1, gy - 1));
            float s_p = c2 * px(gx, gy + 1) + scale * (px(gx - 1, gy + 1) + px(gx + 1, gy + 1));
            float dy = s_p - s_m;
            ca[y * rw + x] = dx * dx; cb[y * rw + x] = dx * dy; cc[y * rw + x] = dy * dy;
        }
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            double sa = 0, sb = 0, sc = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    int yy = reflect101(y + dy, rh), xx = reflect101(x + dx, rw);
                    sa += ca[yy * rw + xx]; sb += cb[yy * rw + xx]; sc += cc[yy * rw + xx];
                }
            float a = (float)sa, b = (float)sb, c = (float)sc;
            harr[y * rw + x] = (float)(a * c - b * b - 0.04 * (a + c) * (a + c));
        }
    std::vector<float> hs(harr);
    for (int y = 4; y < rh - 4; y++)
        for (int x = 4; x < rw - 4; x++) {
            double sum = 0;
            for (int dy = 0; dy < 4; dy++)
                for (int dx = 0; dx < 4; dx++) sum += harr[(y + dy) * rw + x + dx];
            hs[y * rw + x] = (float)sum;
        }
    float bx = -1, by = -1;
    float ccx = (float)(rw / 2), ccy = (float)(rh / 2), den = (float)(rw / 2 + rh / 2);
    double maxv = 0;
    for (int i = 0; i < rh; i++)
        for (int x = 0; x < rw; x++) {
            float d = (float)(fabs(ccx - x) + fabs(ccy - i)) / den;
            float w = (float)(1. - d);
            float v = w * hs[i * rw + x];
            if (v > maxv) { maxv = v; bx = (float)x; by = (float)i; }
        }
    *pt = Pt2f{bx + x0, by + y0};
}

// ------------------------------------------------------------------------------------------------
// cv::solvePnP(ITERATIVE), 4 coplanar points: homography init + damped Gauss-Newton on the reprojection
// error with forward-difference Jacobian
// ------------------------------------------------------------------------------------------------
void rodrigues(const double* r, double* R) {
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0);
        return;
    }
    double c = cos(th), s = sin(th), c1 = 1 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
    double Rm[9] = {c + c1 * x * x, c1 * x * y - s * z, c1 * x * z + s * y, c1 * x * y + s * z, c + c1 * y * y,
                    c1 * y * z - s * x, c1 * x * z - s * y, c1 * y * z + s * x, c + c1 * z * z};
    memcpy(R, Rm, sizeof(Rm));
}

void jacobi_eig3(double A[3][3], double V[3][3]) {  // symmetric 3x3 eigen decomposition
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double offd = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (offd < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (fabs(A[p][q]) < 1e-300) continue;
                double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
                double t = (th >= 0 ? 1 : -1) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}

void nearest_rotation(double* R) {  // R (R^T R)^(-1/2) = U V^T
    double M[3][3], V[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            M[i][j] = 0;
            for (int k = 0; k < 3; k++) M[i][j] += R[k * 3 + i] * R[k * 3 + j];
        }
    jacobi_eig3(M, V);
    double S[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            S[i][j] = 0;
            for (int k = 0; k < 3; k++) S[i][j] += V[i][k] * V[j][k] / sqrt(M[k][k]);
        }
    double O[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            O[i * 3 + j] = 0;
            for (int k = 0; k < 3; k++) O[i * 3 + j] += R[i * 3 + k] * S[k][j];
        }
    memcpy(R, O, sizeof(O));
}

void mat_to_rvec(const double* R, double* r) {
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25), c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : (c < -1. ? -1. : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) r[0] = r[1] = r[2] = 0;
        else {
            double t = (R[0] + 1) * 0.5;
            rx = sqrt(std::max(t, 0.));
            t = (R[4] + 1) * 0.5;
            ry = sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5;
            rz = sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        }
    } else {
        double vth = theta / (2 * s);
        r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
    }
}

void project4(const Cam& c, const double* p, const float* obj, double* uv) {
    double R[9];
    rodrigues(p, R);
    for (int i = 0; i < 4; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + p[3], y = R[3] * X + R[4] * Y + R[5] * Z + p[4],
               z = R[6] * X + R[7] * Y + R[8] * Z + p[5];
        z = z ? 1. / z : 1.;
        project_norm(c, x * z, y * z, &uv[2 * i], &uv[2 * i + 1]);
    }
}

bool solve_pnp(const Cam& cam, const Pt2f* corners, float size, double* rvec, double* tvec) {
    float h = size / 2.f;
    float obj[12] = {-h, -h, 0, -h, h, 0, h, h, 0, h, -h, 0};  // src/marker.cpp:91-108
    double sx[4], sy[4], nx[4], ny[4];
    for (int i = 0; i < 4; i++) {
        double x0 = (corners[i].x - cam.cx) / cam.fx, y0 = (corners[i].y - cam.cy) / cam.fy, x = x0, y = y0;
        for (int j = 0; j < 30; j++) {
            double r2 = x * x + y * y, ic = 1. / (1 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2);
            double dX = 2 * cam.p1 * x * y + cam.p2 * (r2 + 2 * x * x), dY = cam.p1 * (r2 + 2 * y * y) + 2 * cam.p2 * x * y;
            x = (x0 - dX) * ic;
            y = (y0 - dY) * ic;
        }
        nx[i] = x; ny[i] = y;
        sx[i] = obj[3 * i]; sy[i] = obj[3 * i + 1];
    }
    double Hm[9];
    if (!homography4(sx, sy, nx, ny, Hm)) return false;
    double h1[3] = {Hm[0], Hm[3], Hm[6]}, h2[3] = {Hm[1], Hm[4], Hm[7]}, h3[3] = {Hm[2], Hm[5], Hm[8]};
    double n1 = sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]), n2 = sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
    if (!(n1 > 0) || !(n2 > 0)) return false;
    for (int i = 0; i < 3; i++) { h1[i] /= n1; h2[i] /= n2; }
    double p[6];
    for (int i = 0; i < 3; i++) p[3 + i] = h3[i] * 2. / (n1 + n2);
    double c3[3] = {h1[1] * h2[2] - h1[2] * h2[1], h1[2] * h2[0] - h1[0] * h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
    double R[9] = {h1[0], h2[0], c3[0], h1[1], h2[1], c3[1], h1[2], h2[2], c3[2]};
    nearest_rotation(R);
    mat_to_rvec(R, p);
    // Levenberg-Marquardt with the schedule of OpenCV's CvLevMarq (cvFindExtrinsicCameraParams2): lambda = 10^k,
    // k = -3 at start, diag(JtJ) *= 1 + lambda; a step that raises |err| is retried with k+1 (<= 16), an accepted
    // step lowers k; at most 20 accepted steps, stop when |dp|/|p| < FLT_EPSILON.  Jacobian by central differences.
    double m[8], uv[8], err[8];
    for (int i = 0; i < 4; i++) { m[2 * i] = corners[i].x; m[2 * i + 1] = corners[i].y; }
    int k10 = -3, iters = 0;
    double prevErrNorm = 0;
    for (;;) {
        project4(cam, p, obj, uv);
        for (int i = 0; i < 8; i++) err[i] = uv[i] - m[i];
        double J[8][6];
        for (int k = 0; k < 6; k++) {
            double hs = 1e-6 * std::max(1.0, fabs(p[k])), pp[6], up[8], um[8];
            memcpy(pp, p, sizeof(pp));
            pp[k] = p[k] + hs;
            project4(cam, pp, obj, up);
            pp[k] = p[k] - hs;
            project4(cam, pp, obj, um);
            for (int i = 0; i < 8; i++) J[i][k] = (up[i] - um[i]) / (2 * hs);
        }
        double JtJ[6][6], JtE[6], prev[6];
        for (int i = 0; i < 6; i++) {
            JtE[i] = 0;
            for (int k = 0; k < 8; k++) JtE[i] += J[k][i] * err[k];
            for (int j = 0; j < 6; j++) {
                JtJ[i][j] = 0;
                for (int k = 0; k < 8; k++) JtJ[i][j] += J[k][i] * J[k][j];
            }
            prev[i] = p[i];
        }
        if (iters == 0) {
            prevErrNorm = 0;
            for (int i = 0; i < 8; i++) prevErrNorm += err[i] * err[i];
            prevErrNorm = sqrt(prevErrNorm);
        }
        double errNorm = 0;
        for (;;) {
            double lambda = pow(10.0, k10), A[6][7];
            for (int i = 0; i < 6; i++) {
                for (int j = 0; j < 6; j++) A[i][j] = JtJ[i][j];
                A[i][i] *= 1 + lambda;
                A[i][6] = JtE[i];
            }
            bool ok = true;
            for (int i = 0; i < 6 && ok; i++) {  // Gauss-Jordan with partial pivoting
                int piv = i;
                for (int j = i + 1; j < 6; j++)
                    if (fabs(A[j][i]) > fabs(A[piv][i])) piv = j;
                if (fabs(A[piv][i]) < 1e-300) { ok = false; break; }
                for (int c = 0; c < 7; c++) std::swap(A[i][c], A[piv][c]);
                for (int j = 0; j < 6; j++)
                    if (j != i) {
                        double f = A[j][i] / A[i][i];
                        for (int c = i; c < 7; c++) A[j][c] -= f * A[i][c];
                    }
            }
            for (int i = 0; i < 6; i++) p[i] = prev[i] - (ok ? A[i][6] / A[i][i] : 0.0);
            project4(cam, p, obj, uv);
            errNorm = 0;
            for (int i = 0; i < 8; i++) errNorm += (uv[i] - m[i]) * (uv[i] - m[i]);
            errNorm = sqrt(errNorm);
            if (errNorm > prevErrNorm && ++k10 <= 16) continue;
            break;
        }
        k10 = std::max(k10 - 1, -16);
        double dn = 0, pn = 0;
        for (int i = 0; i < 6; i++) { dn += (p[i] - prev[i]) * (p[i] - prev[i]); pn += prev[i] * prev[i]; }
        if (++iters >= 20 || sqrt(dn) / sqrt(pn) < FLT_EPSILON) break;
        prevErrNorm = errNorm;
    }
    for (int i = 0; i < 3; i++) { rvec[i] = p[i]; tvec[i] = p[3 + i]; }
    return true;
}

void rotate_x_axis(double* rvec) {  // src/utils.cpp:16-30
    double Rd[9];
    rodrigues(rvec, Rd);
    float R[9], RX[9] = {1, 0, 0, 0, 0, 0, 0, 0, 0}, O[9];
    for (int i = 0; i < 9; i++) R[i] = (float)Rd[i];
    float a = 3.14159265358979323846 / 2;
    RX[4] = cos(a); RX[5] = -sin(a); RX[7] = sin(a); RX[8] = cos(a);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0;
            for (int k = 0; k < 3; k++) s += R[i * 3 + k] * RX[k * 3 + j];
            O[i * 3 + j] = s;
        }
    for (int i = 0; i < 9; i++) Rd[i] = O[i];
    nearest_rotation(Rd);
    mat_to_rvec(Rd, rvec);
}

}  // namespace

extern "C" {

struct orc_params {
    int32_t thres_method;
    double p1, p2;
    int32_t corner_method;
    float min_size, max_size;
    int32_t warp_size;
    float border_dist;
    int32_t locked_corners, erosion, decoder, set_y_perpendicular;
    int32_t p1_range;  // setThresholdParamRange
};

struct orc_marker {
    int32_t id, has_pose;
    float corners[8];
    float ssize, pad_;
    double rvec[3], tvec[3];
};

struct orc_dict {
    int32_t n, count, tau0;
    float rate;
    const uint8_t* bits;
};

// optional intermediates
struct orc_debug {
    uint8_t* thres;       // W*H or NULL
    int32_t n_contours;
    int32_t n_candidates;
    int32_t cap_candidates;
    float* quads;         // cap*8 or NULL
    int32_t* ids;         // cap or NULL
    int32_t* nrot;        // cap or NULL
    uint8_t* canon;       // cap*S*S or NULL
};

void orc_threshold(const uint8_t* grey, int W, int H, int method, double p1, double p2, uint8_t* out) {
    if (method == 0) {
        int thr = (int)floor(p1);
        for (size_t i = 0; i < (size_t)W * H; i++) out[i] = (int)grey[i] > thr ? 0 : 255;
    } else if (method == 2) {
        canny(grey, W, H, 10, 220, out);
    } else {
        if (p1 < 3) p1 = 3;
        else if (((int)p1) % 2 != 1) p1 = (int)(p1 + 1);
        adaptive_threshold(grey, W, H, (int)p1, p2, out);
    }
}

int orc_find_contours(const uint8_t* bin, int W, int H, int cap_contours, int cap_points, int32_t* lens, int32_t* pts) {
    std::vector<std::vector<Pt>> cs;
    find_contours(bin, W, H, cs);
    if ((int)cs.size() > cap_contours) return -1;
    int np = 0;
    for (size_t i = 0; i < cs.size(); i++) {
        lens[i] = (int)cs[i].size();
        if (np + lens[i] > cap_points) return -1;
        for (auto& p : cs[i]) { pts[2 * np] = p.x; pts[2 * np + 1] = p.y; np++; }
    }
    return (int)cs.size();
}

int orc_approx_poly(const int32_t* pts, int n, double eps, int32_t* out, int cap) {
    std::vector<Pt> src(n), dst;
    for (int i = 0; i < n; i++) src[i] = Pt{pts[2 * i], pts[2 * i + 1]};
    approx_poly_dp(src, eps, dst);
    if ((int)dst.size() > cap) return -1;
    for (size_t i = 0; i < dst.size(); i++) { out[2 * i] = dst[i].x; out[2 * i + 1] = dst[i].y; }
    return (int)dst.size();
}

int orc_warp(const uint8_t* grey, int W, int H, const float* quad, int S, uint8_t* out) {
    Pt2f q[4];
    for (int i = 0; i < 4; i++) q[i] = Pt2f{quad[2 * i], quad[2 * i + 1]};
    return warp_marker(grey, W, H, q, S, out) ? 1 : 0;
}

int orc_otsu(const uint8_t* img, int N) { return otsu(img, N); }

// cv::solve(A[m x 2], B, X, DECOMP_SVD) in CV_32F (test hook for the restated Jacobi SVD)
void orc_svd_solve_f32(const float* A, int m, const float* B, float* X) {
    std::vector<float> c0(m), c1(m);
    for (int i = 0; i < m; i++) { c0[i] = A[2 * i]; c1[i] = A[2 * i + 1]; }
    svd_solve_f32_m2(c0.data(), c1.data(), m, B, X);
}

int orc_solve_pnp(const float* K, const float* D, const float* corners, float size, double* rvec, double* tvec) {
    Cam c;
    c.hasK = true;
    memcpy(c.Kf, K, sizeof(float) * 9);
    c.fx = K[0]; c.cx = K[2]; c.fy = K[4]; c.cy = K[5];
    if (D) { c.hasD = true; c.k1 = D[0]; c.k2 = D[1]; c.p1 = D[2]; c.p2 = D[3]; c.k3 = D[4]; }
    Pt2f q[4];
    for (int i = 0; i < 4; i++) q[i] = Pt2f{corners[2 * i], corners[2 * i + 1]};
    return solve_pnp(c, q, size, rvec, tvec) ? 1 : 0;
}

// MarkerDetector::detect for one grey frame. Returns the number of markers (or -1 if cap is too small,
// -2 for unsupported settings).
int orc_detect(const uint8_t* grey, int W, int H, const orc_params* P, const float* K, const float* D, float marker_size,
               const orc_dict* dict, orc_marker* out, int cap, orc_debug* dbg) {
    const int n_t = 2 * P->p1_range + 1;  // src/markerdetector.cpp:322-334
    std::vector<std::vector<uint8_t>> thr(n_t, std::vector<uint8_t>((size_t)W * H));
    std::vector<const uint8_t*> thr_ptr;
    for (int i = 0; i < n_t; i++) {
        double t1 = n_t == 1 ? P->p1 : P->p1 - P->p1_range + (double)P->p1_range * i;
        orc_threshold(grey, W, H, P->thres_method, t1, P->p2, thr[i].data());
        if (P->erosion) {
            std::vector<uint8_t> t2((size_t)W * H);
            erode3x3(thr[i].data(), W, H, t2.data());
            thr[i].swap(t2);
        }
        thr_ptr.push_back(thr[i].data());
    }
    if (dbg && dbg->thres) memcpy(dbg->thres, thr[n_t / 2].data(), thr[n_t / 2].size());
    std::vector<Candidate> cands;
    int ncont = 0;
    detect_rectangles(thr_ptr, W, H, P->min_size, P->max_size, cands, &ncont);
    Cam cam;
    if (K) {
        cam.hasK = true;
        memcpy(cam.Kf, K, sizeof(float) * 9);
        cam.fx = K[0]; cam.cx = K[2]; cam.fy = K[4]; cam.cy = K[5];
    }
    if (D) { cam.hasD = true; cam.k1 = D[0]; cam.k2 = D[1]; cam.p1 = D[2]; cam.p2 = D[3]; cam.k3 = D[4]; }
    HrmDict HD;
    if (P->decoder == 1) {
        if (!dict) return -2;
        build_dict(HD, dict->bits, dict->n, dict->count, dict->tau0, dict->rate);
    }
    const int S = P->warp_size;
    std::vector<uint8_t> canon((size_t)S * S);
    if (dbg) {
        dbg->n_contours = ncont;
        dbg->n_candidates = (int)cands.size();
    }
    struct Det { int id; Pt2f c[4]; };
    std::vector<Det> det;
    for (size_t i = 0; i < cands.size(); i++) {
        Candidate& cd = cands[i];
        if (dbg && (int)i < dbg->cap_candidates && dbg->quads)
            for (int k = 0; k < 4; k++) { dbg->quads[8 * i + 2 * k] = cd.c[k].x; dbg->quads[8 * i + 2 * k + 1] = cd.c[k].y; }
        warp_marker(grey, W, H, cd.c, S, canon.data());
        if (dbg && (int)i < dbg->cap_candidates && dbg->canon) memcpy(dbg->canon + i * (size_t)S * S, canon.data(), (size_t)S * S);
        int nrot = 0;
        int id = P->decoder == 1 ? hrm_detect(HD, canon.data(), S, &nrot) : fiducidal_detect(canon.data(), S, &nrot);
        cd.id = id;
        cd.nrot = nrot;
        if (dbg && (int)i < dbg->cap_candidates) {
            if (dbg->ids) dbg->ids[i] = id;
            if (dbg->nrot) dbg->nrot[i] = nrot;
        }
        if (id != -1) {
            if (P->corner_method == 3) refine_lines(cd, cam);
            std::rotate(cd.c, cd.c + 4 - nrot, cd.c + 4);
            Det d;
            d.id = id;
            memcpy(d.c, cd.c, sizeof(d.c));
            det.push_back(d);
        }
    }
    if (!det.empty() && (P->corner_method == 1 || P->corner_method == 2)) {  // :388-410
        int w = (int)P->p1;
        for (auto& d : det)
            for (int k = 0; k < 4; k++) {
                if (P->locked_corners) find_corner_maxima(grey, W, H, &d.c[k], w);
                if (P->corner_method == 1) harris_refine(grey, W, H, &d.c[k]);
                else corner_subpix(grey, W, H, &d.c[k], w);
            }
    }
    std::stable_sort(det.begin(), det.end(), [](const Det& a, const Det& b) { return a.id < b.id; });
    std::vector<char> rm(det.size(), 0);
    for (int i = 0; i < (int)det.size() - 1; i++)
        if (det[i].id == det[i + 1].id && !rm[i + 1]) {
            if (perimeter(det[i].c) > perimeter(det[i + 1].c)) rm[i + 1] = 1;
            else rm[i] = 1;
        }
    float bd = P->border_dist;
    int x0 = (int)lrintf(W * bd), y0 = (int)lrintf(H * bd), x1 = (int)lrintf(W * (1.0f - bd)), y1 = (int)lrintf(H * (1.0f - bd));
    int rx0 = std::min(x0, x1), ry0 = std::min(y0, y1), rx1 = std::max(x0, x1), ry1 = std::max(y0, y1);
    for (size_t i = 0; i < det.size(); i++)
        for (int c = 0; c < 4; c++) {
            float x = det[i].c[c].x, y = det[i].c[c].y;
            bool bad = !(std::isfinite(x) && std::isfinite(y));
            if (!bad) {
                int xi = (int)lrintf(x), yi = (int)lrintf(y);
                bad = !(xi >= rx0 && xi < rx1 && yi >= ry0 && yi < ry1);
            }
            if (bad) { rm[i] = 1; break; }
        }
    std::vector<Det> keep;
    for (size_t i = 0; i < det.size(); i++)
        if (!rm[i]) keep.push_back(det[i]);
    if ((int)keep.size() > cap) return -1;
    bool pose = cam.hasK && marker_size > 0;
#pragma omp parallel for
    for (int i = 0; i < (int)keep.size(); i++) {
        orc_marker& m = out[i];
        memset(&m, 0, sizeof(m));
        m.id = keep[i].id;
        for (int k = 0; k < 4; k++) { m.corners[2 * k] = keep[i].c[k].x; m.corners[2 * k + 1] = keep[i].c[k].y; }
        m.ssize = -1;
        if (pose) {
            m.has_pose = solve_pnp(cam, keep[i].c, marker_size, m.rvec, m.tvec) ? 1 : 0;
            if (m.has_pose && P->set_y_perpendicular) rotate_x_axis(m.rvec);
            m.ssize = marker_size;
        }
    }
    return (int)keep.size();
}

// frame-parallel batch for the CPU baseline: one frame per OpenMP thread
int orc_detect_batch(const uint8_t* frames, int W, int H, int n, const orc_params* P, const float* K, const float* D,
                     float marker_size, const orc_dict* dict, orc_marker* out, int cap, int32_t* counts, int threads) {
    int bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    omp_set_max_active_levels(1);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
    for (int f = 0; f < n; f++) {
        int r = orc_detect(frames + (size_t)f * W * H, W, H, P, K, D, marker_size, dict, out + (size_t)f * cap, cap, nullptr);
        counts[f] = r;
        if (r < 0) bad++;
    }
    return bad ? -1 : 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
