// Host-compiled harness around the per-thread logic headers of aruco_b200/csrc (the functions marked
// AB_HD).  TEST INFRASTRUCTURE: lets the `-m "not gpu"` suite exercise the exact code the kernels run per
// thread (border walk, polygon fit, homography, Otsu, decoders, pose ...) against cv2 / the oracle on a
// machine without a GPU.  It is never linked into libaruco_b200.so and is not a CPU fallback.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "../../aruco_b200/csrc/ab_trace.cuh"

extern "C" {

// pack a u8 image (non-zero = fg) into the padded bit layout
void hc_pack_bits(const uint8_t* img, int W, int H, uint32_t* out) {
    int wpr = ab::bit_words_per_row(W);
    memset(out, 0, sizeof(uint32_t) * ab::bit_image_words(W, H));
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (img[(size_t)y * W + x]) out[ab::bit_word_index(wpr, (x + 32 * ab::BIT_PAD) >> 5, y)] |= 1u << ((x + 32 * ab::BIT_PAD) & 31);
}

// all contours with min_len < n < max_len in OpenCV order (reverse discovery). Returns number of contours;
// lens[i], pts = concatenated (x,y) int32 pairs. total_contours = every border incl. isolated pixels.
int hc_find_contours(const uint8_t* img, int W, int H, int min_len, int max_len, int cap_contours, int cap_points,
                     int* lens, int32_t* pts, int* total_contours, int* n_candidates) {
    std::vector<uint32_t> bits(ab::bit_image_words(W, H));
    hc_pack_bits(img, W, H, bits.data());
    ab::BitImage im{bits.data(), ab::bit_words_per_row(W), W, H};
    struct Rec { int64_t key; std::vector<uint32_t> p; };
    std::vector<Rec> recs;
    int total = 0, ncand = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint32_t nb = ab::neighbours8(im, x, y);
            bool fg = img[(size_t)y * W + x] != 0;
            int type = -1;
            if (fg && ab::is_outer_candidate(nb)) type = 0;
            else if (!fg && (nb & (1u << 4)) && (nb & (1u << 2))) type = 1;
            if (type < 0) continue;
            ncand++;
            ab::TraceStart st;
            if (!ab::make_start(im, type, x, y, st)) {
                total++;
                if (1 > min_len && 1 < max_len) { Rec rc; rc.key = st.key; rc.p.push_back((uint32_t)x | ((uint32_t)y << 16)); recs.push_back(std::move(rc)); }
                continue;
            }
            int len = 0;
            int r = ab::find_start_bidir(im, st, 1 << 30, &len);
            if (r != ab::TRACE_OK) continue;
            total++;
            if (len <= min_len || len >= max_len) continue;
            Rec rc; rc.key = st.key; rc.p.resize(len);
            ab::trace_cycle(im, st, 1 << 30, &len, rc.p.data());
            recs.push_back(std::move(rc));
        }
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.key > b.key; });
    *total_contours = total;
    *n_candidates = ncand;
    int np = 0;
    if ((int)recs.size() > cap_contours) return -1;
    for (size_t i = 0; i < recs.size(); i++) {
        lens[i] = (int)recs[i].p.size();
        if (np + lens[i] > cap_points) return -1;
        for (uint32_t v : recs[i].p) { pts[2 * np] = v & 0xFFFF; pts[2 * np + 1] = v >> 16; np++; }
    }
    return (int)recs.size();
}
}

#include "../../aruco_b200/csrc/ab_math.cuh"

extern "C" {

// sequential driver of the approxPolyDP pieces (the kernel runs the same slices with a warp per contour)
int hc_approx_poly(const int32_t* pts, int count, double eps, int32_t* out, int cap) {
    if (count == 0) return 0;
    double E = eps * eps;
    std::vector<int> ox, oy;
    std::vector<std::pair<int, int>> stack;
    int rs = 0, pos = 0;
    bool le = false;
    int sx = 0, sy = 0;
    for (int it = 0; it < 3; it++) {
        long long maxd = 0;
        pos = (pos + rs) % count;
        sx = pts[2 * pos]; sy = pts[2 * pos + 1];
        int p0 = pos;
        for (int j = 1; j < count; j++) {
            int q = (p0 + j) % count;
            long long dx = pts[2 * q] - sx, dy = pts[2 * q + 1] - sy;
            long long d = dx * dx + dy * dy;
            if (d > maxd) { maxd = d; rs = j; }
        }
        le = (double)maxd <= E;
    }
    if (!le) {
        int A = pos % count, B = (rs + A) % count;
        stack.push_back({B, A});
        stack.push_back({A, B});
    } else { ox.push_back(sx); oy.push_back(sy); }
    while (!stack.empty()) {
        auto se = stack.back(); stack.pop_back();
        int s = se.first, e = se.second;
        int m = ((e - s + count) % count) - 1;
        bool lee = true; int split = 0;
        if (m > 0) {
            double maxd = 0;
            for (int i = 0; i < m; i++) {
                int q = (s + 1 + i) % count;
                double d = ab::seg_dist2(pts[2 * q], pts[2 * q + 1], pts[2 * s], pts[2 * s + 1], pts[2 * e], pts[2 * e + 1]);
                if (d > maxd) { maxd = d; split = q; }
            }
            lee = maxd <= E;
        }
        if (lee) { ox.push_back(pts[2 * s]); oy.push_back(pts[2 * s + 1]); }
        else { stack.push_back({split, e}); stack.push_back({s, split}); }
    }
    int n = ab::dp_cleanup(ox.data(), oy.data(), (int)ox.size(), E);
    if (n > cap) return -1;
    for (int i = 0; i < n; i++) { out[2 * i] = ox[i]; out[2 * i + 1] = oy[i]; }
    return n;
}

int hc_is_convex4(const int32_t* p) {
    int x[4] = {p[0], p[2], p[4], p[6]}, y[4] = {p[1], p[3], p[5], p[7]};
    return ab::is_convex4(x, y) ? 1 : 0;
}

int hc_perspective(const float* quad, int S, double* M) {
    float dst[8] = {0, 0, (float)(S - 1), 0, (float)(S - 1), (float)(S - 1), 0, (float)(S - 1)};
    return ab::perspective_transform(quad, dst, M) ? 1 : 0;
}

int hc_warp(const uint8_t* grey, int W, int H, const float* quad, int S, uint8_t* out) {
    double M[9], Mi[9];
    float dst[8] = {0, 0, (float)(S - 1), 0, (float)(S - 1), (float)(S - 1), 0, (float)(S - 1)};
    if (!ab::perspective_transform(quad, dst, M) || !ab::invert3(M, Mi)) return 0;
    int bw = ab::warp_block_width(S);
    for (int y = 0; y < S; y++)
        for (int x = 0; x < S; x++) {
            int sx, sy;
            ab::warp_src_coord(Mi, x, y, bw, &sx, &sy);
            out[y * S + x] = (sx >= 0 && sy >= 0 && sx < W && sy < H) ? grey[(size_t)sy * W + sx] : 0;
        }
    return 1;
}

int hc_otsu(const uint8_t* img, int N) {
    int h[256] = {0};
    for (int i = 0; i < N; i++) h[img[i]]++;
    return ab::otsu_threshold(h, N);
}

int hc_fid_decode(const uint8_t* canon, int S, int* nrot) {
    int t = hc_otsu(canon, S * S);
    int sw = S / 7;
    uint8_t cells[49];
    for (int cy = 0; cy < 7; cy++)
        for (int cx = 0; cx < 7; cx++) {
            int nz = 0;
            for (int y = 0; y < sw; y++)
                for (int x = 0; x < sw; x++) nz += canon[(cy * sw + y) * S + cx * sw + x] > t;
            cells[cy * 7 + cx] = nz > (sw * sw) / 2;
        }
    return ab::fid_decode(cells, nrot);
}

static ab::Camera make_cam(const float* K, const float* D) {
    ab::Camera c;
    memset(&c, 0, sizeof(c));
    c.has_K = K != nullptr;
    c.has_D = D != nullptr;
    if (K) { c.fxf = K[0]; c.cxf = K[2]; c.fyf = K[4]; c.cyf = K[5]; c.fx = K[0]; c.cx = K[2]; c.fy = K[4]; c.cy = K[5]; }
    if (D) { c.k1 = D[0]; c.k2 = D[1]; c.p1 = D[2]; c.p2 = D[3]; c.k3 = D[4]; }
    c.zero_D = D && D[0] == 0.f && D[1] == 0.f && D[2] == 0.f && D[3] == 0.f && D[4] == 0.f;
    return c;
}

int hc_solve_pnp(const float* K, const float* D, const float* corners, float size, double* rvec, double* tvec) {
    ab::Camera c = make_cam(K, D);
    return ab::solve_pnp_marker(c, corners, size, rvec, tvec) ? 1 : 0;
}

void hc_undistort_px(const float* K, const float* D, const float* in, int n, float* out) {
    ab::Camera c = make_cam(K, D);
    for (int i = 0; i < n; i++) ab::undistort_point_px(c, in[2 * i], in[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}

void hc_rotate_x_axis(double* rvec) { ab::rotate_x_axis(rvec); }

// getCrossPoint as k_refine_lines computes it
void hc_cross_point(const float* l1, const float* l2, float* xy) { ab::cross_point_f32(l1, l2, &xy[0], &xy[1]); }
}

extern "C" {
// diagnostic: total walk steps over all start candidates for the forward-only and the bidirectional search
void hc_walk_cost(const uint8_t* img, int W, int H, int max_len, long long* fwd_steps, long long* bidir_steps, long long* kept_points) {
    std::vector<uint32_t> bits(ab::bit_image_words(W, H));
    hc_pack_bits(img, W, H, bits.data());
    ab::BitImage im{bits.data(), ab::bit_words_per_row(W), W, H};
    long long fs = 0, bs = 0, kp = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint32_t nb = ab::neighbours8(im, x, y);
            bool fg = img[(size_t)y * W + x] != 0;
            int type = -1;
            if (fg && ab::is_outer_candidate(nb)) type = 0;
            else if (!fg && (nb & (1u << 4)) && (nb & (1u << 2))) type = 1;
            if (type < 0) continue;
            ab::TraceStart st;
            if (!ab::make_start(im, type, x, y, st)) continue;
            {
                int xx = st.x, yy = st.y, b = st.b, n = 0;
                for (;;) {
                    uint32_t nbb = ab::neighbours8(im, xx, yy);
                    ab::WalkState s{xx, yy, b};
                    if (n > 0 && ab::is_smaller_trigger(im, s, nbb, st.key)) break;
                    n++;
                    ab::walk_forward(s, nbb);
                    xx = s.x; yy = s.y; b = s.b;
                    if (xx == st.x && yy == st.y && b == st.b) break;
                    if (n >= max_len) break;
                }
                fs += n;
            }
            {
                ab::WalkState fw{st.x, st.y, st.b}, bw = fw;
                int nf = 0, ng = 0;
                bool ok = false;
                for (;;) {
                    ab::walk_forward(fw, ab::neighbours8(im, fw.x, fw.y)); nf++;
                    if (ab::same_state(fw, bw)) { ok = true; break; }
                    if (ab::is_smaller_trigger(im, fw, ab::neighbours8(im, fw.x, fw.y), st.key)) break;
                    ab::walk_backward(im, bw); ng++;
                    if (ab::same_state(fw, bw)) { ok = true; break; }
                    if (ab::is_smaller_trigger(im, bw, ab::neighbours8(im, bw.x, bw.y), st.key)) break;
                    if (nf + ng >= max_len) break;
                }
                bs += nf + ng;
                if (ok) kp += nf + ng;
            }
        }
    *fwd_steps = fs; *bidir_steps = bs; *kept_points = kp;
}
}

extern "C" int hc_solve_pnp_planar(const float* K, const float* D, const float* obj, const float* img, int N, double* rvec, double* tvec) {
    ab::Camera c = make_cam(K, D);
    return ab::solve_pnp_planar(c, obj, img, N, rvec, tvec) ? 1 : 0;
}
