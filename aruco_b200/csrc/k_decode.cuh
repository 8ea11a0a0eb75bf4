// Stage 3: per-candidate identification (src/markerdetector.cpp:350-356).
// One WARP per candidate (4 candidates per CTA): getPerspectiveTransform + warpPerspective(INTER_NEAREST) into
// shared memory (MarkerDetector::warp, :684-697), 256-bin histogram -> Otsu (cv::threshold BINARY|OTSU),
// majority vote per cell, then FiducidalMarkers::detect (src/arucofidmarkers.cpp:438-452) or
// HighlyReliableMarkers::detect (src/highlyreliablemarkers.cpp:332-383).  The serial f64 pieces (8x8 LU, the
// Otsu recurrence -- both must follow OpenCV's operation order to stay bit-exact) run on lane 0; giving every
// candidate its own warp keeps ~48 of those chains in flight per SM instead of 6 with a CTA per candidate.
// The canonical image is also written out (getCandidates / host-callback decoders need it).
#pragma once
#include "ab_device.cuh"

namespace ab {

constexpr int MAX_WARP_SIZE = 128;  // S <= 128
constexpr int MAX_CELLS = 100;      // (n+2)^2 with n <= 8
constexpr int DECODE_WARPS = 4;
constexpr int DECODE_LIST = 1024;   // per-warp list of pixels that need the exact f64 coordinate

__device__ __forceinline__ int hrm_decode(const HrmDict& D, const uint8_t* cells, int ncell, int* nrot, int lane) {
    // cells: (n+2)^2 majority bits; HRM ignores the border cells (highlyreliablemarkers.cpp:345)
    const int n = D.n;
    uint8_t code[64];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) code[y * n + x] = cells[(y + 1) * ncell + x + 1];
    uint64_t bits[4];
    uint32_t ids[4];
    hrm_rotations(code, n, bits, ids);
    // 1. exact lookup of the four folded ids in the balanced tree (findId, :483-496)
    for (int r = 0; r < 4; r++) {
        int pos = D.root;
        while (pos != -1) {
            uint32_t pid = D.ord_ids[pos];
            if (pid == ids[r]) {
                *nrot = r;
                return D.ord_pos[pos];
            }
            pos = pid < ids[r] ? D.tree[2 * pos + 1] : D.tree[2 * pos];
        }
    }
    // 2. error correction: min over dictionary of min over rotations (Dictionary::distance, :277-289),
    //    first strict minimum wins in (marker, rotation) order
    unsigned best = 0xFFFFFFFFu;  // (dist << 16) | (marker << 2) | rot
    for (int i = lane; i < D.count; i += 32) {
        uint64_t d0 = D.bits[i];
        unsigned res = (unsigned)(n * n), mr = 0;
        for (unsigned r = 0; r < 4; r++) {
            unsigned hd = (unsigned)__popcll(d0 ^ bits[r]);
            if (hd < res) {
                res = hd;
                mr = r;
            }
        }
        if (res < (unsigned)(n * n)) {
            unsigned key = (res << 16) | ((unsigned)i << 2) | mr;
            best = min(best, key);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
    if (best != 0xFFFFFFFFu && (int)(best >> 16) <= D.correction) {
        *nrot = (int)(best & 3u);
        return (int)((best >> 2) & 0x3FFFu);
    }
    *nrot = 0;
    return -1;
}

// cv::threshold(OTSU) threshold, same arithmetic per bin as otsu_threshold() in ab_math.cuh.  Only
// the recurrence of (q1, mu1) is inherently sequential (its rounding sequence must be reproduced); it runs on
// lane 0 in chunks of 32 bins, the other lanes then evaluate mu2 / sigma of "their" bin, and the first strict
// maximum is found by a warp reduction.  Leading empty bins and the bins after q1 saturates are skipped: both
// are exact no-ops of the sequential loop.  `scratch` = 64 doubles of shared memory.
__device__ __forceinline__ int otsu_threshold_warp(const int* h, int N, double* scratch, int lane) {
    const unsigned FULL = 0xFFFFFFFFu;
    // mu = sum i*h[i] / N : integer valued partial sums are exact in f64, so the order is free
    long long part = 0;
    for (int i = lane; i < 256; i += 32) part += (long long)i * h[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(FULL, part, d);
    const double scale = 1. / N;
    const double mu = (double)part * scale;
    double* s_q = scratch;
    double* s_m = scratch + 32;
    double mu1 = 0, q1 = 0;  // lane 0 only
    bool done = false;
    double best_sigma = 0;
    int best_i = 0;
    for (int c = 0; c < 8; c++) {
        if (lane == 0) {
            for (int k = 0; k < 32; k++) {
                const int i = 32 * c + k;
                double q = -1.0, m = 0;
                if (!done) {
                    const double p_i = h[i] * scale;
                    mu1 *= q1;
                    q1 += p_i;
                    const double q2 = 1. - q1;
                    const double mn = q1 < q2 ? q1 : q2, mx = q1 < q2 ? q2 : q1;
                    if (mx > 1. - FLT_EPSILON && q1 > q2) {
                        done = true;  // q1 only grows: every later bin is skipped by the same test
                    } else if (!(mn < FLT_EPSILON || mx > 1. - FLT_EPSILON)) {
                        mu1 = (mu1 + i * p_i) / q1;
                        q = q1;
                        m = mu1;
                    }
                }
                s_q[k] = q;
                s_m[k] = m;
            }
        }
        __syncwarp();
        const double q = s_q[lane], m = s_m[lane];
        if (q >= 0) {
            const double q2 = 1. - q;
            const double mu2 = (mu - q * m) / q2;
            const double sigma = q * q2 * (m - mu2) * (m - mu2);
            if (sigma > best_sigma) {  // within a lane bins come in increasing order: first strict maximum
                best_sigma = sigma;
                best_i = 32 * c + lane;
            }
        }
        __syncwarp();
    }
    // across lanes: larger sigma wins, ties go to the smaller bin index
    unsigned long long sb = (unsigned long long)__double_as_longlong(best_sigma);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long os = __shfl_xor_sync(FULL, sb, d);
        int oi = __shfl_xor_sync(FULL, best_i, d);
        if (os > sb || (os == sb && oi < best_i)) {
            sb = os;
            best_i = oi;
        }
    }
    return sb ? best_i : 0;
}

inline size_t decode_smem_per_warp(int S) {
    return (((size_t)S * S + 15) & ~(size_t)15) + 256 * sizeof(int) + MAX_CELLS * sizeof(int) + 112 + 9 * sizeof(double) + 8 +
           DECODE_LIST * sizeof(unsigned short);
}

// mode 0: warp + decode; mode 1: warp only (host-callback decoder).  grid = (ceil(cap_c / DECODE_WARPS), B)
__global__ void __launch_bounds__(32 * DECODE_WARPS) k_decode(Batch b, int mode) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ci = blockIdx.x * DECODE_WARPS + wib;
    if (ci >= (int)b.n_cands[f] || ci >= b.cap_c) return;  // whole warp leaves; only warp-level sync below
    const int S = b.S;
    const size_t img_bytes = ((size_t)S * S + 15) & ~(size_t)15;
    const size_t per_warp = img_bytes + 256 * sizeof(int) + MAX_CELLS * sizeof(int) + 112 + 9 * sizeof(double) + 8 + DECODE_LIST * sizeof(unsigned short);
    unsigned char* base = s_raw + (size_t)wib * per_warp;
    uint8_t* s_img = base;
    int* s_hist = reinterpret_cast<int*>(base + img_bytes);
    int* s_cnt = s_hist + 256;
    uint8_t* s_cells = reinterpret_cast<uint8_t*>(s_cnt + MAX_CELLS);
    double* s_Mi = reinterpret_cast<double*>(base + img_bytes + 256 * sizeof(int) + MAX_CELLS * sizeof(int) + 112);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_Mi + 10);
    CandRec* cand = b.cands + (size_t)f * b.cap_c + ci;
    int ok_i = 0;
    if (lane == 0) {
        float dst[8] = {0.f, 0.f, (float)(S - 1), 0.f, (float)(S - 1), (float)(S - 1), 0.f, (float)(S - 1)};
        double M[9], Mi[9];
        bool ok = perspective_transform(cand->c, dst, M) && invert3(M, Mi);
        ok_i = ok;
        if (ok)
            for (int i = 0; i < 9; i++) s_Mi[i] = Mi[i];
    }
    for (int i = lane; i < 256; i += 32) s_hist[i] = 0;
    for (int i = lane; i < MAX_CELLS; i += 32) s_cnt[i] = 0;
    ok_i = __shfl_sync(0xFFFFFFFFu, ok_i, 0);
    __syncwarp();
    const uint8_t* grey = b.grey + (size_t)f * b.grey_frame;
    uint8_t* canon = b.canon + ((size_t)f * b.cap_c + ci) * (size_t)(S * S);
    const int bw = warp_block_width(S);
    const bool ok = ok_i != 0;
    // Filtered arithmetic: OpenCV's source coordinate is rint() of an f64 expression.  An f32 evaluation
    // (error << 0.05 px for |coord| < 32768) decides every pixel whose coordinate is not within 0.05 px of a
    // rounding boundary -- the integer it rounds to is then provably the same; the remaining ~10 % are
    // collected (ballot compaction) and redone densely with the exact f64 sequence.  FP64 issue rate, not
    // memory, was the limiter of this kernel (ncu r1d).
    float Mf[9];
#pragma unroll
    for (int q = 0; q < 9; q++) Mf[q] = ok ? (float)s_Mi[q] : 0.f;
    int nunc = 0;
    int px = lane % S, py = lane / S;  // pixel (x,y) of index i = i0 + lane, advanced incrementally
    for (int i0 = 0; i0 < S * S; i0 += 32) {
        const int i = i0 + lane;
        bool unc = false;
        const int x = px, y = py;
        px += 32;
        while (px >= S) {
            px -= S;
            py++;
        }
        if (i < S * S) {
            if (!ok) {
                s_img[i] = 0;
                canon[i] = 0;
            } else {
                const float fxp = (float)x, fyp = (float)y;
                const float den = Mf[6] * fxp + Mf[7] * fyp + Mf[8];
                const float iden = __frcp_rn(den);
                const float fx = (Mf[0] * fxp + Mf[1] * fyp + Mf[2]) * iden, fy = (Mf[3] * fxp + Mf[4] * fyp + Mf[5]) * iden;
                const float rx = rintf(fx), ry = rintf(fy);
                // forward error bound of the f32 evaluation (unit roundoff 6e-8; 1e-6 leaves > 3x slack), doubled
                const float iad = fabsf(iden);
                const float ex = 2e-6f * ((fabsf(Mf[0] * fxp) + fabsf(Mf[1] * fyp) + fabsf(Mf[2])) * iad + fabsf(fx)) + 2e-4f;
                const float ey = 2e-6f * ((fabsf(Mf[3] * fxp) + fabsf(Mf[4] * fyp) + fabsf(Mf[5])) * iad + fabsf(fy)) + 2e-4f;
                unc = !(fabsf(fx) < 32768.f && fabsf(fy) < 32768.f) || !(0.5f - fabsf(fx - rx) > ex) || !(0.5f - fabsf(fy - ry) > ey);
                if (!unc) {
                    const int sx = (int)rx, sy = (int)ry;
                    uint8_t v = 0;
                    if (sx >= 0 && sy >= 0 && sx < b.W && sy < b.H) v = grey[(size_t)sy * b.grey_row + sx];
                    s_img[i] = v;
                    canon[i] = v;
                }
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, unc);
        if (m) {
            const int pos = nunc + __popc(m & ((1u << lane) - 1u));
            if (unc) {
                if (pos < DECODE_LIST) {
                    s_list[pos] = (unsigned short)i;
                } else {  // list full (cannot happen for guard 0.05 unless the map is degenerate): do it now
                    const int y = i / S, x = i - y * S;
                    int sx, sy;
                    warp_src_coord(s_Mi, x, y, bw, &sx, &sy);
                    uint8_t v = 0;
                    if (sx >= 0 && sy >= 0 && sx < b.W && sy < b.H) v = grey[(size_t)sy * b.grey_row + sx];
                    s_img[i] = v;
                    canon[i] = v;
                }
            }
            nunc += __popc(m);
        }
    }
    __syncwarp();
    const int n2 = min(nunc, DECODE_LIST);
    for (int k = lane; k < n2; k += 32) {
        const int i = (int)s_list[k];
        const int y = i / S, x = i - y * S;
        int sx, sy;
        warp_src_coord(s_Mi, x, y, bw, &sx, &sy);
        uint8_t v = 0;
        if (sx >= 0 && sy >= 0 && sx < b.W && sy < b.H) v = grey[(size_t)sy * b.grey_row + sx];
        s_img[i] = v;
        canon[i] = v;
    }
    __syncwarp();
    if (mode == 1) return;
    // 256-bin histogram: lanes holding the same grey level are matched, one of them adds the group size
    // (shared-memory atomics on a bimodal image serialise almost completely: 18 % of the stall samples in r1e)
    for (int i0 = 0; i0 < S * S; i0 += 32) {
        const int i = i0 + lane;
        const unsigned v = i < S * S ? (unsigned)s_img[i] : 0x100u + (unsigned)lane;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, v);
        if (i < S * S && lane == __ffs((int)peers) - 1) s_hist[v] += __popc(peers);
    }
    __syncwarp();
    const int thr = otsu_threshold_warp(s_hist, S * S, reinterpret_cast<double*>(s_list), lane);
    const int ncell = (b.decoder == AB_DECODER_HRM) ? b.dict.n + 2 : 7;
    const int cell = S / ncell;
    const int span = cell * ncell;
    (void)span;
    for (int cidx = lane; cidx < ncell * ncell; cidx += 32) {  // one lane per cell: no atomics
        const int cy = cidx / ncell, cx = cidx - cy * ncell;
        int cnt = 0;
        for (int yy = 0; yy < cell; yy++) {
            const uint8_t* rowp = s_img + (cy * cell + yy) * S + cx * cell;
            for (int xx = 0; xx < cell; xx++) cnt += rowp[xx] > thr;
        }
        s_cells[cidx] = cnt > (cell * cell) / 2;
    }
    __syncwarp();
    if (b.decoder == AB_DECODER_HRM) {
        int nrot = 0;
        int id = hrm_decode(b.dict, s_cells, ncell, &nrot, lane);
        if (lane == 0) {
            cand->id = id;
            cand->nrot = nrot;
        }
    } else if (lane == 0) {
        int nrot = 0;
        int id = fid_decode(s_cells, &nrot);
        cand->id = id;
        cand->nrot = nrot;
    }
}

// public worker MarkerDetector::warp for one quad (grid 1)
__global__ void k_warp_single(const uint8_t* grey, int W, int H, size_t row, const float* quad, int S, uint8_t* out) {
    __shared__ double s_Mi[9];
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        float dst[8] = {0.f, 0.f, (float)(S - 1), 0.f, (float)(S - 1), (float)(S - 1), 0.f, (float)(S - 1)};
        float q[8];
        for (int i = 0; i < 8; i++) q[i] = quad[i];
        double M[9], Mi[9];
        bool ok = perspective_transform(q, dst, M) && invert3(M, Mi);
        s_ok = ok;
        if (ok)
            for (int i = 0; i < 9; i++) s_Mi[i] = Mi[i];
    }
    __syncthreads();
    const int bw = warp_block_width(S);
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        int y = i / S, x = i - y * S;
        uint8_t v = 0;
        if (s_ok) {
            int sx, sy;
            warp_src_coord(s_Mi, x, y, bw, &sx, &sy);
            if (sx >= 0 && sy >= 0 && sx < W && sy < H) v = grey[(size_t)sy * row + sx];
        }
        out[i] = v;
    }
}

}  // namespace ab
