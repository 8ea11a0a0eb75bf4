// Stage 2a: border extraction (replaces cv::findContours(RETR_LIST, CHAIN_APPROX_NONE),
// src/markerdetector.cpp:510-511, and the length filter at :517).  See ab_trace.cuh for the algorithm.
//   k_scan_starts: bitwise scan of the packed image for start candidates (one thread per 32-pixel word)
//   k_trace:       one thread per candidate walks its border cycle; the Suzuki start of every border with
//                  min_len < n < max_len re-walks it and writes the ordered points into the pool.
#pragma once
#include "ab_device.cuh"

namespace ab {

// one thread per 128 pixels (four packed words, 128-bit loads); candidates are appended with one atomic per warp
__global__ void __launch_bounds__(256) k_scan_starts(Batch b) {
    const int ww = (b.W + 31) >> 5;
    const unsigned nq = (unsigned)(ww + 3) >> 2;  // quads per row
    const unsigned long long total = (unsigned long long)nq * b.H * b.B;
    const int lane = threadIdx.x & 31;
    for (unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; base < total;
         base += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long i = base + lane;
        uint32_t outer[4] = {0, 0, 0, 0}, hole[4] = {0, 0, 0, 0};
        int q = 0, y = 0, f = 0, cnt = 0;
        if (i < total) {
            q = (int)(i % nq);
            unsigned long long r = i / nq;
            y = (int)(r % (unsigned)b.H);
            f = (int)(r / (unsigned)b.H);
            const uint32_t* row = b.bits + (size_t)f * b.bits_words + (size_t)(y + 1) * b.wpr + BIT_PAD + 4 * q;
            const uint32_t* up = row - b.wpr;
            const uint4 c4 = *reinterpret_cast<const uint4*>(row);
            const uint4 u4 = *reinterpret_cast<const uint4*>(up);
            const uint32_t c[6] = {row[-1], c4.x, c4.y, c4.z, c4.w, 0u};
            const uint32_t u[6] = {up[-1], u4.x, u4.y, u4.z, u4.w, up[4]};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t cur = c[k + 1], west = (cur << 1) | (c[k] >> 31);
                uint32_t uu = u[k + 1], uw = (uu << 1) | (u[k] >> 31), ue = (uu >> 1) | (u[k + 2] << 31);
                outer[k] = cur & ~west & ~uu & ~uw & ~ue;  // fg with W, N, NW, NE background
                hole[k] = ~cur & west & uu;                // bg with W and N foreground
                cnt += __popc(outer[k]) + __popc(hole[k]);
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, cnt != 0) == 0) continue;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long wbase = 0;
        if (lane == 31) wbase = atomicAdd(&b.cnt->n_starts, (unsigned long long)tot);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
        if (wbase + (unsigned long long)tot > b.cap_starts) {
            if (lane == 31) atomicOr(&b.cnt->err, ERR_STARTS_OVERFLOW);
            continue;
        }
        unsigned long long o = wbase + (unsigned long long)(incl - cnt);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t m = outer[k];
            const uint32_t xb = (uint32_t)(128 * q + 32 * k);
            while (m) {
                int j = __ffs((int)m) - 1;
                m &= m - 1;
                b.starts[o++] = make_uint2((uint32_t)f, (xb + j) | ((uint32_t)y << 16));
            }
            m = hole[k];
            while (m) {
                int j = __ffs((int)m) - 1;
                m &= m - 1;
                b.starts[o++] = make_uint2((uint32_t)f | 0x80000000u, (xb + j) | ((uint32_t)y << 16));
            }
        }
    }
}

// Lane-scheduled walker.  Every lane owns one start candidate at a time and is refilled from a global work
// counter as soon as it finishes, so a warp never idles behind its longest contour (border lengths range
// from 2 to max_len).  Phase 1 = bidirectional search for a smaller start on the same cycle
// (find_start_bidir in ab_trace.cuh, executed STEPS at a time); phase 2 = the Suzuki start of a kept border
// re-walks it forwards and writes the ordered points into the pool.
__global__ void __launch_bounds__(128) k_trace(Batch b) {
    constexpr int STEPS = 8;
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    unsigned long long n_starts = b.cnt->n_starts;
    if (n_starts > b.cap_starts) return;  // overflow already flagged by k_scan_starts; the list has holes
    int phase = 0;  // 0 idle, 1 search, 2 emit
    bool exhausted = false;
    BitImage im = b.bit_image(0);
    TraceStart st{0, 0, 0, 0};
    WalkState fw{0, 0, 0}, bw{0, 0, 0};
    uint32_t nb_fw = 0;
    int nf = 0, ng = 0, len = 0, frame = 0;
    uint32_t* out = nullptr;
    unsigned int ci = 0;
    unsigned long long off = 0;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, phase == 0 && !exhausted);
        if (idle) {
            unsigned base = 0;
            int leader = __ffs((int)idle) - 1;
            if (lane == leader) base = atomicAdd(&b.cnt->trace_work, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (phase == 0 && !exhausted) {
                unsigned long long i = (unsigned long long)base + (unsigned)__popc(idle & ((1u << lane) - 1u));
                if (i >= n_starts) {
                    exhausted = true;
                } else {
                    uint2 rec = b.starts[i];
                    frame = (int)(rec.x & 0x7FFFFFFFu);
                    im = b.bit_image(frame);
                    if (make_start(im, (int)(rec.x >> 31), (int)(rec.y & 0xFFFFu), (int)(rec.y >> 16), st)) {
                        fw = WalkState{st.x, st.y, st.b};
                        bw = fw;
                        nb_fw = neighbours8(im, fw.x, fw.y);
                        nf = ng = 0;
                        phase = 1;
                    }  // else: isolated pixel, a 1-point contour that is never kept
                }
            }
        }
        if (__ballot_sync(FULL, phase != 0) == 0) {
            if (__ballot_sync(FULL, !exhausted) == 0) break;
            continue;
        }
        if (phase == 1) {
            for (int r = 0; r < STEPS; r++) {
                bool closed = false, dead = false;
                walk_forward(fw, nb_fw);
                nf++;
                if (same_state(fw, bw)) {
                    closed = true;
                } else if (is_smaller_trigger(im, fw, nb_fw = neighbours8(im, fw.x, fw.y), st.key)) {
                    dead = true;
                } else {
                    uint32_t nb_bw = walk_backward(im, bw);
                    ng++;
                    if (same_state(fw, bw)) closed = true;
                    else if (is_smaller_trigger(im, bw, nb_bw, st.key)) dead = true;
                    else if (nf + ng >= b.max_len) dead = true;  // too long: dropped by :517 anyway
                }
                if (closed) {
                    len = nf + ng;
                    if (len <= b.min_len || len >= b.max_len) {  // src/markerdetector.cpp:517
                        phase = 0;
                        break;
                    }
                    ci = atomicAdd(&b.cnt->n_contours, 1u);
                    off = atomicAdd(&b.cnt->pool_used, (unsigned long long)len);
                    if (ci >= b.cap_contours) {
                        atomicOr(&b.cnt->err, ERR_CONTOURS_OVERFLOW);
                        phase = 0;
                    } else if (off + (unsigned long long)len > b.cap_pool) {
                        atomicOr(&b.cnt->err, ERR_POOL_OVERFLOW);
                        b.contours[ci] = ContourRec{(uint32_t)frame, 0u, 0u, (uint32_t)st.key};
                        phase = 0;
                    } else {
                        out = b.pool + off;
                        fw = WalkState{st.x, st.y, st.b};
                        nb_fw = neighbours8(im, fw.x, fw.y);
                        nf = 0;
                        phase = 2;
                    }
                    break;
                }
                if (dead) {
                    phase = 0;
                    break;
                }
            }
        } else if (phase == 2) {
            for (int r = 0; r < 2 * STEPS; r++) {
                out[nf] = (uint32_t)fw.x | ((uint32_t)fw.y << 16);
                walk_forward(fw, nb_fw);
                nb_fw = neighbours8(im, fw.x, fw.y);
                if (++nf == len) {
                    b.contours[ci] = ContourRec{(uint32_t)frame, (uint32_t)off, (uint32_t)len, (uint32_t)st.key};
                    phase = 0;
                    break;
                }
            }
        }
    }
}

}  // namespace ab
