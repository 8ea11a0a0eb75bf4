// Stage 2a: border extraction (replaces cv::findContours(RETR_LIST, CHAIN_APPROX_NONE),
// src/markerdetector.cpp:510-511, and the length filter at :517).  See ab_trace.cuh for the algorithm.
//   k_scan_starts: bitwise scan of the packed image for start candidates (one thread per 32-pixel word)
//   k_trace:       one thread per candidate walks its border cycle; the Suzuki start of every border with
//                  min_len < n < max_len re-walks it and writes the ordered points into the pool.
#pragma once
#include "ab_device.cuh"

namespace ab {

__global__ void k_scan_starts(Batch b) {
    const int ww = (b.W + 31) >> 5;
    const size_t total = (size_t)ww * b.H * b.B;
    const int lane = threadIdx.x & 31;
    for (size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; base < total;
         base += (size_t)gridDim.x * blockDim.x) {
        size_t i = base + lane;
        uint32_t outer = 0, hole = 0;
        int w = 0, y = 0, f = 0;
        if (i < total) {
            w = (int)(i % ww);
            y = (int)((i / ww) % b.H);
            f = (int)(i / ((size_t)ww * b.H));
            const uint32_t* row = b.bits + (size_t)f * b.bits_words + (size_t)(y + 1) * b.wpr + 1 + w;
            const uint32_t* up = row - b.wpr;
            uint32_t cur = row[0], west = (cur << 1) | (row[-1] >> 31);
            uint32_t u = up[0], uw = (u << 1) | (up[-1] >> 31), ue = (u >> 1) | (up[1] << 31);
            outer = cur & ~west & ~u & ~uw & ~ue;  // fg with W, N, NW, NE background
            hole = ~cur & west & u;                // bg with W and N foreground
        }
        int cnt = __popc(outer) + __popc(hole);
        // warp-aggregated reservation
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long wbase = 0;
        if (lane == 31 && tot > 0) wbase = atomicAdd(&b.cnt->n_starts, (unsigned long long)tot);
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
        if (tot == 0) continue;
        unsigned long long o = wbase + (unsigned long long)(incl - cnt);
        if (wbase + (unsigned long long)tot > b.cap_starts) {
            if (lane == 31) atomicOr(&b.cnt->err, ERR_STARTS_OVERFLOW);
            continue;
        }
        while (outer) {
            int j = __ffs((int)outer) - 1;
            outer &= outer - 1;
            b.starts[o++] = make_uint2((uint32_t)f, (uint32_t)(32 * w + j) | ((uint32_t)y << 16));
        }
        while (hole) {
            int j = __ffs((int)hole) - 1;
            hole &= hole - 1;
            b.starts[o++] = make_uint2((uint32_t)f | 0x80000000u, (uint32_t)(32 * w + j) | ((uint32_t)y << 16));
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
                walk_forward(fw, neighbours8(im, fw.x, fw.y));
                nf++;
                if (same_state(fw, bw)) {
                    closed = true;
                } else if (is_smaller_trigger(im, fw, neighbours8(im, fw.x, fw.y), st.key)) {
                    dead = true;
                } else {
                    walk_backward(im, bw);
                    ng++;
                    if (same_state(fw, bw)) closed = true;
                    else if (is_smaller_trigger(im, bw, neighbours8(im, bw.x, bw.y), st.key)) dead = true;
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
                walk_forward(fw, neighbours8(im, fw.x, fw.y));
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
