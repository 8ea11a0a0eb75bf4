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

__global__ void k_trace(Batch b) {
    const int lane = threadIdx.x & 31;
    unsigned long long n_starts = b.cnt->n_starts;
    if (n_starts > b.cap_starts) n_starts = b.cap_starts;
    for (;;) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(&b.cnt->trace_work, 32u);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if ((unsigned long long)base >= n_starts) break;
        unsigned long long i = (unsigned long long)base + lane;
        if (i >= n_starts) continue;
        uint2 rec = b.starts[i];
        int f = (int)(rec.x & 0x7FFFFFFFu), type = (int)(rec.x >> 31);
        int x = (int)(rec.y & 0xFFFFu), y = (int)(rec.y >> 16);
        BitImage im = b.bit_image(f);
        TraceStart st;
        if (!make_start(im, type, x, y, st)) continue;  // isolated pixel: a 1-point contour, never kept
        int len = 0;
        if (trace_cycle(im, st, b.max_len, &len, nullptr) != TRACE_OK) continue;
        if (len <= b.min_len || len >= b.max_len) continue;  // src/markerdetector.cpp:517
        unsigned int ci = atomicAdd(&b.cnt->n_contours, 1u);
        unsigned long long off = atomicAdd(&b.cnt->pool_used, (unsigned long long)len);
        if (ci >= b.cap_contours) {
            atomicOr(&b.cnt->err, ERR_CONTOURS_OVERFLOW);
            continue;
        }
        if (off + (unsigned long long)len > b.cap_pool) {
            atomicOr(&b.cnt->err, ERR_POOL_OVERFLOW);
            b.contours[ci] = ContourRec{(uint32_t)f, 0u, 0u, (uint32_t)st.key};
            continue;
        }
        trace_cycle(im, st, b.max_len + 1, &len, b.pool + off);
        b.contours[ci] = ContourRec{(uint32_t)f, (uint32_t)off, (uint32_t)len, (uint32_t)st.key};
    }
}

}  // namespace ab
