// Stage 2b: quad fitting and candidate filtering (src/markerdetector.cpp:513-627).
//   k_polygon:      one warp per kept contour: approxPolyDP(eps = 0.05 n, closed) -> exactly 4 vertices ->
//                   isContourConvex.  The farthest-point searches are warp reductions with OpenCV's
//                   "first strict maximum wins" tie rule; the slice stack lives in shared memory.
//   k_frame_filter: one CTA per frame: order quads as the reference sees them (reverse discovery order),
//                   orientation swap (:567-581), too-near removal (:586-613), emit candidates.
#pragma once
#include "ab_device.cuh"

namespace ab {

constexpr int DP_MAX_OUT = 16;    // a polygon that emits >= 16 vertices cannot clean up to 4
constexpr int DP_STACK = 40;

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        v = o > v ? o : v;
    }
    return v;
}

__global__ void __launch_bounds__(128) k_polygon(Batch b) {
    __shared__ int s_stack[4][DP_STACK][2];
    __shared__ int s_ox[4][DP_MAX_OUT], s_oy[4][DP_MAX_OUT];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned int nwarps = (gridDim.x * blockDim.x) >> 5;
    unsigned int ncont = b.cnt->n_contours;
    if (ncont > b.cap_contours) ncont = b.cap_contours;
    for (unsigned int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < ncont; ci += nwarps) {
        AB_BOUND((unsigned)ci < b.cap_contours);
        ContourRec rec = b.contours[ci];
        AB_BOUND((unsigned long long)rec.off + rec.n <= b.cap_pool);
        rec.frame &= CONTOUR_FRAME_MASK;  // bits 31/30 = border type / long contour, used by the emit kernels only
        const int n = (int)rec.n;
        if (n < 4) continue;
        const uint32_t* pts = b.pool + rec.off;
        const double eps = (double)n * 0.05;  // src/markerdetector.cpp:522
        const double E = eps * eps;
        // 1. three passes of "farthest point from the current start"
        int p0 = 0, rs = 0;
        long long maxd = 0;
        for (int it = 0; it < 3; it++) {
            p0 = (p0 + rs) % n;
            uint32_t sp = pts[p0];
            int sx = (int)(sp & 0xFFFFu), sy = (int)(sp >> 16);
            // per lane: first strict maximum (j ascends, so `>` keeps the smallest j of equal distances); coordinates are
            // < 2^14, the squared distance fits 32 bits.  The 64-bit key (dist << 32) | ~j is only built for the warp reduction.
            unsigned bestd = 0u, bestj = 0u;
            for (int j = 1 + lane; j < n; j += 32) {
                int q = p0 + j;
                if (q >= n) q -= n;
                uint32_t pp = pts[q];
                int dx = (int)(pp & 0xFFFFu) - sx, dy = (int)(pp >> 16) - sy;
                const unsigned d = (unsigned)(dx * dx + dy * dy);
                if (d > bestd) {
                    bestd = d;
                    bestj = (unsigned)j;
                }
            }
            unsigned long long best = bestd ? ((unsigned long long)bestd << 32) | (unsigned long long)(0xFFFFFFFFu - bestj) : 0ull;
            best = warp_max_u64(best);
            maxd = (long long)(best >> 32);
            if (best) rs = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFu));
        }
        int nout = 0, top = 0;
        bool reject = false;
        if ((double)maxd <= E) {
            reject = true;  // single-point polygon
        } else {
            int A = p0 % n, Bv = (rs + A) % n;
            if (lane == 0) {
                s_stack[wib][0][0] = Bv;
                s_stack[wib][0][1] = A;
                s_stack[wib][1][0] = A;
                s_stack[wib][1][1] = Bv;
            }
            top = 2;
        }
        __syncwarp();
        // 2. recursive subdivision with an explicit stack
        while (top > 0 && !reject) {
            top--;
            const int s = s_stack[wib][top][0], e = s_stack[wib][top][1];
            __syncwarp();
            const uint32_t ps = pts[s], pe = pts[e];
            const int sx = (int)(ps & 0xFFFFu), sy = (int)(ps >> 16), ex = (int)(pe & 0xFFFFu), ey = (int)(pe >> 16);
            int m = e - s;
            if (m < 0) m += n;
            m -= 1;  // interior points
            bool le = true;
            int split = 0;
            if (m > 0) {
                // seg_dist2 (ab_math.cuh) on integers.  Coordinates are < 2^14, so t, dd, the cross product c and the
                // end-point distances are exact in int32.  Points that project inside the segment have distance c*c/dd:
                // for one segment that is ordered exactly like |c| (distinct |c| differ by >= 2^-28 relative, far above
                // the rounding of the f64 quotient; equal |c| give equal quotients), so only the lane's best inside
                // point pays for the f64 division (it was 13 % of the kernel's instructions, ncu r1o).
                const int dxi = ex - sx, dyi = ey - sy, ddi = dxi * dxi + dyi * dyi;
                int best_c = 0, best_ci = 0x7FFFFFFF, best_o = 0, best_oi = 0x7FFFFFFF;
                for (int i = lane; i < m; i += 32) {
                    int q = s + 1 + i;
                    if (q >= n) q -= n;
                    const uint32_t pp = pts[q];
                    const int px = (int)(pp & 0xFFFFu), py = (int)(pp >> 16);
                    const int qx = px - sx, qy = py - sy;
                    const int tt = qx * dxi + qy * dyi;
                    if (tt < 0 || tt > ddi) {
                        const int fx = tt < 0 ? qx : px - ex, fy = tt < 0 ? qy : py - ey;
                        const int v = fx * fx + fy * fy;
                        if (v > best_o) {
                            best_o = v;
                            best_oi = i;
                        }
                    } else {
                        const int c = abs(qx * dyi - qy * dxi);
                        if (c > best_c) {
                            best_c = c;
                            best_ci = i;
                        }
                    }
                }
                // the lane's first strict maximum over both kinds of points
                const double d_in = best_c ? ((double)best_c * (double)best_c) / (double)ddi : 0.0, d_out = (double)best_o;
                const bool take_in = d_in > d_out || (d_in == d_out && best_ci < best_oi);
                const double dbest = take_in ? d_in : d_out;
                unsigned long long bd = dbest > 0 ? (unsigned long long)__double_as_longlong(dbest) : 0ull;  // bit pattern of a non-negative double orders like the value
                int bi = dbest > 0 ? (take_in ? best_ci : best_oi) : 0x7FFFFFFF;
#pragma unroll
                for (int dlt = 16; dlt > 0; dlt >>= 1) {
                    unsigned long long od = __shfl_xor_sync(0xFFFFFFFFu, bd, dlt);
                    int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, dlt);
                    if (od > bd || (od == bd && oi < bi)) {
                        bd = od;
                        bi = oi;
                    }
                }
                double md = __longlong_as_double((long long)bd);
                le = md <= E;
                split = s + 1 + bi;
                if (split >= n) split -= n;
            }
            if (le) {
                if (nout >= DP_MAX_OUT) {
                    reject = true;
                } else {
                    if (lane == 0) {
                        s_ox[wib][nout] = sx;
                        s_oy[wib][nout] = sy;
                    }
                    nout++;
                }
            } else {
                if (top + 2 > DP_STACK || nout + top + 2 > DP_MAX_OUT) {
                    reject = true;  // every stacked slice emits at least one vertex
                } else {
                    if (lane == 0) {
                        s_stack[wib][top][0] = split;
                        s_stack[wib][top][1] = e;
                        s_stack[wib][top + 1][0] = s;
                        s_stack[wib][top + 1][1] = split;
                    }
                    top += 2;
                }
            }
            __syncwarp();
        }
        if (reject || nout < 4) continue;
        // 3. clean-up, ==4 vertices, convexity (lane 0)
        if (lane == 0) {
            int ox[DP_MAX_OUT], oy[DP_MAX_OUT];
            for (int i = 0; i < nout; i++) {
                ox[i] = s_ox[wib][i];
                oy[i] = s_oy[wib][i];
            }
            int cnt = dp_cleanup(ox, oy, nout, E);
            if (cnt == 4 && is_convex4(ox, oy)) {
                // the reference's min-side>10 test (:542-552) indexes out of bounds and never rejects
                // reference order of the joined candidate list (:561-563): threshold image t ascending, then
                // reverse discovery order inside an image
                const unsigned rf = rec.frame / (unsigned)b.n_t, tt = rec.frame % (unsigned)b.n_t;
                unsigned int q = atomicAdd(&b.n_quads[rf], 1u);
                if (q >= (unsigned)b.cap_q) {
                    atomicOr(&b.cnt->err, ERR_QUADS_OVERFLOW);
                } else {
                    QuadRec qr;
                    for (int i = 0; i < 4; i++) {
                        qr.x[i] = (short)ox[i];
                        qr.y[i] = (short)oy[i];
                    }
                    qr.key = ((uint32_t)(b.n_t - 1 - (int)tt) << 28) | rec.key;
                    qr.contour = ci;
                    AB_BOUND((int)q < b.cap_q && rf * (unsigned)b.n_t < (unsigned)b.B);
                    b.quads[(size_t)rf * b.cap_q + q] = qr;
                }
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_frame_filter(Batch b) {
    __shared__ short s_xy[MAX_QUADS][8];
    __shared__ uint32_t s_key[MAX_QUADS];
    __shared__ uint32_t s_contour[MAX_QUADS];
    __shared__ float s_per[MAX_QUADS];
    __shared__ uint8_t s_swapped[MAX_QUADS];
    __shared__ uint8_t s_remove[MAX_QUADS];
    __shared__ int s_count;
    const int f = blockIdx.x, t = threadIdx.x;
    int nq = (int)b.n_quads[f];
    if (nq > b.cap_q) nq = b.cap_q;
    const QuadRec* q = b.quads + (size_t)f * b.cap_q;
    // rank sort by key descending (keys are unique per frame) = the order of contours2 in the reference
    for (int i = t; i < nq; i += blockDim.x) {
        uint32_t key = q[i].key;
        int rank = 0;
        for (int j = 0; j < nq; j++) rank += q[j].key > key;
        QuadRec r = q[i];
        float c[8];
        for (int k = 0; k < 4; k++) {
            c[2 * k] = (float)r.x[k];
            c[2 * k + 1] = (float)r.y[k];
        }
        // orientation (:567-581), f32 arithmetic as in the reference
        float d1x = c[2] - c[0], d1y = c[3] - c[1], d2x = c[4] - c[0], d2y = c[5] - c[1];
        float o = __fsub_rn(__fmul_rn(d1x, d2y), __fmul_rn(d1y, d2x));
        uint8_t sw = 0;
        if (o < 0.0f) {
            float tx = c[2], ty = c[3];
            c[2] = c[6];
            c[3] = c[7];
            c[6] = tx;
            c[7] = ty;
            sw = 1;
        }
        for (int k = 0; k < 8; k++) s_xy[rank][k] = (short)c[k];
        s_key[rank] = key;
        s_contour[rank] = r.contour;
        s_swapped[rank] = sw;
        s_per[rank] = perimeter4(c);
        s_remove[rank] = 0;
    }
    __syncthreads();
    // too-near pairs (:586-613): all four same-index corners closer than 6 px -> drop the smaller perimeter
    for (int i = t; i < nq; i += blockDim.x) {
        for (int j = i + 1; j < nq; j++) {
            bool near = true;
            for (int k = 0; k < 4 && near; k++) {
                int dx = s_xy[i][2 * k] - s_xy[j][2 * k], dy = s_xy[i][2 * k + 1] - s_xy[j][2 * k + 1];
                near = dx * dx + dy * dy < 36;  // norm < 6 on integer coordinates
            }
            if (near) {
                if (s_per[i] > s_per[j]) s_remove[j] = 1;
                else s_remove[i] = 1;
            }
        }
    }
    __syncthreads();
    if (t == 0) {
        int n = 0;
        bool overflow = false;
        CandRec* out = b.cands + (size_t)f * b.cap_c;
        for (int i = 0; i < nq; i++) {
            if (s_remove[i]) continue;
            if (n >= b.cap_c) {
                overflow = true;
                break;
            }
            CandRec cr;
            for (int k = 0; k < 8; k++) cr.c[k] = cr.refined[k] = (float)s_xy[i][k];
            cr.contour = s_contour[i];
            cr.swapped = s_swapped[i];
            cr.id = -1;
            cr.nrot = 0;
            out[n++] = cr;
        }
        if (overflow) atomicOr(&b.cnt->err, ERR_CANDS_OVERFLOW);
        b.n_cands[f] = (unsigned)n;
        atomicAdd(&b.cnt->n_quads_total, (unsigned long long)nq);
        atomicAdd(&b.cnt->n_cands_total, (unsigned long long)n);
        s_count = n;
    }
}

}  // namespace ab
