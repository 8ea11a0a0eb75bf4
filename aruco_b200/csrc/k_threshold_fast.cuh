// Fast path of the adaptive threshold for the common odd block sizes (compile-time K): same arithmetic as
// k_threshold_adaptive (k_threshold.cuh), restructured to cut shared-memory traffic (the generic kernel is
// LSU/shared-memory bound: ncu r1a showed memory pipes 93 % busy at 12.7 % DRAM throughput):
//   * the K most recent source rows of a thread's 4 columns live in REGISTERS (ring unrolled K times),
//   * vertical sums are published as 4 x u16 in one 8-byte store, double buffered -> one barrier per row,
//   * the horizontal window is read back with 8-byte loads (4 + 2*R4 values = 3 loads for K = 7),
//   * 128 output threads are word aligned (thread t owns columns X0+4t), halo columns are computed by a few
//     extra threads, so the 1-bit packed copy is assembled with three warp shuffles and no second barrier.
#pragma once
#include "k_threshold.cuh"

namespace ab {

constexpr int THR_OUT_THREADS = 128;            // output threads per CTA
constexpr int THR_TWO = 4 * THR_OUT_THREADS;    // 512 output columns per CTA
constexpr int THR_RH = 128;                     // output rows per CTA

template <int K>
__global__ void __launch_bounds__(160) k_threshold_fast(ThrArgs a) {
    constexpr int R = K / 2, R4 = (R + 3) & ~3, HT = R4 / 4, NV = 4 + 2 * R4, CSW = THR_TWO + 2 * R4, K2 = K * K;
    __shared__ __align__(16) unsigned short cs[2][CSW];
    const int t = threadIdx.x;
    if (t >= THR_OUT_THREADS + 2 * HT) return;  // spare lanes of the halo warp
    const int X0 = blockIdx.x * THR_TWO, y0 = blockIdx.y * THR_RH, f = blockIdx.z;
    const bool is_out = t < THR_OUT_THREADS;
    int c0, ci;  // first column of this thread, its index in the column-sum row
    if (is_out) {
        c0 = X0 + 4 * t;
        ci = R4 + 4 * t;
    } else if (t < THR_OUT_THREADS + HT) {
        ci = 4 * (t - THR_OUT_THREADS);
        c0 = X0 - R4 + ci;
    } else {
        ci = R4 + THR_TWO + 4 * (t - THR_OUT_THREADS - HT);
        c0 = X0 - R4 + ci;
    }
    const uint8_t* src = a.grey + (size_t)f * a.grey_frame;
    uint8_t* dst = a.thres + (size_t)f * a.W * a.H;
    uint32_t* bits = a.bits + (size_t)f * a.bits_words;
    const bool fast = a.aligned4 && c0 >= 0 && c0 + 3 < a.W;
    const bool live = c0 < a.W + R4;  // columns far right of the image are never needed
    int xc0 = min(max(c0, 0), a.W - 1), xc1 = min(max(c0 + 1, 0), a.W - 1), xc2 = min(max(c0 + 2, 0), a.W - 1),
        xc3 = min(max(c0 + 3, 0), a.W - 1);
    const int yEnd = min(y0 + THR_RH, a.H);
    const int nrows = (yEnd - y0) + 2 * R;
    const bool store_vec = (c0 + 3 < a.W) && ((a.W & 3) == 0);
    uint32_t ring[K];
#pragma unroll
    for (int j = 0; j < K; j++) ring[j] = 0u;
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int buf = 0;
    for (int base = 0; base < nrows; base += K) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int i = base + j;
            if (i < nrows) {
                uint32_t p = 0;
                if (live) {
                    int yy = min(max(y0 - R + i, 0), a.H - 1);
                    const uint8_t* rowp = src + (size_t)yy * a.grey_row;
                    if (fast)
                        p = *reinterpret_cast<const uint32_t*>(rowp + c0);
                    else
                        p = (uint32_t)rowp[xc0] | ((uint32_t)rowp[xc1] << 8) | ((uint32_t)rowp[xc2] << 16) | ((uint32_t)rowp[xc3] << 24);
                }
                const uint32_t old = ring[j];
                ring[j] = p;
                s0 += (int)(p & 255u) - (int)(old & 255u);
                s1 += (int)((p >> 8) & 255u) - (int)((old >> 8) & 255u);
                s2 += (int)((p >> 16) & 255u) - (int)((old >> 16) & 255u);
                s3 += (int)(p >> 24) - (int)(old >> 24);
                if (i >= 2 * R) {
                    const int yo = y0 + i - 2 * R;
                    *reinterpret_cast<uint2*>(&cs[buf][ci]) = make_uint2((uint32_t)s0 | ((uint32_t)s1 << 16), (uint32_t)s2 | ((uint32_t)s3 << 16));
                    __syncthreads();
                    if (is_out) {
                        int v[NV];
                        const uint2* wp = reinterpret_cast<const uint2*>(&cs[buf][ci - R4]);
#pragma unroll
                        for (int q = 0; q < NV / 4; q++) {
                            uint2 u = wp[q];
                            v[4 * q] = (int)(u.x & 0xFFFFu);
                            v[4 * q + 1] = (int)(u.x >> 16);
                            v[4 * q + 2] = (int)(u.y & 0xFFFFu);
                            v[4 * q + 3] = (int)(u.y >> 16);
                        }
                        int S = 0;
#pragma unroll
                        for (int d = R4 - R; d <= R4 + R; d++) S += v[d];
                        const uint32_t c = ring[(j + K - R) % K];  // centre row
                        uint32_t nibble = 0, outb = 0;
#pragma unroll
                        for (int jj = 0; jj < 4; jj++) {
                            int T = (int)((c >> (8 * jj)) & 255u) + a.idelta;
                            bool on = (2 * S + K2 >= 2 * K2 * T) && (c0 + jj < a.W);
                            if (on) {
                                nibble |= 1u << jj;
                                outb |= 255u << (8 * jj);
                            }
                            if (jj < 3) S += v[R4 + jj + 1 + R] - v[R4 + jj - R];
                        }
                        if (c0 < a.W) {
                            uint8_t* orow = dst + (size_t)yo * a.W + c0;
                            if (store_vec) {
                                *reinterpret_cast<uint32_t*>(orow) = outb;
                            } else {
                                for (int jj = 0; jj < 4; jj++)
                                    if (c0 + jj < a.W) orow[jj] = (uint8_t)(outb >> (8 * jj));
                            }
                        }
                        uint32_t word = nibble << (4 * (t & 7));
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
                        if ((t & 7) == 0 && c0 < a.W) bits[(size_t)(yo + 1) * a.wpr + BIT_PAD + (X0 >> 5) + (t >> 3)] = word;
                    }
                    buf ^= 1;
                }
            }
        }
    }
}

// host-side dispatch: returns false when K has no compiled fast path
inline bool launch_threshold_fast(const ThrArgs& a, int B, cudaStream_t st) {
    dim3 grid((a.W + THR_TWO - 1) / THR_TWO, (a.H + THR_RH - 1) / THR_RH, B);
    switch (a.k) {
#define AB_THR_CASE(KK) \
    case KK:            \
        k_threshold_fast<KK><<<grid, 160, 0, st>>>(a); \
        return true;
        AB_THR_CASE(3)
        AB_THR_CASE(5)
        AB_THR_CASE(7)
        AB_THR_CASE(9)
        AB_THR_CASE(11)
        AB_THR_CASE(13)
        AB_THR_CASE(15)
        AB_THR_CASE(17)
        AB_THR_CASE(19)
        AB_THR_CASE(21)
#undef AB_THR_CASE
        default:
            return false;
    }
}

}  // namespace ab
