// Fast path of the adaptive threshold for the common odd block sizes (compile-time K): same arithmetic as
// k_threshold_adaptive (k_threshold.cuh), restructured around instruction count (ncu r1a/r1d: the kernel is
// issue bound -- 68 % issue slots busy at 13 % DRAM throughput -- so every removed instruction is time):
//   * the K most recent source rows of a thread's 4 columns live in REGISTERS (ring unrolled K times),
//   * vertical sums are kept as packed 2 x u16 (even / odd columns) and updated with plain 32-bit adds,
//   * they are published as 4 x u16 in one 8-byte store, double buffered -> one barrier per row,
//   * the horizontal window is read back with 8-byte loads and, for K <= 15 (sums < 65536), summed as packed
//     u16 pairs: 15 integer ops give the four window sums of a thread for K = 7,
//   * mean >= T is tested as  S >= K^2 (src + idelta) - (K^2 - 1)/2  (no division, one IMAD per pixel),
//   * 128 output threads are word aligned (thread t owns columns X0+4t), halo columns are computed by a few
//     extra threads, so the 1-bit packed copy is assembled with three warp shuffles and no second barrier.
#pragma once
#include "k_threshold.cuh"

namespace ab {

constexpr int THR_OUT_THREADS = 128;            // output threads per CTA
constexpr int THR_TWO = 4 * THR_OUT_THREADS;    // 512 output columns per CTA
constexpr int THR_RH = 128;                     // output rows per CTA

// (v[i] | v[i+1] << 16) from the packed words w[] (w[q] = v[2q] | v[2q+1] << 16)
template <int I>
__device__ __forceinline__ uint32_t thr_pair(const uint32_t* w) {
    if constexpr ((I & 1) == 0) return w[I / 2];
    else return __byte_perm(w[(I - 1) / 2], w[(I + 1) / 2], 0x5432);
}
template <int I, int END>
__device__ __forceinline__ uint32_t thr_pair_sum(const uint32_t* w) {
    if constexpr (I > END) return 0u;
    else return thr_pair<I>(w) + thr_pair_sum<I + 1, END>(w);
}

template <int K>
__global__ void __launch_bounds__(160) k_threshold_fast(ThrArgs a) {
    constexpr int R = K / 2, R4 = (R + 3) & ~3, HT = R4 / 4, NV = 4 + 2 * R4, NW = NV / 2, CSW = THR_TWO + 2 * R4, K2 = K * K;
    constexpr bool PACKED = K2 * 255 < 65536;
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
    const bool fast = a.aligned4 && c0 >= 0 && c0 + 3 < a.W;
    const bool live = c0 < a.W + R4;  // columns far right of the image are never needed
    const int xc0 = min(max(c0, 0), a.W - 1), xc1 = min(max(c0 + 1, 0), a.W - 1), xc2 = min(max(c0 + 2, 0), a.W - 1),
              xc3 = min(max(c0 + 3, 0), a.W - 1);
    const int yEnd = min(y0 + THR_RH, a.H);
    const int nrows = (yEnd - y0) + 2 * R;
    const bool store_vec = (c0 + 3 < a.W) && ((a.W & 3) == 0);
    // validity of the 4 columns as a nibble
    const uint32_t vmask = c0 >= a.W ? 0u : (c0 + 3 < a.W ? 15u : ((1u << (a.W - c0)) - 1u));
    const size_t fo = (size_t)f * a.out_mul + a.out_off;
    uint8_t* orow = a.thres + fo * a.W * a.H + (size_t)y0 * a.W + c0;
    uint32_t* brow = a.bits + fo * a.bits_words + bit_word_index(a.wpr, BIT_PAD + (X0 >> 5) + (t >> 3), y0);
    int btr = (y0 + 1) & 31;  // row inside the bit tile
    const int bjump = a.wpr * BIT_TILE - (BIT_TILE - 1);
    const int cst = K2 * a.idelta - (K2 - 1) / 2;  // S >= K2*src + cst  <=>  src - mean <= -idelta
    const uint32_t M = 0x00FF00FFu;
    uint32_t ring[K];
#pragma unroll
    for (int j = 0; j < K; j++) ring[j] = 0u;
    uint32_t E = 0u, O = 0u;  // vertical sums of columns (0,2) and (1,3), 16 bits each
    int buf = 0;
    // software pipeline: the pixels of row i+1 are requested before row i is processed (ncu r1e: 43 % of the
    // stall samples of the non-pipelined loop sat on the first use of the freshly loaded word)
    auto load_row = [&](int i) -> uint32_t {
        if (!live || i >= nrows) return 0u;
        int yy = min(max(y0 - R + i, 0), a.H - 1);
        const uint8_t* rowp = src + (size_t)yy * a.grey_row;
        if (fast) return __ldg(reinterpret_cast<const uint32_t*>(rowp + c0));
        return (uint32_t)rowp[xc0] | ((uint32_t)rowp[xc1] << 8) | ((uint32_t)rowp[xc2] << 16) | ((uint32_t)rowp[xc3] << 24);
    };
    uint32_t p_next = load_row(0), p_next2 = load_row(1);
    for (int base = 0; base < nrows; base += K) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int i = base + j;
            if (i < nrows) {
                const uint32_t p = p_next;
                p_next = p_next2;
                p_next2 = load_row(i + 2);
                const uint32_t old = ring[j];
                ring[j] = p;
                E = E + (p & M) - (old & M);
                O = O + ((p >> 8) & M) - ((old >> 8) & M);
                if (i >= 2 * R) {
                    *reinterpret_cast<uint2*>(&cs[buf][ci]) = make_uint2(__byte_perm(E, O, 0x5410), __byte_perm(E, O, 0x7632));
                    __syncthreads();
                    if (is_out) {
                        uint32_t w[NW];
                        const uint2* wp = reinterpret_cast<const uint2*>(&cs[buf][ci - R4]);
#pragma unroll
                        for (int q = 0; q < NW / 2; q++) {
                            uint2 u = wp[q];
                            w[2 * q] = u.x;
                            w[2 * q + 1] = u.y;
                        }
                        int S0, S1, S2, S3;
                        if constexpr (PACKED) {
                            // (S0,S1) = sum of pairs starting at R4-R .. R4+R; (S2,S3) = the same window two columns on
                            uint32_t s01 = thr_pair_sum<R4 - R, R4 + R>(w);
                            uint32_t s23 = s01 - thr_pair<R4 - R>(w) - thr_pair<R4 - R + 1>(w) + thr_pair<R4 + R + 1>(w) + thr_pair<R4 + R + 2>(w);
                            S0 = (int)(s01 & 0xFFFFu);
                            S1 = (int)(s01 >> 16);
                            S2 = (int)(s23 & 0xFFFFu);
                            S3 = (int)(s23 >> 16);
                        } else {
                            int v[NV];
#pragma unroll
                            for (int q = 0; q < NW; q++) {
                                v[2 * q] = (int)(w[q] & 0xFFFFu);
                                v[2 * q + 1] = (int)(w[q] >> 16);
                            }
                            S0 = 0;
#pragma unroll
                            for (int d = R4 - R; d <= R4 + R; d++) S0 += v[d];
                            S1 = S0 + v[R4 + 1 + R] - v[R4 - R];
                            S2 = S1 + v[R4 + 2 + R] - v[R4 + 1 - R];
                            S3 = S2 + v[R4 + 3 + R] - v[R4 + 2 - R];
                        }
                        const uint32_t c = ring[(j + K - R) % K];  // centre row
                        uint32_t nibble = (uint32_t)(S0 >= (int)(c & 255u) * K2 + cst) | ((uint32_t)(S1 >= (int)((c >> 8) & 255u) * K2 + cst) << 1) |
                                          ((uint32_t)(S2 >= (int)((c >> 16) & 255u) * K2 + cst) << 2) | ((uint32_t)(S3 >= (int)(c >> 24) * K2 + cst) << 3);
                        nibble &= vmask;
                        const uint32_t outb = ((nibble * 0x00204081u) & 0x01010101u) * 255u;  // bit j -> byte j = 0xFF
                        if (store_vec) {
                            *reinterpret_cast<uint32_t*>(orow) = outb;
                        } else {
                            for (int jj = 0; jj < 4; jj++)
                                if ((vmask >> jj) & 1u) orow[jj] = (uint8_t)(outb >> (8 * jj));
                        }
                        uint32_t word = nibble << (4 * (t & 7));
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
                        word |= __shfl_xor_sync(0xFFFFFFFFu, word, 4);
                        if ((t & 7) == 0 && vmask) *brow = word;
                        orow += a.W;
                        brow += btr == 31 ? bjump : 1;
                        btr = (btr + 1) & 31;
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
