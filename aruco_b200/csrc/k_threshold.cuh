// Stage 1: MarkerDetector::thresHold (src/markerdetector.cpp:643-677) for a batch of grey frames.
//   ADPT_THRES: cv::adaptiveThreshold(MEAN_C, BINARY_INV, k, C)  -- k x k box mean with replicate border,
//               mean = (2S + k^2) / (2 k^2) (integer), dst = (src - mean <= -floor(C)) ? 255 : 0   
//   FIXED_THRES: cv::threshold(BINARY_INV, p1): dst = src > floor(p1) ? 0 : 255
// Output: the u8 {0,255} image (getThresholdedImage is API) and a 1-bit-per-pixel packed copy with zero
// padding that the contour stage reads (8x less traffic than re-reading the u8 image).
//
// Kernel shape: one CTA owns a strip of TWo output columns x RH output rows.  Each thread owns 4 adjacent
// columns (one 32-bit load per source row), keeps their vertical k-sums in registers (sliding down the
// strip: + new row - row k back, the k most recent rows live in a shared-memory ring), publishes them to
// shared memory and forms the horizontal k-sums from its neighbours' columns.  The division of the mean
// is folded into the comparison:  mean >= T  <=>  2S + k^2 >= 2 k^2 T.
#pragma once
#include "ab_device.cuh"

namespace ab {

struct ThrArgs {
    const uint8_t* grey;
    size_t grey_row, grey_frame;
    uint8_t* thres;
    uint32_t* bits;
    size_t bits_words;
    int W, H, wpr;
    int k, idelta;
    int TWo, RH, R4;
    int aligned4;  // source rows allow 32-bit loads
    int out_mul, out_off;  // output (virtual) frame = f * out_mul + out_off  (setThresholdParamRange: several images per frame)
    int skip_u8;           // erosion follows and writes the u8 image itself: kernels that honour this write only the packed image
};

__global__ void k_threshold_adaptive(ThrArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int nth = blockDim.x, t = threadIdx.x, SPAN = 4 * nth;
    const int k = a.k, r = k >> 1, k2 = k * k;
    uint8_t* ring = smem;                                            // k rows of SPAN source pixels
    uint32_t* cs = (uint32_t*)(smem + (((size_t)k * SPAN + 15) & ~(size_t)15));  // SPAN column sums
    uint8_t* nib = (uint8_t*)(cs + SPAN);                            // nth result nibbles
    const int X0 = blockIdx.x * a.TWo, y0 = blockIdx.y * a.RH, f = blockIdx.z;
    const int c0 = X0 - a.R4 + 4 * t;
    const uint8_t* src = a.grey + (size_t)f * a.grey_frame;
    const size_t fo = (size_t)f * a.out_mul + a.out_off;
    uint8_t* dst = a.thres + fo * a.W * a.H;
    uint32_t* bits = a.bits + fo * a.bits_words;
    const bool fast = a.aligned4 && c0 >= 0 && c0 + 3 < a.W;
    int xc[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xc[j] = min(max(c0 + j, 0), a.W - 1);
    const int yEnd = min(y0 + a.RH, a.H);
    const int nrows = (yEnd - y0) + 2 * r;
    const bool is_out = c0 >= X0 && c0 < X0 + a.TWo && c0 < a.W;
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int slot = 0, cslot = 0;  // ring slot of the incoming row / of the centre row
    for (int i = 0; i < nrows; i++) {
        int yy = min(max(y0 - r + i, 0), a.H - 1);
        const uint8_t* rowp = src + (size_t)yy * a.grey_row;
        uint32_t p;
        if (fast)
            p = *reinterpret_cast<const uint32_t*>(rowp + c0);
        else
            p = (uint32_t)rowp[xc[0]] | ((uint32_t)rowp[xc[1]] << 8) | ((uint32_t)rowp[xc[2]] << 16) | ((uint32_t)rowp[xc[3]] << 24);
        uint32_t* rp = reinterpret_cast<uint32_t*>(ring + (size_t)slot * SPAN + 4 * t);
        uint32_t old = (i >= k) ? *rp : 0u;
        *rp = p;
        s0 += (int)(p & 255u) - (int)(old & 255u);
        s1 += (int)((p >> 8) & 255u) - (int)((old >> 8) & 255u);
        s2 += (int)((p >> 16) & 255u) - (int)((old >> 16) & 255u);
        s3 += (int)(p >> 24) - (int)(old >> 24);
        if (++slot == k) slot = 0;
        if (i >= 2 * r) {
            const int yo = y0 + i - 2 * r;
            *reinterpret_cast<uint4*>(cs + 4 * t) = make_uint4((uint32_t)s0, (uint32_t)s1, (uint32_t)s2, (uint32_t)s3);
            __syncthreads();
            uint32_t nibble = 0;
            if (is_out) {
                uint32_t c = *reinterpret_cast<const uint32_t*>(ring + (size_t)cslot * SPAN + 4 * t);
                const uint32_t* w = cs + 4 * t;
                int S = 0;
                for (int d = -r; d <= r; d++) S += (int)w[d];
                uint32_t outb = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int T = (int)((c >> (8 * j)) & 255u) + a.idelta;
                    bool on = (2 * S + k2 >= 2 * k2 * T) && (c0 + j < a.W);
                    if (on) {
                        nibble |= 1u << j;
                        outb |= 255u << (8 * j);
                    }
                    if (j < 3) S += (int)w[j + 1 + r] - (int)w[j - r];
                }
                uint8_t* orow = dst + (size_t)yo * a.W + c0;
                if (c0 + 3 < a.W && (a.W & 3) == 0) {
                    *reinterpret_cast<uint32_t*>(orow) = outb;
                } else {
                    for (int j = 0; j < 4; j++)
                        if (c0 + j < a.W) orow[j] = (uint8_t)(outb >> (8 * j));
                }
            }
            nib[t] = (uint8_t)nibble;
            __syncthreads();
            if (t < (a.TWo >> 5) && X0 + 32 * t < a.W) {
                const uint8_t* nb = nib + (a.R4 >> 2) + 8 * t;
                uint32_t word = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) word |= (uint32_t)nb[q] << (4 * q);
                bits[bit_word_index(a.wpr, BIT_PAD + (X0 >> 5) + t, yo)] = word;
            }
            if (++cslot == k) cslot = 0;
        } else if (i >= r) {
            if (++cslot == k) cslot = 0;
        }
    }
}

// FIXED_THRES and bit packing of an existing binary image: one thread per 32-pixel word.
// mode 0: dst = src > thr ? 0 : 255 (threshold BINARY_INV);  mode 1: dst = src (non-zero = fg), pack only
__global__ void k_threshold_fixed(const uint8_t* grey, size_t grey_row, size_t grey_frame, uint8_t* thres,
                                  uint32_t* bits, size_t bits_words, int W, int H, int wpr, int thr, int mode, int B,
                                  int out_mul, int out_off) {
    int ww = (W + 31) >> 5;
    size_t total = (size_t)ww * H * B;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int w = (int)(i % ww);
        int y = (int)((i / ww) % H);
        int f = (int)(i / ((size_t)ww * H));
        const uint8_t* rowp = grey + (size_t)f * grey_frame + (size_t)y * grey_row;
        const size_t fo = (size_t)f * out_mul + out_off;
        uint8_t* orow = thres + (fo * H + y) * W;
        uint32_t word = 0;
        for (int j = 0; j < 32; j++) {
            int x = 32 * w + j;
            if (x >= W) break;
            uint8_t v = rowp[x];
            bool on = mode == 0 ? !((int)v > thr) : (v != 0);
            if (on) word |= 1u << j;
            if (mode == 0) orow[x] = on ? 255 : 0;
            else if (orow + x != rowp + x) orow[x] = v;
        }
        bits[fo * bits_words + bit_word_index(wpr, BIT_PAD + w, y)] = word;
    }
}

// 3x3 erosion (cv::erode(thres, Mat()), out-of-image = 255) on the packed image: 32 pixels per AND.  Writes the eroded
// packed image and the eroded u8 image (the threshold kernel skips its own u8 store when erosion is on, so the binarised
// frame is written to HBM once).  One thread per 32-pixel word: nine word loads (three rows, L1/L2 hits), two shifts and
// four ANDs per row, then the 32 result bytes go out as two 16-byte stores.
__device__ __forceinline__ uint32_t bits4_to_bytes(uint32_t n) {  // bit j of the nibble -> byte j = 0x00 / 0xFF
    return (((n & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu;
}
constexpr int ERODE_CH = 128;             // word columns per shared-memory chunk
constexpr int ERODE_PITCH = ERODE_CH + 3;  // odd multiple: rows of a tile land in different banks
__global__ void __launch_bounds__(256) k_erode(const uint32_t* in, uint32_t* out, uint8_t* thres, size_t bits_words, int W, int H, int wpr, int B) {
    // The packed image is stored in tiles of 32 rows x one word column (one 128-byte line, ab_trace.cuh), so a row-wise
    // reader touches one line per word (r2m: 0.62 ms per 64 4K frames).  A CTA therefore takes a strip of one tile row:
    // (1) its tiles go to shared memory line by line (lane = row of the tile), with a one-word / one-row apron in which
    // everything outside the image counts as 255; (2) the 3x3 AND per word, lane = word column; (3) the eroded words go back
    // tile by tile and the u8 rows go out as consecutive 16-byte pieces.
    __shared__ uint32_t s_in[34][ERODE_PITCH];
    __shared__ uint32_t s_out[32][ERODE_PITCH];
    const int ww = (W + 31) >> 5, ntr = (H + 2 + BIT_TILE - 1) / BIT_TILE;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = blockDim.x >> 5;
    const bool vec = (W & 15) == 0 && (((uintptr_t)thres) & 15) == 0;
    for (int strip = blockIdx.x; strip < ntr * B; strip += gridDim.x) {
        const int f = strip / ntr, tr = strip - f * ntr;
        const uint32_t* base = in + (size_t)f * bits_words;
        uint32_t* obase = out + (size_t)f * bits_words;
        for (int c0 = 0; c0 < ww; c0 += ERODE_CH) {
            const int nc = min(ERODE_CH, ww - c0);
            // (1) columns c0-1 .. c0+nc of padded rows 32 tr - 1 .. 32 tr + 32
            for (int cc = warp; cc < nc + 2; cc += nwarp) {
                const int c = c0 - 1 + cc;
                const bool col_ok = c >= 0 && c < ww;
                const uint32_t colmask = col_ok ? (c == ww - 1 && (W & 31) ? ((1u << (W & 31)) - 1u) : 0xFFFFFFFFu) : 0u;
                {   // the tile itself: lane = row in tile, image row y = 32 tr + lane - 1
                    const int y = 32 * tr + lane - 1;
                    uint32_t v = 0xFFFFFFFFu;
                    if (col_ok && y >= 0 && y < H) v = base[((size_t)tr * wpr + BIT_PAD + c) * BIT_TILE + lane] | ~colmask;
                    s_in[lane + 1][cc] = v;
                }
                if (lane < 2) {  // the rows above and below the tile
                    const int y = lane == 0 ? 32 * tr - 2 : 32 * tr + 31;
                    uint32_t v = 0xFFFFFFFFu;
                    if (col_ok && y >= 0 && y < H) v = base[bit_word_index(wpr, BIT_PAD + c, y)] | ~colmask;
                    s_in[lane == 0 ? 0 : 33][cc] = v;
                }
            }
            __syncthreads();
            // (2) eroded word of (row r, column c0 + cc): lane = column
            for (int i = t; i < 32 * nc; i += blockDim.x) {
                const int r = i / nc, cc = i - r * nc;
                uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    const uint32_t prv = s_in[r + dy][cc], cur = s_in[r + dy][cc + 1], nxt = s_in[r + dy][cc + 2];
                    acc &= cur & ((cur << 1) | (prv >> 31)) & ((cur >> 1) | (nxt << 31));
                }
                if (c0 + cc == ww - 1 && (W & 31)) acc &= (1u << (W & 31)) - 1u;
                s_out[r][cc] = acc;
            }
            __syncthreads();
            // (3a) packed image, tile by tile (rows outside the image stay zero: the frame of the padded buffer)
            for (int cc = warp; cc < nc; cc += nwarp) {
                const int y = 32 * tr + lane - 1;
                if (y >= 0 && y < H) obase[((size_t)tr * wpr + BIT_PAD + c0 + cc) * BIT_TILE + lane] = s_out[lane][cc];
            }
            // (3b) u8 image: 16 bytes per thread, consecutive threads consecutive pieces of a row
            for (int i = t; i < 32 * 2 * nc; i += blockDim.x) {
                const int r = i / (2 * nc), hw = i - r * (2 * nc), cc = hw >> 1, half = hw & 1;
                const int y = 32 * tr + r - 1;
                if (y < 0 || y >= H) continue;
                const int x0 = 32 * (c0 + cc) + 16 * half, nv = min(16, W - x0);
                if (nv <= 0) continue;
                const uint32_t h16 = s_out[r][cc] >> (16 * half);
                uint8_t* orow = thres + ((size_t)f * H + y) * W + x0;
                if (vec && nv == 16) {
                    *reinterpret_cast<uint4*>(orow) = make_uint4(bits4_to_bytes(h16), bits4_to_bytes(h16 >> 4), bits4_to_bytes(h16 >> 8), bits4_to_bytes(h16 >> 12));
                } else {
                    for (int j = 0; j < nv; j++) orow[j] = (h16 >> j) & 1u ? 255 : 0;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace ab
