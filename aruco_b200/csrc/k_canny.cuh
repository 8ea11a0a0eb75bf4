// ThresholdMethods::CANNY (src/markerdetector.cpp:664-675): cv::Canny(grey, out, 10, 220), aperture 3, L1 gradient.
// Restated from OpenCV's algorithm (bit-exact against cv2 4.13 on the reference frames and on noise):
//   Sobel 3x3 with replicated border -> |dx|+|dy| (magnitude outside the image = 0) -> non-maximum suppression
//   with the fixed-point tangent tests (TG22 = 13573 / 2^15; comparisons "> left && >= right", "> up && >= down",
//   strict on the diagonals) -> candidates (m > low) and seeds (m > high) -> hysteresis: every candidate
//   8-connected to a seed through candidates is an edge.  The edge set is unique, so the propagation order is free:
//   k_canny_hyst relaxes 32x32 tiles to a fixed point in shared memory and is re-launched until no tile changes.
#pragma once
#include "ab_device.cuh"

namespace ab {

// map values: 0 = not an edge candidate, 1 = candidate (weak), 2 = edge
__global__ void __launch_bounds__(256) k_canny_nms(const uint8_t* grey, size_t grey_row, size_t grey_frame, uint8_t* map, int W, int H,
                                                   int low, int high) {
    constexpr int TX = 32, TY = 8;
    __shared__ uint8_t s_pix[TY + 4][TX + 4];
    __shared__ int s_mag[TY + 2][TX + 2];
    const int f = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.y * TX + threadIdx.x;
    const uint8_t* src = grey + (size_t)f * grey_frame;
    for (int i = tid; i < (TY + 4) * (TX + 4); i += TX * TY) {
        int py = i / (TX + 4), px = i - py * (TX + 4);
        int gy = min(max(y0 - 2 + py, 0), H - 1), gx = min(max(x0 - 2 + px, 0), W - 1);
        s_pix[py][px] = src[(size_t)gy * grey_row + gx];
    }
    __syncthreads();
    for (int i = tid; i < (TY + 2) * (TX + 2); i += TX * TY) {
        int my = i / (TX + 2), mx = i - my * (TX + 2);
        int gy = y0 - 1 + my, gx = x0 - 1 + mx;
        int m = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const int py = my + 1, px = mx + 1;  // centre in s_pix
            int dx = (s_pix[py - 1][px + 1] - s_pix[py - 1][px - 1]) + 2 * (s_pix[py][px + 1] - s_pix[py][px - 1]) + (s_pix[py + 1][px + 1] - s_pix[py + 1][px - 1]);
            int dy = (s_pix[py + 1][px - 1] - s_pix[py - 1][px - 1]) + 2 * (s_pix[py + 1][px] - s_pix[py - 1][px]) + (s_pix[py + 1][px + 1] - s_pix[py - 1][px + 1]);
            m = abs(dx) + abs(dy);
        }
        s_mag[my][mx] = m;
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= W || gy >= H) return;
    const int py = threadIdx.y + 2, px = threadIdx.x + 2, my = threadIdx.y + 1, mx = threadIdx.x + 1;
    const int xs = (s_pix[py - 1][px + 1] - s_pix[py - 1][px - 1]) + 2 * (s_pix[py][px + 1] - s_pix[py][px - 1]) + (s_pix[py + 1][px + 1] - s_pix[py + 1][px - 1]);
    const int ys = (s_pix[py + 1][px - 1] - s_pix[py - 1][px - 1]) + 2 * (s_pix[py + 1][px] - s_pix[py - 1][px]) + (s_pix[py + 1][px + 1] - s_pix[py - 1][px + 1]);
    const int m = s_mag[my][mx];
    uint8_t v = 0;
    if (m > low) {
        const long long x = abs(xs), y = (long long)abs(ys) << 15;
        const long long tg22x = x * 13573;
        bool keep;
        if (y < tg22x) {
            keep = m > s_mag[my][mx - 1] && m >= s_mag[my][mx + 1];
        } else {
            const long long tg67x = tg22x + (x << 16);
            if (y > tg67x) {
                keep = m > s_mag[my - 1][mx] && m >= s_mag[my + 1][mx];
            } else {
                const int s = (xs ^ ys) < 0 ? -1 : 1;
                keep = m > s_mag[my - 1][mx - s] && m > s_mag[my + 1][mx + s];
            }
        }
        if (keep) v = m > high ? 2 : 1;
    }
    map[((size_t)f * H + gy) * W + gx] = v;
}

// one relaxation pass over 32x32 tiles (iterated to a fixed point inside the tile); *changed counts tiles that changed
__global__ void __launch_bounds__(256) k_canny_hyst(uint8_t* map, int W, int H, unsigned int* changed) {
    constexpr int T = 32;
    __shared__ uint8_t s[T + 2][T + 2];
    __shared__ int s_changed, s_any;
    const int f = blockIdx.z, x0 = blockIdx.x * T, y0 = blockIdx.y * T;
    const int tid = threadIdx.x;
    uint8_t* m = map + (size_t)f * W * H;
    for (int i = tid; i < (T + 2) * (T + 2); i += 256) {
        int py = i / (T + 2), px = i - py * (T + 2);
        int gy = y0 - 1 + py, gx = x0 - 1 + px;
        s[py][px] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? m[(size_t)gy * W + gx] : 0;
    }
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (;;) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < T * T; i += 256) {
            int py = i / T + 1, px = i % T + 1;
            if (s[py][px] == 1) {
                bool near = s[py - 1][px - 1] == 2 || s[py - 1][px] == 2 || s[py - 1][px + 1] == 2 || s[py][px - 1] == 2 || s[py][px + 1] == 2 ||
                            s[py + 1][px - 1] == 2 || s[py + 1][px] == 2 || s[py + 1][px + 1] == 2;
                if (near) {
                    s[py][px] = 2;  // benign race: values only ever go 1 -> 2
                    s_changed = 1;
                }
            }
        }
        __syncthreads();
        if (!s_changed) break;
        if (tid == 0) s_any = 1;
        __syncthreads();
    }
    if (s_any) {
        for (int i = tid; i < T * T; i += 256) {
            int py = i / T + 1, px = i % T + 1;
            int gy = y0 + py - 1, gx = x0 + px - 1;
            if (gy < H && gx < W && s[py][px] == 2) m[(size_t)gy * W + gx] = 2;
        }
        if (tid == 0) atomicAdd(changed, 1u);
    }
}

// map (0/1/2) -> binary image {0,255} in place + packed bits
__global__ void k_canny_finish(uint8_t* thres, uint32_t* bits, size_t bits_words, int W, int H, int wpr, int B, int out_mul, int out_off) {
    int ww = (W + 31) >> 5;
    size_t total = (size_t)ww * H * B;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int w = (int)(i % ww);
        int y = (int)((i / ww) % H);
        int f = (int)(i / ((size_t)ww * H));
        const size_t fo = (size_t)f * out_mul + out_off;
        uint8_t* row = thres + (fo * H + y) * W;
        uint32_t word = 0;
        for (int j = 0; j < 32; j++) {
            int x = 32 * w + j;
            if (x >= W) break;
            bool on = row[x] == 2;
            row[x] = on ? 255 : 0;
            if (on) word |= 1u << j;
        }
        bits[fo * bits_words + bit_word_index(wpr, BIT_PAD + w, y)] = word;
    }
}

}  // namespace ab
