// Marker / board rendering on the device:
//   FiducidalMarkers::createMarkerImage, createBoardImage, createBoardImage_ChessBoard, createBoardImage_Frame
//   (src/arucofidmarkers.cpp:213-407) and MarkerCode::getImg (src/highlyreliablemarkers.cpp:234-256).
// A canvas is filled with its background, then one kernel draws every marker rectangle: a pixel of a marker of side
// `size` lies in cell (px / (size/7), py / (size/7)); cells 1..5 x 1..5 carry the 5x5 code, everything else (border
// cells and the size % 7 remainder strip) stays black -- exactly what the reference's Rect assignments produce.
#pragma once
#include <stdint.h>

namespace ab {

struct RenderRect {
    int x0, y0, size;  // top-left corner and side of the marker on the canvas
    int id;            // Fiducidal id (0..1023)
};

__global__ void k_fill_u8(uint8_t* img, size_t n, uint8_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) img[i] = v;
}

// rectangles [0, n_black) are plain black squares (locked-marker corner blocks), the rest are Fiducidal markers
__global__ void k_render_fiducidal(uint8_t* img, int W, int H, const RenderRect* rects, int n_black) {
    const RenderRect r = rects[blockIdx.y];
    const int sw = r.size / 7;
    // words of the Hamming-style row code, bit 4 = column 0 (arucofidmarkers.cpp:221)
    const int words[4] = {0x10, 0x17, 0x09, 0x0e};
    const int total = r.size * r.size;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int py = i / r.size, px = i - py * r.size;
        uint8_t v = 0;
        if ((int)blockIdx.y >= n_black && sw > 0) {
            const int cx = px / sw, cy = py / sw;
            if (cx >= 1 && cx <= 5 && cy >= 1 && cy <= 5) {
                const int val = words[(r.id >> (2 * (4 - (cy - 1)))) & 3];
                v = ((val >> (4 - (cx - 1))) & 1) ? 255 : 0;
            }
        }
        const int X = r.x0 + px, Y = r.y0 + py;
        if (X >= 0 && Y >= 0 && X < W && Y < H) img[(size_t)Y * W + X] = v;
    }
}

// MarkerCode::getImg: black canvas of side pix (a multiple of n+2), cell (i+1, j+1) white where bit i*n+j is set
__global__ void k_render_hrm(uint8_t* img, int pix, int n, const uint8_t* bits) {
    const int cell = pix / (n + 2);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pix * pix; i += gridDim.x * blockDim.x) {
        const int py = i / pix, px = i - py * pix;
        const int cy = py / cell - 1, cx = px / cell - 1;
        img[i] = (cy >= 0 && cy < n && cx >= 0 && cx < n && bits[cy * n + cx]) ? 255 : 0;
    }
}

// HighlyReliableMarkers::createBoardImage (src/highlyreliablemarkers.cpp:498-545): rect.id indexes the dictionary;
// every marker is MarkerCode::getImg at side rect.size (a multiple of n+2 by construction)
__global__ void k_render_hrm_board(uint8_t* img, int W, int H, const RenderRect* rects, int n, const uint8_t* bits) {
    const RenderRect r = rects[blockIdx.y];
    const int cell = r.size / (n + 2);
    const uint8_t* code = bits + (size_t)r.id * n * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < r.size * r.size; i += gridDim.x * blockDim.x) {
        const int py = i / r.size, px = i - py * r.size;
        const int cy = py / cell - 1, cx = px / cell - 1;
        const int X = r.x0 + px, Y = r.y0 + py;
        if (X < W && Y < H) img[(size_t)Y * W + X] = (cy >= 0 && cy < n && cx >= 0 && cx < n && code[cy * n + cx]) ? 255 : 0;
    }
}

}  // namespace ab
