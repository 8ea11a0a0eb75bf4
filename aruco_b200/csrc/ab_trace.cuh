// Border following as a cycle walk over (pixel, back-direction) states.
//
// Replaces cv::findContours(RETR_LIST, CHAIN_APPROX_NONE) at src/markerdetector.cpp:510-511 of the
// reference.  OpenCV's Suzuki-Abe tracer is a serial raster scan that marks pixels so that every border
// is started once; here every border is a cycle of a local successor function (the 3x3 neighbourhood
// decides the next state), and the Suzuki start of a cycle is the trigger with the smallest raster scan
// position that lies on it.  Only pixels that *can* be such a minimum are start candidates:
//   outer border: foreground pixel whose W, NW, N, NE neighbours are background
//                 (the raster-first pixel of an 8-connected component always is one)
//   hole border:  background pixel h whose W and N neighbours are foreground (the raster-first pixel of a
//                 4-connected hole always is one); the contour then starts at the pixel left of h.
// A candidate walks its cycle and gives up as soon as it meets the start state of a candidate with a
// smaller scan position; the survivor reproduces OpenCV's contour point for point.
//
// Directions (image y grows downwards):  0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE.
#pragma once
#include <stdio.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#else
#define AB_HD inline
#endif

namespace ab {

// Binary image packed 1 bit per pixel, LSB first, stored in TILES: one 32-pixel word column x 32 rows = 32
// consecutive words = 128 B = one L1 line.  With plain row-major rows a walker's 3x3 neighbourhood sat in three
// lines; 1024 resident walkers per SM then need 3072 lines of an L1 that holds ~1800, and the hit rate of the
// walker kernels was 17 % (ncu r1m) with L2 bandwidth as the bound.  Tiled, a neighbourhood is one line (two when
// it straddles a tile edge) and stays there for ~32 steps.
// One zero word column left and >= 1 right of every row, one zero row above and >= 1 below, so a 3x3 neighbourhood
// never needs a bounds test.
// -DAB_DEBUG_BOUNDS: in-kernel index asserts (compute-sanitizer is closed on the GPU pool).  A violated bound prints its
// source position and traps: the launch fails with an error instead of silently writing out of range.
#if defined(AB_DEBUG_BOUNDS) && defined(__CUDA_ARCH__)
#define AB_BOUND(cond)                                                             \
    do {                                                                           \
        if (!(cond)) {                                                             \
            printf("AB_DEBUG_BOUNDS violated: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            __trap();                                                              \
        }                                                                          \
    } while (0)
#else
#define AB_BOUND(cond) ((void)0)
#endif

constexpr int BIT_PAD = 1;     // zero word columns left of the image
constexpr int BIT_TILE = 32;   // rows (= words) per tile
AB_HD int bit_words_per_row(int W) { return ((W + 31) >> 5) + BIT_PAD + 1; }  // word columns incl. padding
AB_HD size_t bit_image_words(int W, int H) { return (size_t)bit_words_per_row(W) * BIT_TILE * (size_t)((H + 2 + BIT_TILE - 1) / BIT_TILE); }
// index of the word in padded word column wc (image pixels 32*(wc-BIT_PAD) ..+31) of image row y (-1 <= y <= H)
// (a frame's packed image has < 2^28 words for W, H <= 16384: 32-bit index arithmetic)
AB_HD uint32_t bit_word_index(int wpr, int wc, int y) {
    const int yp = y + 1;
    return (uint32_t)(((yp >> 5) * wpr + wc) * BIT_TILE + (yp & 31));
}

struct BitImage {
    const uint32_t* bits;  // points at the padded buffer
    int wpr;               // word columns per padded row (bit_words_per_row)
    int W, H;
    AB_HD const uint32_t* word(int wc, int y) const { return bits + bit_word_index(wpr, wc, y); }
};

// 3x3 window of pixel (x,y) as 9 bits: bits 0-2 = row above (x-1, x, x+1), 3-5 = own row, 6-8 = row below.
// The three rows are read with ONE word each; the word of the next column is fetched only when the 3-pixel window
// straddles a word boundary (2 of 32 positions).
AB_HD uint32_t window9(const BitImage& im, int x, int y) {
    const int p = x - 1 + 32 * BIT_PAD;  // pixel x lives at padded bit x + 32*BIT_PAD
    const int sh = p & 31, yr = (y + 1) & 31;
    AB_BOUND(x >= 0 && x < im.W && y >= 0 && y < im.H);  // a walker never leaves the image
    const uint32_t* r1 = im.word(p >> 5, y);
    const int jump = im.wpr * BIT_TILE - (BIT_TILE - 1);  // to the same column of the next tile row, minus 31
    const uint32_t* r0 = r1 - (yr == 0 ? jump : 1);
    const uint32_t* r2 = r1 + (yr == 31 ? jump : 1);
    uint32_t t = r0[0], m = r1[0], b = r2[0];
    uint32_t th = 0, mh = 0, bh = 0;
    if (sh > 29) {
        th = r0[BIT_TILE];
        mh = r1[BIT_TILE];
        bh = r2[BIT_TILE];
    }
#if defined(__CUDA_ARCH__)
    t = __funnelshift_r(t, th, sh) & 7u;
    m = __funnelshift_r(m, mh, sh) & 7u;
    b = __funnelshift_r(b, bh, sh) & 7u;
#else
    t = (uint32_t)((((uint64_t)th << 32) | t) >> sh) & 7u;
    m = (uint32_t)((((uint64_t)mh << 32) | m) >> sh) & 7u;
    b = (uint32_t)((((uint64_t)bh << 32) | b) >> sh) & 7u;
#endif
    return t + m * 8u + b * 64u;
}

// 8-neighbour mask from the window: bit d set <=> neighbour in direction d is foreground
AB_HD uint32_t nb_from_window(uint32_t w) {
    const uint32_t t = w & 7u, m = (w >> 3) & 7u, b = (w >> 6) & 7u;
    return ((m >> 2) & 1u) | (((t >> 2) & 1u) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) | ((m & 1u) << 4) |
           ((b & 1u) << 5) | (((b >> 1) & 1u) << 6) | (((b >> 2) & 1u) << 7);
}

AB_HD uint32_t neighbours8(const BitImage& im, int x, int y) { return nb_from_window(window9(im, x, y)); }

AB_HD int dir_dx(int d) { return (d == 0 || d == 1 || d == 7) ? 1 : ((d >= 3 && d <= 5) ? -1 : 0); }
AB_HD int dir_dy(int d) { return (d >= 1 && d <= 3) ? -1 : ((d >= 5) ? 1 : 0); }

// first foreground neighbour going CLOCKWISE from direction `from` (exclusive), -1 if none
AB_HD int first_clockwise(uint32_t nb, int from) {
    // branch-free: rotate so that direction `from` sits at bit 0; the clockwise-nearest neighbour is then the
    // HIGHEST set bit (bit 7 = from-1, ..., bit 0 = from itself after a full turn)
    uint32_t r = (((nb << 8) | nb) >> from) & 0xFFu;
    if (r == 0) return -1;
#if defined(__CUDA_ARCH__)
    int j = 31 - __clz((int)r);
#else
    int j = 31 - __builtin_clz(r);
#endif
    return (from + j) & 7;
}

// successor: first foreground neighbour going COUNTER-CLOCKWISE from back-direction b (exclusive).
// nb always has bit b set (we came from there), so this terminates.
AB_HD int next_dir(uint32_t nb, int b) {
    uint32_t r = ((nb >> (b + 1)) | (nb << (7 - b))) & 0xFFu;  // rotate right by b+1 within 8 bits
#if defined(__CUDA_ARCH__)
    int t = __ffs((int)r) - 1;
#else
    int t = __builtin_ffs((int)r) - 1;
#endif
    return (b + 1 + t) & 7;
}

AB_HD bool is_outer_candidate(uint32_t nb) { return (nb & 0x1Eu) == 0; }            // NE,N,NW,W all zero
AB_HD bool is_hole_candidate_east(uint32_t nb) { return (nb & 1u) == 0 && (nb & 2u); }  // E zero, NE set

enum TraceResult { TRACE_NOT_START = 0, TRACE_OK = 1, TRACE_TOO_LONG = 2, TRACE_ISOLATED = 3 };

struct TraceStart {
    int x, y;      // first contour point
    int b;         // back-direction of the start state
    int64_t key;   // raster scan position of the trigger
};

// Start state of a candidate. type 0: outer border starting at fg pixel (x,y); type 1: hole border whose
// trigger is the bg pixel (x,y) -- the contour starts at (x-1,y).  Returns false for an isolated pixel.
AB_HD bool make_start(const BitImage& im, int type, int x, int y, TraceStart& st) {
    st.key = (int64_t)y * im.W + x;
    if (type == 0) {
        st.x = x;
        st.y = y;
        st.b = first_clockwise(neighbours8(im, x, y), 4);
    } else {
        st.x = x - 1;
        st.y = y;
        st.b = first_clockwise(neighbours8(im, x - 1, y), 0);
    }
    return st.b >= 0;
}

struct WalkState {
    int x, y, b;  // at pixel (x,y), came from the neighbour in direction b
};

AB_HD bool same_state(const WalkState& a, const WalkState& c) { return a.x == c.x && a.y == c.y && a.b == c.b; }

// Is `s` the start state of a start candidate whose trigger scans before `key0`?  (nb = neighbours of s)
AB_HD bool is_smaller_trigger(const BitImage& im, const WalkState& s, uint32_t nb, int64_t key0) {
    int64_t k = (int64_t)s.y * im.W + s.x;
    if (is_outer_candidate(nb) && k < key0 && s.b == first_clockwise(nb, 4)) return true;
    if (is_hole_candidate_east(nb) && k + 1 < key0 && s.b == first_clockwise(nb, 0)) return true;
    return false;
}

// successor state (the border-following step of OpenCV's tracer)
AB_HD void walk_forward(WalkState& s, uint32_t nb) {
    int d = next_dir(nb, s.b);
    s.x += dir_dx(d);
    s.y += dir_dy(d);
    s.b = (d + 4) & 7;
}

// predecessor state: the successor function is a permutation of the states, its inverse probes clockwise
// returns the neighbour mask of the NEW position (callers reuse it for the trigger test)
AB_HD uint32_t walk_backward(const BitImage& im, WalkState& s) {
    int qx = s.x + dir_dx(s.b), qy = s.y + dir_dy(s.b);
    int d = (s.b + 4) & 7;  // direction from the predecessor pixel to the current one
    uint32_t nbq = neighbours8(im, qx, qy);
    s.b = first_clockwise(nbq, d);
    s.x = qx;
    s.y = qy;
    return nbq;
}

// ---- table-driven step (k_trace): all of the combinational logic of a step is a function of the 9-bit window and
// a 3-bit direction, so the kernel looks it up instead of computing it (the walkers are issue bound).
//   WALK_LUT_FW[w | b << 9]: state (pixel with window w, back-direction b) -> bits 0-2 direction of the successor,
//                            bit 3 / bit 4: the state is the start state of an outer / hole start candidate
//   WALK_LUT_BW[w | d << 9]: w = window of the predecessor pixel q, d = direction from q to the current pixel
//                            -> bits 0-2 back-direction b' of the predecessor state, bits 3/4 as above for (q, b')
constexpr int WALK_LUT_SIZE = 4096;
constexpr uint32_t WALK_TRIG_OUTER = 8u, WALK_TRIG_HOLE = 16u;
AB_HD uint8_t walk_lut_flags(uint32_t nb, int b) {
    uint32_t f = 0;
    if (is_outer_candidate(nb) && b == first_clockwise(nb, 4)) f |= WALK_TRIG_OUTER;
    if (is_hole_candidate_east(nb) && b == first_clockwise(nb, 0)) f |= WALK_TRIG_HOLE;
    return (uint8_t)f;
}
AB_HD uint8_t walk_lut_fw_entry(uint32_t idx) {
    const uint32_t nb = nb_from_window(idx & 511u);
    const int b = (int)(idx >> 9);
    return (uint8_t)((uint32_t)(next_dir(nb, b) & 7) | walk_lut_flags(nb, b));
}
AB_HD uint8_t walk_lut_bw_entry(uint32_t idx) {
    const uint32_t nb = nb_from_window(idx & 511u);
    int b = first_clockwise(nb, (int)(idx >> 9));
    if (b < 0) b = 0;  // isolated pixel: never reached by a walk
    return (uint8_t)((uint32_t)b | walk_lut_flags(nb, b));
}
// dx / dy of direction d, two bits per direction biased by 1
AB_HD int step_dx(int d) { return (int)((0x901Au >> (2 * d)) & 3u) - 1; }
AB_HD int step_dy(int d) { return (int)((0xA901u >> (2 * d)) & 3u) - 1; }

// Bidirectional search: is `st` the Suzuki start of its border?  Walks forwards and backwards alternately
// and stops as soon as either walker stands on the start state of a candidate with a smaller scan position
// (the expected walk is then ~ the distance to the next higher candidate instead of most of the border).
// When the walkers meet the whole cycle has been seen: *len = its length.
AB_HD int find_start_bidir(const BitImage& im, const TraceStart& st, int max_len, int* len) {
    WalkState fw{st.x, st.y, st.b}, bw = fw;
    int nf = 0, ng = 0;
    uint32_t nb_fw = neighbours8(im, fw.x, fw.y);
    for (;;) {
        walk_forward(fw, nb_fw);
        nf++;
        if (same_state(fw, bw)) break;
        nb_fw = neighbours8(im, fw.x, fw.y);
        if (is_smaller_trigger(im, fw, nb_fw, st.key)) return TRACE_NOT_START;
        uint32_t nb_bw = walk_backward(im, bw);
        ng++;
        if (same_state(fw, bw)) break;
        if (is_smaller_trigger(im, bw, nb_bw, st.key)) return TRACE_NOT_START;
        if (nf + ng >= max_len) return TRACE_TOO_LONG;
    }
    *len = nf + ng;
    return TRACE_OK;
}

// Walks the cycle of `st`.  Returns TRACE_OK with the length in *len when `st` is the Suzuki start of its
// border and the length is < max_len; TRACE_NOT_START when a smaller trigger lies on the cycle;
// TRACE_TOO_LONG when max_len points were passed without closing.  When `emit` is non-null the points are
// written as (x | y<<16).
AB_HD int trace_cycle(const BitImage& im, const TraceStart& st, int max_len, int* len, uint32_t* emit) {
    int x = st.x, y = st.y, b = st.b, n = 0;
    for (;;) {
        uint32_t nb = neighbours8(im, x, y);
        if (n > 0) {
            if (is_outer_candidate(nb) && (int64_t)y * im.W + x < st.key && b == first_clockwise(nb, 4))
                return TRACE_NOT_START;
            if (is_hole_candidate_east(nb) && (int64_t)y * im.W + x + 1 < st.key && b == first_clockwise(nb, 0))
                return TRACE_NOT_START;
        }
        if (emit) emit[n] = (uint32_t)x | ((uint32_t)y << 16);
        n++;
        int s = next_dir(nb, b);
        x += dir_dx(s);
        y += dir_dy(s);
        b = (s + 4) & 7;
        if (x == st.x && y == st.y && b == st.b) break;
        if (n >= max_len) return TRACE_TOO_LONG;
    }
    *len = n;
    return TRACE_OK;
}

}  // namespace ab
