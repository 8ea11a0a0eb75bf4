// Headless equivalent of the reference's utils/aruco_create_board.cpp on the C++ facade:
//   aruco_create_board <XSize> <YSize> <pixSize> <type 0|1|2> <interMarkerDistance> <out.raw> <out.yml> id0 id1 ...
// writes the board image (raw 8-bit, size printed on stdout) and its BoardConfiguration YAML, then detects the board in
// its own image and prints the detection probability.  Used by tests/test_gpu_api.py.
#include <cstdio>
#include <cstdlib>
#include "../../include/aruco/arucofidmarkers.hpp"

int main(int argc, char** argv) {
    if (argc < 9) { std::fprintf(stderr, "usage: aruco_create_board X Y pix type dist out.raw out.yml ids...\n"); return 2; }
    const int X = std::atoi(argv[1]), Y = std::atoi(argv[2]), pix = std::atoi(argv[3]), type = std::atoi(argv[4]), dist = std::atoi(argv[5]);
    std::vector<int> ids;
    for (int i = 8; i < argc; i++) ids.push_back(std::atoi(argv[i]));
    try {
        aruco::MarkerDetector md;
        aruco::BoardConfiguration BInfo;
        aruco::Image8 img = type == 0 ? aruco::FiducidalMarkers::createBoardImage(md, aruco::Size(X, Y), pix, dist, BInfo, ids)
                          : type == 1 ? aruco::FiducidalMarkers::createBoardImage_ChessBoard(md, aruco::Size(X, Y), pix, BInfo, ids)
                                      : aruco::FiducidalMarkers::createBoardImage_Frame(md, aruco::Size(X, Y), pix, dist, BInfo, ids);
        FILE* f = std::fopen(argv[6], "wb");
        if (!f || std::fwrite(img.data.data(), 1, img.data.size(), f) != img.data.size()) return 2;
        std::fclose(f);
        BInfo.saveToFile(argv[7]);
        // a white margin, then detect the board in its own picture
        const int m = pix / 2, W = img.cols + 2 * m, H = img.rows + 2 * m;
        std::vector<uint8_t> page((size_t)W * H, 255);
        for (int y = 0; y < img.rows; y++) std::memcpy(&page[(size_t)(y + m) * W + m], &img.data[(size_t)y * img.cols], (size_t)img.cols);
        std::vector<aruco::Marker> markers;
        md.detect(aruco::ImageView(page.data(), H, W), markers);
        float K[9] = {(float)W, 0, W / 2.f, 0, (float)W, H / 2.f, 0, 0, 1}, D[5] = {0, 0, 0, 0, 0};
        aruco::CameraParameters cp(K, D, aruco::Size(W, H));
        aruco::BoardDetector bd(md);
        aruco::Board board;
        float prob = bd.detect(markers, BInfo, board, cp, 0.05f);
        std::printf("%d %d %zu %zu %.6f\n", img.cols, img.rows, BInfo.size(), markers.size(), prob);
    } catch (const aruco::Exception& e) {
        std::fprintf(stderr, "aruco::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
