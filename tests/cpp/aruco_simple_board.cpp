// Headless equivalent of the reference's utils/aruco_simple_board.cpp on the C++ facade, driven entirely by the
// reference's YAML files:
//   aruco_simple_board <frame.raw> <width> <height> <boardConfig.yml> <intrinsics.yml> <markerSizeMeters> <out.yml>
// detects the markers, runs BoardDetector::detect and saves the Board in the format of the reference's golden files
// (test/core_tests.cpp:164-195).  Used by tests/test_gpu_api.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/aruco/serialization.hpp"

int main(int argc, char** argv) {
    if (argc < 8) { std::fprintf(stderr, "usage: aruco_simple_board frame.raw W H board.yml intrinsics.yml size out.yml\n"); return 2; }
    int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    std::vector<uint8_t> img((size_t)W * H);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "cannot read frame\n"); return 2; }
    std::fclose(f);
    try {
        aruco::BoardConfiguration TheBoardConfig;
        TheBoardConfig.readFromFile(argv[4]);
        aruco::CameraParameters CamParam = aruco::readCameraParameters(argv[5]);
        CamParam.resize(aruco::Size(W, H));
        const float MarkerSize = (float)std::atof(argv[6]);
        aruco::MarkerDetector MDetector;
        std::vector<aruco::Marker> Markers;
        MDetector.detect(aruco::ImageView(img.data(), H, W), Markers);  // utils/aruco_simple_board.cpp:78
        aruco::BoardDetector BD(MDetector);
        aruco::Board TheBoardDetected;
        float prob = BD.detect(Markers, TheBoardConfig, TheBoardDetected, CamParam, MarkerSize);  // :82
        std::printf("%zu markers, board probability %.6f, %zu board markers\n", Markers.size(), prob, TheBoardDetected.size());
        aruco::saveBoard(TheBoardDetected, argv[7]);
    } catch (const aruco::Exception& e) {
        std::fprintf(stderr, "aruco::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
