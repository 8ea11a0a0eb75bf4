// Headless equivalent of the reference's utils/aruco_simple.cpp (config C1 driver) on the C++ facade:
//   aruco_simple <frame.raw> <width> <height> [fx fy cx cy k1 k2 p1 p2 k3 markerSize]
// prints one line per marker: id, 4 corners, Rvec, Tvec.  Used by tests/test_gpu_api.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/aruco/markerdetector.hpp"

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: aruco_simple frame.raw W H [fx fy cx cy k1 k2 p1 p2 k3 size]\n"); return 2; }
    int W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    std::vector<uint8_t> img((size_t)W * H);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "cannot read frame\n"); return 2; }
    std::fclose(f);
    try {
        aruco::MarkerDetector MDetector;
        std::vector<aruco::Marker> Markers;
        aruco::CameraParameters CamParam;
        float size = -1;
        if (argc >= 14) {
            float K[9] = {(float)std::atof(argv[4]), 0, (float)std::atof(argv[6]), 0, (float)std::atof(argv[5]), (float)std::atof(argv[7]), 0, 0, 1};
            float D[5];
            for (int i = 0; i < 5; i++) D[i] = (float)std::atof(argv[8 + i]);
            size = (float)std::atof(argv[13]);
            CamParam.setParams(K, D, aruco::Size(W, H));
            CamParam.resize(aruco::Size(W, H));
        }
        MDetector.detect(aruco::ImageView(img.data(), H, W), Markers, CamParam, size);  // utils/aruco_simple.cpp:77
        {   // the reference holds detectors by value (boarddetector.h:146): a copy must detect the same markers
            aruco::MarkerDetector copy(MDetector);
            std::vector<aruco::Marker> again;
            copy.detect(aruco::ImageView(img.data(), H, W), again, CamParam, size);
            bool same = again.size() == Markers.size();
            for (size_t i = 0; same && i < again.size(); i++)
                same = again[i].id == Markers[i].id && again[i][0].x == Markers[i][0].x && again[i][2].y == Markers[i][2].y;
            if (!same) { std::fprintf(stderr, "a copied MarkerDetector detects differently\n"); return 4; }
            // a rejected setter value must not poison the next call
            bool threw2 = false;
            try { copy.setThresholdParamRange(99); } catch (const aruco::Exception&) { threw2 = true; }
            copy.setThresholdParams(7, 7);
            if (!threw2) { std::fprintf(stderr, "setThresholdParamRange(99) did not throw\n"); return 5; }
        }
        for (const auto& m : Markers) {
            std::printf("%d", m.id);
            for (const auto& p : m) std::printf(" %.9g %.9g", p.x, p.y);
            std::printf(" %d", (int)m.hasPose);
            for (int k = 0; k < 3; k++) std::printf(" %.17g", m.Rvec[k]);
            for (int k = 0; k < 3; k++) std::printf(" %.17g", m.Tvec[k]);
            std::printf("\n");
        }
        // error behaviour mirrors the reference's CV_Asserts
        bool threw = false;
        try { MDetector.setWarpSize(5); } catch (const aruco::Exception&) { threw = true; }
        if (!threw) { std::fprintf(stderr, "setWarpSize(5) did not throw\n"); return 3; }
    } catch (const aruco::Exception& e) {
        std::fprintf(stderr, "aruco::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
