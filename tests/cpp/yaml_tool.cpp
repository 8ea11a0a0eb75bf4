// Exercises include/aruco/serialization.hpp without a GPU (tests/test_yaml.py):
//   yaml_tool <camera|markers|board|boardconf|dict> <in.yml> <out.yml>
// parses <in.yml> the way the reference's cv::FileStorage readers do, prints every value on stdout with full
// precision, and writes the object back to <out.yml> in the reference's wire format.
#include <cstdio>
#include <string>
#include "../../include/aruco/serialization.hpp"

static void print_marker(const aruco::Marker& m) {
    std::printf("marker %d %d", m.id, (int)m.hasPose);
    for (int k = 0; k < 3; k++) std::printf(" %.17g", m.Rvec[k]);
    for (int k = 0; k < 3; k++) std::printf(" %.17g", m.Tvec[k]);
    std::printf(" %zu", m.size());
    for (const auto& p : m) std::printf(" %.9g %.9g", p.x, p.y);
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: yaml_tool kind in.yml out.yml\n"); return 2; }
    const std::string kind = argv[1], in = argv[2], out = argv[3];
    try {
        if (kind == "camera") {
            aruco::CameraParameters cp = aruco::readCameraParameters(in);
            std::printf("camera %d %d", cp.CamSize.width, cp.CamSize.height);
            for (int i = 0; i < 9; i++) std::printf(" %.9g", cp.CameraMatrix[i]);
            for (int i = 0; i < 5; i++) std::printf(" %.9g", cp.Distorsion[i]);
            std::printf("\n");
            aruco::saveCameraParameters(cp, out);
        } else if (kind == "markers") {
            std::vector<aruco::Marker> ms = aruco::readMarkers(in);
            for (const auto& m : ms) print_marker(m);
            aruco::saveMarkers(ms, out);
        } else if (kind == "board") {
            aruco::Board b = aruco::readBoard(in);
            std::printf("board %d", (int)b.hasPose);
            for (int k = 0; k < 3; k++) std::printf(" %.17g", b.Rvec[k]);
            for (int k = 0; k < 3; k++) std::printf(" %.17g", b.Tvec[k]);
            std::printf("\n");
            for (const auto& m : b) print_marker(m);
            aruco::saveBoard(b, out);
        } else if (kind == "boardconf") {
            aruco::BoardConfiguration bc;
            bc.readFromFile(in);
            std::printf("boardconf %zu %d\n", bc.size(), bc.mInfoType);
            for (size_t i = 0; i < bc.size(); i++) {
                std::printf("bm %d", bc.ids[i]);
                for (int k = 0; k < 12; k++) std::printf(" %.9g", bc.objPoints[i][k]);
                std::printf("\n");
            }
            bc.saveToFile(out);
        } else if (kind == "dict") {
            aruco::Dictionary d;
            d.fromFile(in);
            std::printf("dict %zu %d %d\n", d.size(), d.markersize, d.tau0);
            for (const auto& c : d.codes) std::printf("code %s\n", c.c_str());
            d.toFile(out);
        } else {
            return 2;
        }
    } catch (const aruco::Exception& e) {
        std::fprintf(stderr, "aruco::Exception %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
