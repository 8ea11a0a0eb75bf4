// aruco::FiducidalMarkers generators on the C++ facade (src/arucofidmarkers.h:49-88): the pixels are drawn on the
// device through the C ABI (ab_create_marker_image / ab_create_board_image).  Images are returned as tightly packed
// 8-bit buffers; the text watermark of the reference (cv::putText) is not reproduced and the board generators take the
// marker ids from the caller instead of rand().
#pragma once
#include "serialization.hpp"

namespace aruco {

struct Image8 {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    ImageView view() const { return ImageView(data.data(), rows, cols); }
};

class FiducidalMarkers {
public:
    // createMarkerImage(id, size, addWaterMark = false, locked) -- arucofidmarkers.cpp:213
    static Image8 createMarkerImage(MarkerDetector& md, int id, int size, bool locked = false) {
        int side = 0;
        check(md, ab_create_marker_image(md.handle(), id, size, locked, nullptr, 0, &side));
        Image8 im;
        im.rows = im.cols = side;
        im.data.resize((size_t)side * side);
        check(md, ab_create_marker_image(md.handle(), id, size, locked, im.data.data(), (size_t)side, &side));
        return im;
    }
    // createBoardImage (:283), createBoardImage_ChessBoard (:337), createBoardImage_Frame (:397)
    static Image8 createBoardImage(MarkerDetector& md, Size gridSize, int MarkerSize, int MarkerDistance, BoardConfiguration& TInfo,
                                   const std::vector<int>& ids) {
        return board(md, 0, gridSize, MarkerSize, MarkerDistance, true, TInfo, ids);
    }
    static Image8 createBoardImage_ChessBoard(MarkerDetector& md, Size gridSize, int MarkerSize, BoardConfiguration& TInfo,
                                              const std::vector<int>& ids, bool centerData = true) {
        return board(md, 1, gridSize, MarkerSize, 0, centerData, TInfo, ids);
    }
    static Image8 createBoardImage_Frame(MarkerDetector& md, Size gridSize, int MarkerSize, int MarkerDistance, BoardConfiguration& TInfo,
                                         const std::vector<int>& ids, bool centerData = true) {
        return board(md, 2, gridSize, MarkerSize, MarkerDistance, centerData, TInfo, ids);
    }

private:
    static void check(MarkerDetector& md, int rc) {
        if (rc != AB_OK) throw Exception(rc, ab_last_error(md.handle()));
    }
    static Image8 board(MarkerDetector& md, int kind, Size g, int ms, int dist, bool center, BoardConfiguration& TInfo, const std::vector<int>& ids) {
        int w = 0, h = 0, n = 0;
        check(md, ab_create_board_image(md.handle(), kind, g.width, g.height, ms, dist, center, nullptr, 0, nullptr, 0, &w, &h, nullptr, nullptr, 0, &n));
        std::vector<int32_t> in(ids.begin(), ids.end()), out_ids((size_t)n);
        std::vector<float> corners((size_t)n * 12);
        Image8 im;
        im.rows = h;
        im.cols = w;
        im.data.resize((size_t)w * h);
        check(md, ab_create_board_image(md.handle(), kind, g.width, g.height, ms, dist, center, in.data(), (int)in.size(), im.data.data(), (size_t)w,
                                        &w, &h, out_ids.data(), corners.data(), n, &n));
        TInfo.mInfoType = BoardConfiguration::PIX;
        TInfo.ids.assign(out_ids.begin(), out_ids.end());
        TInfo.objPoints.assign((size_t)n, std::array<float, 12>());
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 12; k++) TInfo.objPoints[(size_t)i][(size_t)k] = corners[(size_t)i * 12 + k];
        return im;
    }
};

}  // namespace aruco
