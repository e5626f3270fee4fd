// Inert stand-ins for the LiveVisionKit library's debug-HUD helpers (Functions/Drawing.hpp: colour tables indexed by
// frame format, draw_text / draw_rect).  They are OpenCL drawing kernels off the accelerated path; a real build keeps
// the reference's own.  Included behind lvk-compat by the redirect header of the OBS-plugin compile test.
#pragma once
#include <string>

namespace lvk
{
namespace col
{
    const cv::Scalar BLACK[6] = {}, WHITE[6] = {}, MAGENTA[6] = {}, GREEN[6] = {}, BLUE[6] = {}, RED[6] = {};
    inline cv::Scalar rgb2yuv(const cv::Scalar& rgb) { return rgb; }
}
template <typename T>
inline void draw_rect(VideoFrame&, const cv::Rect_<T>&, const cv::Scalar&, const int = 3) {}
inline void draw_text(VideoFrame&, const std::string&, const cv::Point&, const cv::Scalar&, const double = 1.5, const int = 2) {}
}  // namespace lvk
