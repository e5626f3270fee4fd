// Stubs for everything the reference's call sites reference that is NOT part of the drop-in boundary (libobs, the
// plugin's interop/drawing helpers, the editor's loggers).  tests/test_compat_cpu.py prepends this header to lines
// extracted at test time from /root/reference (VSFilter.cpp / VSFilter.hpp / VideoProcessor.cpp) and compiles the
// result against lvk-compat (-DLVK_COMPAT_USE_OPENCV, mock opencv2/): the lvk:: calls in those lines must compile
// UNCHANGED.  Test infrastructure only.
#pragma once
#define LVK_COMPAT_USE_OPENCV
#include "../../livevisionkit_b200/compat/lvk/lvk.hpp"

#include <array>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <optional>
#include <sstream>
#include <string>

// ---- libobs ---------------------------------------------------------------------------------------------------------
struct obs_data_t {};
struct obs_source_t {};
struct obs_properties_t;
struct obs_property_t;
struct obs_video_info { uint32_t fps_num = 60, fps_den = 1; };
inline bool obs_data_get_bool(obs_data_t*, const char*) { return false; }
inline double obs_data_get_double(obs_data_t*, const char*) { return 5.0; }
inline long long obs_data_get_int(obs_data_t*, const char*) { return 0; }
inline const char* obs_data_get_string(obs_data_t*, const char*) { return ""; }
inline void obs_data_set_int(obs_data_t*, const char*, long long) {}
inline void obs_source_update_properties(obs_source_t*) {}
inline bool obs_get_video_info(obs_video_info*) { return true; }

#define LVK_PROFILE
#define L(text) text

namespace lvk
{
// ---- Modules/OBS-Plugin/Interop + Utility ---------------------------------------------------------------------------------
struct OBSFrame : public VideoFrame { using VideoFrame::VideoFrame; };
namespace col
{
    inline cv::Scalar rgb2yuv(const cv::Scalar& rgb) { return rgb; }
    const std::array<cv::Scalar, 7> GREEN{}, RED{}, MAGENTA{};
}
inline void draw_text(VideoFrame&, const std::string&, const cv::Point&, const cv::Scalar&) {}
inline void draw_rect(VideoFrame&, const cv::Rect&, const cv::Scalar&) {}
class VisionFilter
{
public:
    explicit VisionFilter(obs_source_t*) {}
    virtual ~VisionFilter() = default;
    VideoFrame::Format format() const { return VideoFrame::YUV; }
    bool is_asynchronous() const { return true; }
protected:
    virtual void filter(OBSFrame& frame) = 0;
};

// ---- Modules/VideoEditor ----------------------------------------------------------------------------------------------------
struct ConsoleLogger
{
    struct NextTag {};
    static constexpr NextTag Next{};
    std::ostringstream text;
    template <typename T> ConsoleLogger& operator<<(const T& value) { text << value; return *this; }
    ConsoleLogger& operator<<(NextTag) { text << '\n'; return *this; }
    ConsoleLogger& operator<<(std::ios_base& (*manip)(std::ios_base&)) { text << manip; return *this; }
};
struct CSVLogger
{
    bool has_started() const { return started; }
    template <typename T> CSVLogger& operator<<(const T&) { started = true; return *this; }
    void next() {}
    bool started = false;
};
class VideoProcessor
{
public:
    void print_filter_timings();
    void log_timing_data();
private:
    ConsoleLogger m_ConsoleLogger;
    std::optional<CSVLogger> m_DataLogger = CSVLogger{};
    CompositeFilter m_Processor;
    TickTimer m_FrameTimer;
};
}  // namespace lvk
