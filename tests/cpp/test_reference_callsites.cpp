// Compile-and-link check of the drop-in boundary against the API the reference's two callers use, with the reference's
// own VideoFrame declaration (struct VideoFrame : cv::UMat, -DLVK_COMPAT_USE_OPENCV) on a mock opencv2/.  The calls
// below are the ones VSFilter.cpp (OBS plugin) and VideoProcessor.cpp (video editor) make - re-typed here, not copied;
// tests/test_compat_cpu.py additionally compiles the reference's real lines where /root/reference is present.
//   Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:235-298 (reconfigure / settings / frame_delay),
//   :346-364 (set_timing_samples, apply(std::move(frame), frame[, true]), draw_*), :372-383 (timings, stable_region)
//   Modules/VideoEditor/VideoProcessor.cpp:296-353 (filters(i)->alias() / timings().average() / deviation() / frequency())
#define LVK_COMPAT_USE_OPENCV
#include "../../livevisionkit_b200/compat/lvk/lvk.hpp"

#include <cmath>
#include <iostream>
#include <sstream>

namespace
{
struct OBSFrame : public lvk::VideoFrame  // Modules/OBS-Plugin/Interop/OBSFrame.hpp: a VideoFrame the plugin owns
{
    using lvk::VideoFrame::VideoFrame;
};

class StabilisationSource
{
public:
    StabilisationSource() { m_Filter.set_timing_samples(30); }

    int configure(bool apply_crop, bool disabled, float crop_x, float crop_y, int samples, bool field_subsystem, float fps)
    {
        m_Filter.reconfigure([&](lvk::StabilizationFilterSettings& stab_settings) {
            stab_settings.crop_to_stable_region = apply_crop && !m_TestMode;
            stab_settings.stabilize_output = !disabled;
            stab_settings.corrective_limits.height = crop_y;
            stab_settings.corrective_limits.width = crop_x;
            stab_settings.predictive_samples = static_cast<size_t>(samples);
            stab_settings.background_colour[0] = 16.0;
            stab_settings.background_colour = cv::Scalar(16, 128, 128);
            stab_settings.detection_resolution = {480, 270};
            stab_settings.motion_resolution = field_subsystem ? cv::Size(16, 16) : cv::Size(2, 2);
            stab_settings.detection_regions = {2, 1};
            stab_settings.track_local_motions = field_subsystem;
            stab_settings.acceptance_threshold = 3.0f;
            stab_settings.min_scene_quality = 0.95f;
            stab_settings.min_tracking_quality = 0.35f;
        });
        const auto delay_ms = static_cast<int>(std::round((1000.0f / fps) * static_cast<float>(m_Filter.frame_delay())));
        std::ostringstream log;
        log << m_Filter.settings().predictive_samples << m_Filter.settings().corrective_limits.width
            << m_Filter.settings().crop_to_stable_region << m_Filter.settings().stabilize_output;
        return delay_ms + static_cast<int>(log.str().size());
    }

    void filter(OBSFrame& frame)
    {
        if (m_TestMode)
        {
            m_Filter.apply(std::move(frame), frame, true);
            m_Filter.draw_motion_mesh();
            m_Filter.draw_trackers();
            hud(frame);
        }
        else m_Filter.apply(std::move(frame), frame);
    }

    std::string hud(OBSFrame& frame)
    {
        const double frame_time_ms = m_Filter.timings().average().milliseconds();
        const double deviation_ms = m_Filter.timings().deviation().milliseconds();
        const auto& crop_region = m_Filter.stable_region();
        const cv::Point anchor = crop_region.tl() + cv::Point(5, 40);
        return cv::format("%.2fms (%.2fms) @%d,%d %d", frame_time_ms, deviation_ms, anchor.x, anchor.y, static_cast<int>(frame.format));
    }

    bool m_TestMode = false;

private:
    lvk::StabilizationFilter m_Filter;
};

std::string print_filter_timings(lvk::CompositeFilter& processor)
{
    std::ostringstream out;
    for (size_t i = 0; i < processor.filter_count(); i++)
    {
        auto filter = processor.filters(i);
        auto average_timing = filter->timings().average();
        out << std::to_string(i) << ".   " << filter->alias() << "\t" << average_timing.milliseconds() << "ms"
            << " +/- " << filter->timings().deviation().milliseconds() << "ms"
            << "   (" << static_cast<uint64_t>(average_timing.frequency()) << "FPS) uid " << filter->uid() << "\n";
    }
    for (auto& filter : processor.filters()) out << filter->timings().history().size() << filter->timings().elapsed().hms();
    return out.str();
}
}  // namespace

int main(int argc, char**)
{
    if (argc > 100)  // never taken: this translation unit is a compile-and-link check (running needs a GPU)
    {
        StabilisationSource source;
        source.configure(true, false, 0.05f, 0.05f, 10, false, 60.0f);
        OBSFrame frame;
        frame.create(360, 640, CV_8UC3);
        frame.format = lvk::VideoFrame::YUV;
        source.filter(frame);
        lvk::CompositeFilter processor({std::make_shared<lvk::DeblockingFilter>(), std::make_shared<lvk::StabilizationFilter>()});
        std::cout << print_filter_timings(processor);
        cv::VideoCapture capture;
        lvk::StabilizationFilter streaming;
        streaming.stream(capture, [](lvk::Frame& f) { return f.empty(); });
        lvk::Stopwatch watch(4);
        watch.sync_gpu().start();
        watch.wait_until(lvk::Time::Milliseconds(0.01));
        watch.set_history_size(8);
        std::cout << watch.restart().microseconds() << lvk::Time::Timestep(60.0).frequency() << lvk::Time::Timestamp();
        lvk::VideoFrame view = frame(cv::Rect(0, 0, 16, 16)), deep = frame.clone();
        frame.copyTo(deep);
        std::cout << view.cols << deep.rows << lvk::Unique<>().uid();
    }
    return 0;
}
