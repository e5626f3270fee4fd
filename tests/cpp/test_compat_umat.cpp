// lvk-compat built with the reference's own VideoFrame declaration (struct VideoFrame : cv::UMat, mock opencv2/) and
// driven the way the OBS plugin drives it (VSFilter.cpp:352-364): apply(std::move(frame), frame) on one frame object,
// plus VideoFilter::stream(cv::VideoCapture&, callback).  Prints "<index> <empty|timestamp> <sum of output bytes>".
#define LVK_COMPAT_USE_OPENCV
#include "../../livevisionkit_b200/compat/lvk/lvk.hpp"

#include <cstdio>
#include <cstdlib>

static void render(int i, int w, int h, cv::UMat& frame)
{
    frame.create(h, w, CV_8UC3);
    cv::Mat m = frame.getMat(cv::ACCESS_WRITE);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            const int xs = x + 2 * i, ys = y + i;
            const uint8_t v = static_cast<uint8_t>(((xs / 16 + ys / 16) % 2) * 120 + ((xs * 7 + ys * 13) % 61) + 40);
            uint8_t* p = m.data + static_cast<size_t>(y) * m.step + static_cast<size_t>(x) * 3;
            p[0] = v; p[1] = static_cast<uint8_t>(v * 9 / 10); p[2] = static_cast<uint8_t>(v * 8 / 10);
        }
}

static unsigned long long total(const lvk::VideoFrame& f)
{
    cv::Mat m = f.getMat(cv::ACCESS_READ);
    unsigned long long sum = 0;
    for (int y = 0; y < m.rows; y++)
        for (size_t b = 0; b < static_cast<size_t>(m.cols) * 3; b++) sum += m.data[y * m.step + b];
    return sum;
}

int main(int argc, char** argv)
{
    const int w = 640, h = 360, frames = argc > 1 ? std::atoi(argv[1]) : 14;
    lvk::StabilizationFilterSettings settings;
    settings.detection_resolution = {480, 270};  // OBS "Homography" preset (VSFilter.cpp:269-280)
    settings.detection_regions = {2, 1};
    settings.max_feature_density = 0.12f;
    settings.min_feature_density = 0.04f;
    settings.accumulation_rate = 3.0f;
    settings.track_local_motions = false;
    settings.acceptance_threshold = 3.0f;
    lvk::StabilizationFilter filter(settings);
    filter.set_timing_samples(30);
    for (int i = 0; i < frames; i++)
    {
        lvk::VideoFrame frame;
        render(i, w, h, frame);
        frame.timestamp = 1000 + i;
        frame.format = lvk::VideoFrame::BGR;
        filter.apply(std::move(frame), frame, i % 2 == 0);  // input and output are the same object, as in OBS
        if (frame.empty()) std::printf("%d empty 0\n", i);
        else std::printf("%d %llu %llu\n", i, static_cast<unsigned long long>(frame.timestamp), total(frame));
    }
    std::printf("timings %zu %.4f %.4f\n", filter.timings().history().size(), filter.timings().average().milliseconds(),
                filter.timings().deviation().milliseconds());

    cv::VideoCapture capture;  // the mock hands out the frames pushed into it
    for (int i = 0; i < frames; i++)
    {
        cv::UMat f;
        render(i, w, h, f);
        capture.push(f);
    }
    lvk::StabilizationFilter streamed(settings);
    int delivered = 0;
    streamed.stream(capture, [&](lvk::Frame& out) {
        std::printf("s%d %llu %llu\n", delivered, static_cast<unsigned long long>(out.timestamp), total(out));
        delivered++;
        return false;
    });
    std::printf("stream %d\n", delivered);
    return 0;
}
