// Compiles against the lvk-compat header exactly like a reference caller would (VSFilter.cpp:352-364 /
// FilterParser.tpp:48,60 style) and runs a few frames through lvk::StabilizationFilter.  Prints one line per frame:
//   <index> <empty|timestamp> <sum of output bytes>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../livevisionkit_b200/compat/lvk/lvk.hpp"

int main(int argc, char** argv)
{
    const int w = 640, h = 360, frames = argc > 1 ? std::atoi(argv[1]) : 14;
    int failures = 0;
    lvk::context::assert_handler = [&](std::string file, std::string function, std::string assertion) {
        std::fprintf(stderr, "assert: %s %s %s\n", file.c_str(), function.c_str(), assertion.c_str());
        failures++;
    };

    lvk::StabilizationFilterSettings settings;           // defaults of the reference
    settings.detection_resolution = {480, 270};          // OBS "Homography" preset (VSFilter.cpp:269-280)
    settings.detection_regions = {2, 1};
    settings.max_feature_density = 0.12f;
    settings.min_feature_density = 0.04f;
    settings.accumulation_rate = 3.0f;
    settings.track_local_motions = false;
    settings.acceptance_threshold = 3.0f;
    lvk::StabilizationFilter filter(settings);
    filter.reconfigure([](lvk::StabilizationFilterSettings& s) { s.crop_to_stable_region = false; });
    if (filter.frame_delay() != 10 || filter.alias() != "Stabilization Filter") return 2;

    std::vector<uint8_t> pixels(static_cast<size_t>(w) * h * 3);
    // deterministic textured frame, shifted by i pixels
    auto render = [&](int i, uint8_t* dst) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
            {
                const int xs = x + 2 * i, ys = y + i;
                const uint8_t v = static_cast<uint8_t>(((xs / 16 + ys / 16) % 2) * 120 + ((xs * 7 + ys * 13) % 61) + 40);
                uint8_t* p = dst + (static_cast<size_t>(y) * w + x) * 3;
                p[0] = v; p[1] = static_cast<uint8_t>(v * 9 / 10); p[2] = static_cast<uint8_t>(v * 8 / 10);
            }
    };
    std::vector<std::pair<unsigned long long, unsigned long long>> applied;  // (timestamp, byte sum) of every output
    for (int i = 0; i < frames; i++)
    {
        render(i, pixels.data());
        lvk::VideoFrame input(pixels.data(), w, h, static_cast<size_t>(w) * 3, lvk::VideoFrame::BGR, 1000 + i);
        lvk::VideoFrame output;
        filter.apply(input, output, /*profile=*/i == frames - 1);
        if (output.empty())
            std::printf("%d empty 0\n", i);
        else
        {
            unsigned long long sum = 0;
            for (int y = 0; y < output.rows; y++)
                for (size_t b = 0; b < static_cast<size_t>(output.cols) * 3; b++) sum += output.data[y * output.step + b];
            std::printf("%d %llu %llu\n", i, static_cast<unsigned long long>(output.timestamp), sum);
            applied.emplace_back(output.timestamp, sum);
            if (output.format != lvk::VideoFrame::BGR) return 3;
        }
    }
    // error behaviour: a failed precondition reaches the assert handler instead of throwing
    lvk::StabilizationFilterSettings bad = settings;
    bad.min_tracking_quality = 2.0f;  // LVK_ASSERT_01 (StabilizationFilter.cpp:44)
    const int before = failures;
    filter.configure(bad);
    if (failures != before + 1) return 4;
    std::printf("timing_ms %.3f\n", filter.timings().elapsed().milliseconds());

    // ---- VideoFilter::stream (VideoFilter.cpp:62-209) on a fresh filter: the pipelined device path must deliver exactly
    // the frames apply() produced, and a true return from the callback must terminate it
    {
        struct Capture  // stands in for cv::VideoCapture: read() fills a frame the reader owns
        {
            int next = 0, count = 0, w = 0, h = 0;
            const std::function<void(int, uint8_t*)>* render = nullptr;
            bool read(lvk::VideoFrame& f)
            {
                if (next >= count) return false;
                f.create(h, w);
                (*render)(next, f.data);
                f.format = lvk::VideoFrame::BGR;
                f.timestamp = 1000 + next;
                next++;
                return true;
            }
        };
        const std::function<void(int, uint8_t*)> render_fn = render;
        lvk::StabilizationFilter streamed(settings);
        streamed.reconfigure([](lvk::StabilizationFilterSettings& s) { s.crop_to_stable_region = false; });
        Capture cap{0, frames, w, h, &render_fn};
        size_t k = 0;
        bool same = true;
        streamed.stream(cap, [&](lvk::Frame& out) {
            unsigned long long sum = 0;
            for (int y = 0; y < out.rows; y++)
                for (size_t b = 0; b < static_cast<size_t>(out.cols) * 3; b++) sum += out.data[y * out.step + b];
            same = same && k < applied.size() && applied[k].first == out.timestamp && applied[k].second == sum;
            k++;
            return false;
        });
        if (!same || k != applied.size()) { std::fprintf(stderr, "stream: %zu outputs, expected %zu\n", k, applied.size()); return 10; }
        lvk::StabilizationFilter stopped(settings);
        Capture cap2{0, frames, w, h, &render_fn};
        size_t seen = 0;
        stopped.stream(cap2, [&](lvk::Frame&) { return ++seen >= 2; });
        if (seen != 2) return 11;
        std::printf("stream %zu\n", k);
    }
    render(frames - 1, pixels.data());

    // ---- the editor's other filters on the last frame (FilterParser.tpp style): Scaling, Deblocking, a Composite chain
    auto checksum = [](const lvk::VideoFrame& f) {
        unsigned long long sum = 0;
        for (int y = 0; y < f.rows; y++)
            for (size_t b = 0; b < static_cast<size_t>(f.cols) * 3; b++) sum += static_cast<unsigned long long>(f.data[y * f.step + b]) * (1 + (b + y) % 7);
        return sum;
    };
    {
        lvk::ScalingFilter scaler(lvk::ScalingFilterSettings{{960, 540}, 0.8f, false});
        lvk::VideoFrame input(pixels.data(), w, h, static_cast<size_t>(w) * 3, lvk::VideoFrame::BGR, 77), output;
        scaler.apply(input, output);
        if (output.empty() || output.cols != 960 || output.rows != 540 || output.timestamp != 77) return 5;
        std::printf("scaling %llu\n", checksum(output));

        lvk::DeblockingFilter deblocker;
        lvk::VideoFrame output2;
        deblocker.apply(input, output2);
        if (output2.empty() || output2.cols != w || deblocker.filter_region().width != w / 16 * 16) return 6;
        std::printf("deblocking %llu\n", checksum(output2));

        lvk::CompositeFilter chain({std::make_shared<lvk::DeblockingFilter>(),
                                    std::make_shared<lvk::ScalingFilter>(lvk::ScalingFilterSettings{{960, 540}, 0.8f, false})});
        lvk::VideoFrame output3;
        chain.apply(input, output3);
        if (output3.empty() || output3.cols != 960 || chain.filter_count() != 2) return 7;
        std::printf("composite %llu\n", checksum(output3));
        chain.disable_filter(0);
        lvk::VideoFrame output4, output5;
        chain.apply(input, output4);
        scaler.apply(input, output5);  // (the in-place deblockers above have modified `pixels`: scale them again)
        if (checksum(output4) != checksum(output5)) return 8;  // with the deblocker disabled the chain is the scaler alone

        const int before2 = failures;
        scaler.reconfigure([](lvk::ScalingFilterSettings& s) { s.sharpness = 1.5f; });  // LVK_ASSERT_01 (ScalingFilter.cpp:43)
        if (failures != before2 + 1) return 9;
    }
    return 0;
}
