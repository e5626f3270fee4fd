// lvk-compat — C++ mirror of the reference's filter interface for the stabilization path, forwarding to the C-ABI
// (include/lvkb200.h).  Same names, argument meaning and error behaviour as
//   LiveVisionKit/Filters/VideoFilter.hpp:32-61          lvk::VideoFilter (apply / stream / alias / timings / filter)
//   LiveVisionKit/Filters/StabilizationFilter.hpp:28-77  lvk::StabilizationFilterSettings, lvk::StabilizationFilter
//   LiveVisionKit/Filters/DeblockingFilter.hpp:26-59     lvk::DeblockingFilterSettings, lvk::DeblockingFilter
//   LiveVisionKit/Filters/ScalingFilter.hpp:27-52        lvk::ScalingFilterSettings, lvk::ScalingFilter
//   LiveVisionKit/Filters/CompositeFilter.hpp:27-75      lvk::CompositeFilterSettings, lvk::CompositeFilter
//   LiveVisionKit/Vision/FrameTracker.hpp:31-44, FeatureDetector.hpp:28-37, PathSmoother.hpp:29-39  settings bases
//   LiveVisionKit/Utility/Configurable.hpp:27-44         lvk::Configurable<T>
//   LiveVisionKit/Utility/Unique.hpp:25-45               lvk::Unique<Scope>
//   LiveVisionKit/Timing/Time.hpp:25-108, Stopwatch.hpp:25-75, Data/StreamBuffer.hpp  lvk::Time, lvk::Stopwatch, history
//   LiveVisionKit/Data/VideoFrame.hpp:25-79              lvk::VideoFrame (format, timestamp, width/height)
//   LiveVisionKit/Directives.hpp:37-95                   lvk::context::assert_handler, LVK_ASSERT
// Two builds of lvk::VideoFrame:
//   * default (no OpenCV in the build): VideoFrame owns / views a plain packed 8-bit buffer in host or CUDA device
//     memory, reference-counted like a cv::UMat (copies are shallow, clone() is deep); cv::Size / Size2f / Scalar /
//     Rect / Point are replaced by minimal structs in lvk::cvlite;
//   * -DLVK_COMPAT_USE_OPENCV: `struct VideoFrame : public cv::UMat` exactly as Data/VideoFrame.hpp:28 declares it, the
//     real cv:: geometry types, and VideoFilter::stream(cv::VideoCapture&, ...) — what the OBS plugin and the video
//     editor compile against.  Pixels reach the library through cv::UMat::getMat (host mapping).
//     tests/test_compat_cpu.py compiles the reference's own call sites (VSFilter.cpp, VideoProcessor.cpp) against this
//     build with a mock opencv2/ (tests/cpp/mock_opencv).
// Header-only; link with liblvkb200.so.
#pragma once

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <initializer_list>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../../include/lvkb200.h"

#ifdef LVK_COMPAT_USE_OPENCV
#include <opencv2/core.hpp>
#include <opencv2/videoio.hpp>
#endif

namespace lvk
{

#ifdef LVK_COMPAT_USE_OPENCV
namespace cvlite { using Size = cv::Size; using Size2f = cv::Size2f; using Scalar = cv::Scalar; using Rect = cv::Rect; using Point = cv::Point; }
#else
namespace cvlite
{
    struct Point { int x = 0, y = 0; Point() = default; Point(int px, int py) : x(px), y(py) {}
                   Point operator+(const Point& o) const { return {x + o.x, y + o.y}; } };
    struct Size { int width = 0, height = 0; Size() = default; Size(int w, int h) : width(w), height(h) {}
                  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
                  bool operator!=(const Size& o) const { return !(*this == o); } };
    struct Size2f { float width = 0, height = 0; Size2f() = default; Size2f(float w, float h) : width(w), height(h) {} };
    struct Scalar { double val[4] = {0, 0, 0, 0}; Scalar() = default;
                    Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
                    double operator[](int i) const { return val[i]; } double& operator[](int i) { return val[i]; } };
    struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() = default;
                  Rect(int px, int py, int w, int h) : x(px), y(py), width(w), height(h) {}
                  Point tl() const { return {x, y}; } Point br() const { return {x + width, y + height}; } };
}
#endif

// ---- Directives.hpp:37-95 ----------------------------------------------------------------------------------------
namespace context
{
    // Directives.cpp:27-42: the default prints and aborts
    inline std::function<void(std::string file, std::string function, std::string assertion)> assert_handler =
        [](std::string file, std::string function, std::string assertion) {
            std::fprintf(stderr, "LiveVisionKit failed %s@%s(..) ` %s ` \n", file.c_str(), function.c_str(), assertion.c_str());
            std::abort();
        };
}

// The assertion family of Directives.hpp:46-95: a failed check reports (file, function, text) to the handler and carries
// on; the range forms report the violated relation ("0 <= x <= 1", "lo < x < hi") like the reference does.
#ifndef LVK_DISABLE_CHECKS
#define LVK_COMPAT_CHECK(passes, text) \
    if (!(passes)) { lvk::context::assert_handler(__FILE__, __func__, text); }
#else
#define LVK_COMPAT_CHECK(passes, text)
#endif
#define LVK_ASSERT(assertion) LVK_COMPAT_CHECK(assertion, #assertion)
#define LVK_ASSERT_IF(condition, assertion) LVK_COMPAT_CHECK(!(condition) || (assertion), #assertion)
#define LVK_ASSERT_01(value) LVK_COMPAT_CHECK(!(value < 0 || value > 1), "0 <= " #value " <= 1")
#define LVK_ASSERT_01_STRICT(value) LVK_COMPAT_CHECK(!(value <= 0 || value >= 1), "0 < " #value " < 1")
#define LVK_ASSERT_RANGE(value, min, max) LVK_COMPAT_CHECK(!(value < min || value > max), #min " <= " #value " <= " #max)
#define LVK_ASSERT_RANGE_STRICT(value, min, max) LVK_COMPAT_CHECK(!(value <= min || value >= max), #min " < " #value " < " #max)

// ---- Utility/Unique.hpp:25-45, Unique.tpp:27-60 -----------------------------------------------------------------------
struct GlobalScope;
template <typename Scope = GlobalScope>
class Unique
{
public:
    Unique() : m_UID(next()) {}
    Unique(const Unique&) : m_UID(next()) {}           // a copy is a new object
    Unique(Unique&& other) noexcept : m_UID(other.m_UID) {}
    Unique& operator=(const Unique&) { return *this; }
    uint64_t uid() const { return m_UID; }
private:
    static uint64_t next() { static uint64_t s_NextUID = 1; return s_NextUID++; }
    uint64_t m_UID;
};

// ---- Timing/Time.hpp:25-108 (Time.cpp) -----------------------------------------------------------------------------------
class Time
{
public:
    using TimePoint = std::chrono::high_resolution_clock::time_point;
    static Time Now() { return Time(std::chrono::high_resolution_clock::now()); }
    static std::string Timestamp(const char* format = "%F %T")
    {
        const std::time_t now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
        char buffer[128] = {0};
        std::strftime(buffer, sizeof(buffer), format, std::localtime(&now));
        return buffer;
    }
    static Time Hours(const double amount) { return Seconds(amount * 3600.0); }
    static Time Minutes(const double amount) { return Seconds(amount * 60.0); }
    static Time Seconds(const double amount) { return Time(static_cast<uint64_t>(amount * 1e9)); }
    static Time Milliseconds(const double amount) { return Time(static_cast<uint64_t>(amount * 1e6)); }
    static Time Microseconds(const double amount) { return Time(static_cast<uint64_t>(amount * 1e3)); }
    static Time Nanoseconds(const uint64_t amount) { return Time(amount); }
    static Time Timestep(const double frequency) { return Seconds(1.0 / frequency); }

    Time() : m_Time(0) {}
    explicit Time(const TimePoint& time) : m_Time(std::chrono::duration_cast<std::chrono::nanoseconds>(time.time_since_epoch())) {}
    explicit Time(const uint64_t nanoseconds) : m_Time(static_cast<std::chrono::nanoseconds::rep>(nanoseconds)) {}
    explicit Time(const std::chrono::nanoseconds nanoseconds) : m_Time(nanoseconds) {}
    Time(const Time& other) = default;

    double hours() const { return seconds() / 3600.0; }
    double minutes() const { return seconds() / 60.0; }
    double seconds() const { return nanoseconds() * 1e-9; }
    double milliseconds() const { return nanoseconds() * 1e-6; }
    double microseconds() const { return nanoseconds() * 1e-3; }
    double nanoseconds() const { return static_cast<double>(m_Time.count()); }
    double frequency() const { return 1.0 / seconds(); }
    std::string hms() const
    {
        const uint64_t total = static_cast<uint64_t>(seconds());
        char buffer[32];
        std::snprintf(buffer, sizeof(buffer), "%02llu:%02llu:%02llu", static_cast<unsigned long long>(total / 3600),
                      static_cast<unsigned long long>((total / 60) % 60), static_cast<unsigned long long>(total % 60));
        return buffer;
    }
    bool is_zero() const { return m_Time.count() == 0; }

    Time& operator=(const Time& other) = default;
    void operator+=(const Time& other) { m_Time += other.m_Time; }
    void operator-=(const Time& other) { m_Time -= other.m_Time; }
    Time operator+(const Time& other) const { return Time(m_Time + other.m_Time); }
    Time operator-(const Time& other) const { return Time(m_Time - other.m_Time); }
    bool operator==(const Time& other) const { return m_Time == other.m_Time; }
    bool operator!=(const Time& other) const { return m_Time != other.m_Time; }
    bool operator>(const Time& other) const { return m_Time > other.m_Time; }
    bool operator>=(const Time& other) const { return m_Time >= other.m_Time; }
    bool operator<(const Time& other) const { return m_Time < other.m_Time; }
    bool operator<=(const Time& other) const { return m_Time <= other.m_Time; }
    Time operator*(const double multiplier) const { return Time(std::chrono::nanoseconds(static_cast<std::chrono::nanoseconds::rep>(nanoseconds() * multiplier))); }
    Time operator/(const double divisor) const { return Time(std::chrono::nanoseconds(static_cast<std::chrono::nanoseconds::rep>(nanoseconds() / divisor))); }
private:
    std::chrono::nanoseconds m_Time;
};

// ---- Data/StreamBuffer.hpp (the part Stopwatch::history() exposes: a fixed-capacity ring, oldest -> newest) ----------------
template <typename T>
class StreamBuffer
{
public:
    explicit StreamBuffer(const size_t capacity) : m_Capacity(capacity ? capacity : 1) {}
    void push(const T& element)  // a full buffer overwrites its oldest element (StreamBuffer.tpp:37-84)
    {
        if (m_Data.size() == m_Capacity) m_Data.erase(m_Data.begin());
        m_Data.push_back(element);
    }
    void resize(const size_t capacity)
    {
        m_Capacity = capacity ? capacity : 1;
        while (m_Data.size() > m_Capacity) m_Data.erase(m_Data.begin());
    }
    void clear() { m_Data.clear(); }
    T& at(const size_t index) { return m_Data.at(index); }
    const T& at(const size_t index) const { return m_Data.at(index); }
    T& operator[](const size_t index) { return m_Data[index]; }
    const T& operator[](const size_t index) const { return m_Data[index]; }
    const T& oldest(const int offset = 0) const { return m_Data[static_cast<size_t>(offset)]; }
    const T& newest(const int offset = 0) const { return m_Data[m_Data.size() - 1 + static_cast<size_t>(offset)]; }
    bool is_full() const { return m_Data.size() == m_Capacity; }
    bool is_empty() const { return m_Data.empty(); }
    size_t size() const { return m_Data.size(); }
    size_t capacity() const { return m_Capacity; }
    typename std::vector<T>::const_iterator begin() const { return m_Data.begin(); }
    typename std::vector<T>::const_iterator end() const { return m_Data.end(); }
private:
    std::vector<T> m_Data;
    size_t m_Capacity;
};

// ---- Timing/Stopwatch.hpp:25-75 (Stopwatch.cpp:27-166) -----------------------------------------------------------------------
class Stopwatch
{
public:
    explicit Stopwatch(const size_t history = 1) : m_History(history) {}
    void start() { m_Running = true; m_StartTime = Time::Now(); }
    Time stop()
    {
        if (is_running() || is_paused())
        {
            m_ElapsedTime = pause();
            m_History.push(m_ElapsedTime);
            m_Memory = Time(0);
            return m_ElapsedTime;
        }
        return Time(0);
    }
    Time pause()
    {
        if (!is_running()) return m_Memory;
        m_Memory += (Time::Now() - m_StartTime);
        m_ElapsedTime = m_Memory;
        m_Running = false;
        return m_Memory;
    }
    Time restart() { const Time elapsed = stop(); start(); return elapsed; }
    bool is_paused() const { return !m_Running && m_Memory.nanoseconds() > 0; }
    bool is_running() const { return m_Running; }
    Time wait_until(const Time& target_elapsed_time)
    {
        if (!is_running()) start();
        Time elapsed_time = elapsed();
        while (elapsed_time < target_elapsed_time)
        {
            std::this_thread::yield();
            elapsed_time = elapsed();
        }
        return elapsed_time;
    }
    // Stopwatch.cpp:127-131 calls cv::ocl::finish(): here every CUDA stream of the calling thread's device is drained
    Stopwatch& sync_gpu(const bool trigger = true)
    {
        if (trigger) lvkb200_device_synchronize();
        return *this;
    }
    Time elapsed() const { return is_running() ? (m_Memory + (Time::Now() - m_StartTime)) : m_ElapsedTime; }
    Time average() const
    {
        if (m_History.is_empty()) return Time(0);
        Time total(0);
        for (const Time& t : m_History) total += t;
        return total / static_cast<double>(m_History.size());
    }
    Time deviation() const  // mean absolute deviation, Stopwatch.cpp:142-160
    {
        if (m_History.size() < 2) return Time(0);
        const Time average_time = average();
        Time total_deviation(0);
        for (const Time& current_time : m_History)
            total_deviation += (average_time > current_time) ? (average_time - current_time) : (current_time - average_time);
        return total_deviation / static_cast<double>(m_History.size());
    }
    void reset_history() { m_History.clear(); }
    const StreamBuffer<Time>& history() const { return m_History; }
    void set_history_size(const size_t history) { m_History.resize(history); }
private:
    bool m_Running = false;
    StreamBuffer<Time> m_History;
    Time m_ElapsedTime{0}, m_StartTime{0}, m_Memory{0};
};

// ---- Timing/TickTimer.hpp:24-44: a Stopwatch that counts its laps (the editor's frame timer, VideoProcessor.hpp:68) ---
class TickTimer : public Stopwatch
{
public:
    explicit TickTimer(const uint32_t history = 1) : Stopwatch(history) { LVK_ASSERT(history > 0); }
    Time tick()  // closes the running lap (into the history) and opens the next one
    {
        ++m_Ticks;
        return m_LastLap = restart();
    }
    Time tick(const Time& timestep)  // fixed-rate ticking: the lap is stretched to at least `timestep`
    {
        wait_until(timestep);
        return tick();
    }
    uint64_t tick_count() const { return m_Ticks; }
    void reset_counter() { m_Ticks = 0; }
    Time delta() const { return m_LastLap; }
private:
    uint64_t m_Ticks = 0;
    Time m_LastLap{0};
};

// ---- Data/VideoFrame.hpp:25-79 -------------------------------------------------------------------------------------
#ifdef LVK_COMPAT_USE_OPENCV
// The reference's own declaration: a cv::UMat with a timestamp and a format.  (reformat / reformatTo / viewAsFormat are
// cv::cvtColor wrappers, VideoFrame.cpp:116-317, used by ConversionFilter and the drawing helpers - outside this path.)
struct VideoFrame : public cv::UMat
{
    enum Format { BGR, BGRA, RGB, RGBA, YUV, GRAY, UNKNOWN };

    uint64_t timestamp = 0;
    Format format = UNKNOWN;
    int& width = cols; int& height = rows;

    VideoFrame() : cv::UMat() {}
    VideoFrame(const VideoFrame& frame) : cv::UMat(frame), timestamp(frame.timestamp), format(frame.format) {}
    VideoFrame(VideoFrame&& frame) noexcept : cv::UMat(std::move(frame)), timestamp(frame.timestamp), format(frame.format) {}
    explicit VideoFrame(const uint64_t ts) : cv::UMat(), timestamp(ts) {}
    explicit VideoFrame(const cv::UMat& frame, const uint64_t ts = 0, const Format fmt = UNKNOWN) : cv::UMat(frame), timestamp(ts), format(fmt) {}
    explicit VideoFrame(cv::UMat&& frame, const uint64_t ts = 0, const Format fmt = UNKNOWN) noexcept : cv::UMat(std::move(frame)), timestamp(ts), format(fmt) {}
    virtual ~VideoFrame() = default;

    VideoFrame& operator=(VideoFrame&& frame) noexcept
    {
        timestamp = frame.timestamp; format = frame.format;
        cv::UMat::operator=(std::move(frame));
        return *this;
    }
    VideoFrame& operator=(const VideoFrame& frame) noexcept
    {
        timestamp = frame.timestamp; format = frame.format;
        cv::UMat::operator=(frame);
        return *this;
    }
    VideoFrame clone() const { return VideoFrame(cv::UMat::clone(), timestamp, format); }
    void copyTo(VideoFrame& dst) const { cv::UMat::copyTo(dst); dst.timestamp = timestamp; dst.format = format; }
    VideoFrame operator()(const cv::Rect& roi) const { return VideoFrame(cv::UMat::operator()(roi), timestamp, format); }
    bool has_known_format() const { return format != UNKNOWN; }
};
#else
struct VideoFrame
{
    enum Format { BGR, BGRA, RGB, RGBA, YUV, GRAY, UNKNOWN };

    uint8_t* data = nullptr;   // packed 8-bit pixels (3 channels for every format the stabilizer accepts)
    size_t step = 0;           // bytes per row
    int cols = 0, rows = 0;
    bool on_device = false;    // true: `data` is CUDA device memory (used in place, never copied through the host)
    uint64_t timestamp = 0;
    Format format = UNKNOWN;
    int& width = cols; int& height = rows;

    VideoFrame() = default;
    explicit VideoFrame(const uint64_t ts) : timestamp(ts) {}
    // view onto caller memory (no ownership)
    VideoFrame(uint8_t* pixels, int w, int h, size_t row_step, Format fmt, uint64_t ts = 0, bool device = false)
        : data(pixels), step(row_step), cols(w), rows(h), on_device(device), timestamp(ts), format(fmt) {}
    // Copies are SHALLOW, like the reference's cv::UMat base (VideoFrame.cpp:37-44, VideoFilter.cpp:55-58): owned pixels
    // are reference-counted and shared; clone() makes the deep copy.
    VideoFrame(const VideoFrame& o) { *this = o; }
    VideoFrame(VideoFrame&& o) noexcept { *this = std::move(o); }
    virtual ~VideoFrame() = default;
    VideoFrame& operator=(const VideoFrame& o) noexcept
    {
        if (this == &o) return *this;
        storage = o.storage; data = o.data;
        step = o.step; cols = o.cols; rows = o.rows; on_device = o.on_device; timestamp = o.timestamp; format = o.format;
        return *this;
    }
    VideoFrame& operator=(VideoFrame&& o) noexcept
    {
        if (this == &o) return *this;
        storage = std::move(o.storage); data = o.data;
        step = o.step; cols = o.cols; rows = o.rows; on_device = o.on_device; timestamp = o.timestamp; format = o.format;
        o.release();
        return *this;
    }

    bool empty() const { return data == nullptr || cols <= 0 || rows <= 0; }
    void release() { storage.reset(); data = nullptr; step = 0; cols = rows = 0; on_device = false; }
    // cv::UMat::create(rows, cols, CV_8UC3): (re)allocates an owned host buffer when the geometry differs or the pixels are
    // shared with another frame
    void create(int h, int w, int channels = 3)
    {
        if (storage && storage.use_count() == 1 && h == rows && w == cols && step == static_cast<size_t>(w) * channels) return;
        storage = std::make_shared<std::vector<uint8_t>>(static_cast<size_t>(w) * channels * h, 0);
        data = storage->data(); step = static_cast<size_t>(w) * channels; cols = w; rows = h; on_device = false;
    }
    VideoFrame clone() const  // deep copy (host frames; a device view is cloned as the same view)
    {
        VideoFrame c;
        if (empty() || on_device) { c = *this; return c; }
        c.create(rows, cols, static_cast<int>(step / static_cast<size_t>(cols)) >= 3 ? 3 : 1);
        copyTo(c);
        return c;
    }
    void copyTo(VideoFrame& dst) const
    {
        if (empty() || on_device) { dst = *this; return; }
        const int channels = 3;
        dst.create(rows, cols, channels);
        for (int y = 0; y < rows; y++) std::memcpy(dst.data + static_cast<size_t>(y) * dst.step, data + static_cast<size_t>(y) * step, static_cast<size_t>(cols) * channels);
        dst.timestamp = timestamp; dst.format = format;
    }
    VideoFrame operator()(const cvlite::Rect& roi) const  // a view sharing the pixels (cv::UMat::operator())
    {
        VideoFrame v = *this;
        v.data = data + static_cast<size_t>(roi.y) * step + static_cast<size_t>(roi.x) * 3;
        v.cols = roi.width; v.rows = roi.height;
        return v;
    }
    bool has_known_format() const { return format != UNKNOWN; }

private:
    std::shared_ptr<std::vector<uint8_t>> storage;
};
#endif
typedef VideoFrame Frame;

namespace detail
{
// Pixel access for the C-ABI: pointer, row pitch and memory space of a frame.  With cv::UMat frames the pixels are mapped
// to the host (cv::UMat::getMat) for the duration of the call.
#ifdef LVK_COMPAT_USE_OPENCV
struct Pixels
{
    cv::Mat mapped;
    uint8_t* data = nullptr; size_t step = 0; lvkb200_memspace space = LVKB200_MEM_HOST;
    Pixels(const VideoFrame& f, bool write) : mapped(f.getMat(write ? cv::ACCESS_RW : cv::ACCESS_READ)), data(mapped.data), step(mapped.step) {}
};
inline bool on_device(const VideoFrame&) { return false; }
inline void allocate(VideoFrame& f, int rows, int cols) { f.create(rows, cols, CV_8UC3); }
#else
struct Pixels
{
    uint8_t* data; size_t step; lvkb200_memspace space;
    Pixels(const VideoFrame& f, bool) : data(f.data), step(f.step), space(f.on_device ? LVKB200_MEM_DEVICE : LVKB200_MEM_HOST) {}
};
inline bool on_device(const VideoFrame& f) { return f.on_device; }
inline void allocate(VideoFrame& f, int rows, int cols) { f.create(rows, cols); }
#endif
}  // namespace detail

// ---- Utility/Configurable.hpp:27-44 -----------------------------------------------------------------------------------
template <typename T>
class Configurable
{
public:
    explicit Configurable(const T& settings = {}) : m_Settings(settings) {}
    virtual ~Configurable() = default;
    void configure_default() { configure(T{}); }
    virtual void configure(const T& settings) = 0;
    void reconfigure(const std::function<void(T&)>& updater) { T s = m_Settings; updater(s); configure(s); }
    const T& settings() const { return m_Settings; }
protected:
    T m_Settings;
};

// ---- Filters/VideoFilter.hpp:32-61 -------------------------------------------------------------------------------------
class VideoFilter : public Unique<VideoFilter>
{
public:
    explicit VideoFilter(const std::string& filter_name = "Identity Filter") : m_Alias(filter_name) {}
    virtual ~VideoFilter() = default;
    const std::string& alias() const { return m_Alias; }

    // VideoFilter.cpp:46-51 (profile => synchronise the device around filter(), Stopwatch::sync_gpu)
    void apply(VideoFrame&& input, VideoFrame& output, const bool profile = false)
    {
        if (profile) sync_device();
        m_FrameTimer.start();
        filter(std::move(input), output);
        if (profile) sync_device();
        m_FrameTimer.stop();
    }
    // VideoFilter.cpp:55-58: a reference-counted (shallow) copy of the input is moved in
    void apply(const VideoFrame& input, VideoFrame& output, const bool profile = false) { apply(Frame(input), output, profile); }

    // VideoFilter.cpp:62-209: reads frames until the input is exhausted, filters them and hands every non-empty output
    // to `callback`; a TRUE return terminates the stream (:180-206).  `Capture` is cv::VideoCapture in the reference;
    // here anything with `bool read(VideoFrame&)` (format / timestamp set by the reader).  The reference overlaps
    // input, filtering and output with three host threads; filters with a device pipeline override stream_frames().
    template <typename Capture>
    void stream(Capture& input, const std::function<bool(Frame&)>& callback, const bool profile = false)
    {
        stream_frames([&input](Frame& f) { return input.read(f); }, callback, profile);
    }
#ifdef LVK_COMPAT_USE_OPENCV
    // the reference's signature (VideoFilter.hpp:47): frames read from the capture are BGR (VideoFilter.cpp:96-100)
    void stream(cv::VideoCapture& input, const std::function<bool(Frame&)>& callback, const bool profile = false)
    {
        uint64_t index = 0;
        stream_frames([&input, &index](Frame& f) {
            if (!input.read(f)) return false;
            f.timestamp = index++;
            f.format = VideoFrame::BGR;
            return true;
        }, callback, profile);
    }
#endif

    void set_timing_samples(const size_t samples) { m_FrameTimer.set_history_size(samples); }
    const Stopwatch& timings() const { return m_FrameTimer; }

protected:
    virtual void filter(VideoFrame&& input, VideoFrame& output) { output = std::move(input); }  // VideoFilter.cpp:229-233
    virtual void sync_device() {}
    virtual void stream_frames(const std::function<bool(Frame&)>& read, const std::function<bool(Frame&)>& callback,
                               const bool profile)
    {
        Frame input_frame, filtered_frame;
        while (read(input_frame))
        {
            apply(std::move(input_frame), filtered_frame, profile);
            if (filtered_frame.empty()) continue;  // VideoFilter.cpp:137-138
            if (callback(filtered_frame)) return;
        }
    }
private:
    Stopwatch m_FrameTimer;
    const std::string m_Alias;
};
typedef VideoFilter IdentityFilter;

// ---- settings (field names and defaults are the reference's API) -----------------------------------------------------------
struct FeatureDetectorSettings  // Vision/FeatureDetector.hpp:28-37
{
    cvlite::Size detection_resolution = {256, 256};
    cvlite::Size detection_regions = {2, 2};
    bool force_detection = false;
    float max_feature_density = 0.20f;
    float min_feature_density = 0.05f;
    float accumulation_rate = 2.0f;
};

struct FrameTrackerSettings : public FeatureDetectorSettings  // Vision/FrameTracker.hpp:31-44
{
    cvlite::Size motion_resolution = {16, 16};
    bool track_local_motions = true;
    float temporal_smoothing = 1.0f;
    float local_smoothing = 20.0f;
    size_t min_motion_samples = 75;
    float acceptance_threshold = 8.0f;
    float uniformity_threshold = 0.20f;
};

struct PathSmootherSettings  // Vision/PathSmoother.hpp:29-39
{
    size_t predictive_samples = 10;
    cvlite::Size motion_resolution = {2, 2};
    cvlite::Size2f corrective_limits = {0.1f, 0.1f};
    float smoothing_steps = 20.0f;
    float response_rate = 0.04f;
};

struct StabilizationFilterSettings : public FrameTrackerSettings, public PathSmootherSettings  // StabilizationFilter.hpp:28-39
{
    cvlite::Size motion_resolution = {2, 2};
    cvlite::Scalar background_colour = {255, 0, 255};
    bool crop_to_stable_region = false;
    bool stabilize_output = true;
    float min_scene_quality = 0.8f;
    float min_tracking_quality = 0.3f;
};

// ---- Filters/StabilizationFilter.hpp:42-77 ---------------------------------------------------------------------------------
class StabilizationFilter final : public VideoFilter, public Configurable<StabilizationFilterSettings>
{
public:
    explicit StabilizationFilter(const StabilizationFilterSettings& settings = {}, int cuda_device = 0)
        : VideoFilter("Stabilization Filter"), m_Device(cuda_device)
    {
        install_assert_bridge();
        const lvkb200_settings pod = to_pod(settings);
        check(lvkb200_stream_create(m_Device, &pod, &m_Stream), "lvkb200_stream_create");
        m_Settings = settings;
        link_motion_resolution();
    }
    ~StabilizationFilter() override { if (m_Stream) lvkb200_stream_destroy(m_Stream); }
    StabilizationFilter(const StabilizationFilter&) = delete;
    StabilizationFilter& operator=(const StabilizationFilter&) = delete;

    void configure(const StabilizationFilterSettings& settings) override  // StabilizationFilter.cpp:42-65
    {
        const lvkb200_settings pod = to_pod(settings);
        if (check(lvkb200_stream_configure(m_Stream, &pod), "StabilizationFilter::configure"))
        {
            m_Settings = settings;
            link_motion_resolution();
        }
    }
    void restart() { check(lvkb200_stream_restart(m_Stream), "StabilizationFilter::restart"); }
    bool ready() const { return lvkb200_stream_ready(m_Stream) != 0; }
    void reset_context() { check(lvkb200_stream_reset_context(m_Stream), "StabilizationFilter::reset_context"); }
    void draw_trackers() {}      // OBS test-mode overlay (StabilizationFilter.cpp:163-177): not on the hot path
    void draw_motion_mesh() {}   // (:181-188)
    size_t frame_delay() const { return static_cast<size_t>(lvkb200_stream_frame_delay(m_Stream)); }
    cvlite::Rect stable_region() const
    {
        cvlite::Rect r;
        lvkb200_stream_stable_region(m_Stream, m_LastWidth, m_LastHeight, &r.x, &r.y, &r.width, &r.height);
        return r;
    }
    const lvkb200_result& last_result() const { return m_Result; }
    lvkb200_stream* native_handle() const { return m_Stream; }

private:
    void filter(VideoFrame&& input, VideoFrame& output) override  // StabilizationFilter.cpp:69-135
    {
        LVK_ASSERT(input.has_known_format());
        LVK_ASSERT(!input.empty());
        m_LastWidth = input.cols; m_LastHeight = input.rows;
        const int w = input.cols, h = input.rows;
        const uint64_t ts = input.timestamp;
        const lvkb200_format fmt = static_cast<lvkb200_format>(input.format);
        const bool device = detail::on_device(input);
        // the reference moves `input` into its queue and overwrites `output`; OBS passes the same object for both
        const bool alias = (&input == &output);
        VideoFrame* dst = &output;
        if (!alias && (output.empty() || output.cols != w || output.rows != h || detail::on_device(output) != device))
        {
            if (device) { dst = &m_Scratch; if (m_Scratch.cols != w || m_Scratch.rows != h) detail::allocate(m_Scratch, h, w); }
            else detail::allocate(output, h, w);
        }
        lvkb200_status st;
        {
            const detail::Pixels src(input, false);
            if (alias)
                st = lvkb200_stream_submit(m_Stream, src.data, src.step, w, h, fmt, ts, src.space, src.data, src.step, src.space, &m_Result);
            else
            {
                const detail::Pixels out(*dst, true);
                st = lvkb200_stream_submit(m_Stream, src.data, src.step, w, h, fmt, ts, src.space, out.data, out.step, out.space, &m_Result);
            }
            // A device output may still be pending (the library can hold the remap back, lvkb200.h): the reference hands out
            // a complete frame (a UMat access synchronises), so does this mirror.  Callers that want the overlap use stream().
            if (st == LVKB200_OK && m_Result.has_output && device) lvkb200_stream_sync(m_Stream);
        }
        if (!check(st, "StabilizationFilter::filter") || !m_Result.has_output)
        {
            output.release();  // StabilizationFilter.cpp:94,134
            return;
        }
        if (dst != &output) output = *dst;
        output.timestamp = m_Result.out_timestamp;                                   // WarpMesh.cpp:221
        output.format = static_cast<VideoFrame::Format>(m_Result.out_format);         // WarpMesh.cpp:222
    }
    void sync_device() override { lvkb200_stream_sync(m_Stream); }  // Stopwatch::sync_gpu (Timing/Stopwatch.cpp:127-131)

    // VideoFilter::stream on the device pipeline: the upload of frame t+1 (lvkb200_stream_prefetch) and the download of
    // output t-1 (submit_async / wait_output) overlap the processing of frame t; two outputs stay in flight, so output
    // t is delivered after submit t+2.  Same outputs as apply(), frame for frame.
    void stream_frames(const std::function<bool(Frame&)>& read, const std::function<bool(Frame&)>& callback,
                       const bool profile) override
    {
        struct Pending { uint64_t ticket; size_t slot; lvkb200_result res; };
        Frame current, next;
        if (!read(current)) return;
        if (profile || detail::on_device(current))  // profiling synchronises around every frame; device frames need no copies
        {
            Frame out;
            do
            {
                apply(std::move(current), out, profile);
                if (!out.empty() && callback(out)) return;
            } while (read(current));
            return;
        }
        VideoFrame outputs[3];
        std::vector<Pending> pending;
        auto deliver_front = [&]() -> bool {
            const Pending p = pending.front();
            pending.erase(pending.begin());
            check(lvkb200_stream_wait_output(m_Stream, p.ticket), "StabilizationFilter::stream");
            VideoFrame& out = outputs[p.slot];
            out.timestamp = p.res.out_timestamp;
            out.format = static_cast<VideoFrame::Format>(p.res.out_format);
            return callback(out);
        };
        auto drain = [&]() { for (const Pending& p : pending) lvkb200_stream_wait_output(m_Stream, p.ticket); pending.clear(); };
        bool have = true;
        for (size_t i = 0; have; i++)
        {
            LVK_ASSERT(current.has_known_format());
            LVK_ASSERT(!current.empty());
            m_LastWidth = current.cols; m_LastHeight = current.rows;
            const bool have_next = read(next);
            if (have_next && !next.empty() && !detail::on_device(next))
            {
                const detail::Pixels np(next, false);
                check(lvkb200_stream_prefetch_frame(m_Stream, np.data, np.step, next.cols, next.rows,
                                                    static_cast<lvkb200_format>(next.format), LVKB200_MEM_HOST),
                      "StabilizationFilter::stream");
            }
            VideoFrame& out = outputs[i % 3];
            if (out.empty() || out.cols != current.cols || out.rows != current.rows) detail::allocate(out, current.rows, current.cols);
            Pending p{0, i % 3, {}};
            const detail::Pixels cp(current, false), op(out, true);
            const lvkb200_status st = lvkb200_stream_submit_async(
                m_Stream, cp.data, cp.step, current.cols, current.rows, static_cast<lvkb200_format>(current.format),
                current.timestamp, LVKB200_MEM_HOST, op.data, op.step, LVKB200_MEM_HOST, &p.res, &p.ticket);
            if (!check(st, "StabilizationFilter::stream")) { drain(); return; }
            m_Result = p.res;
            if (p.res.has_output) pending.push_back(p);
            while (pending.size() > 2)
                if (deliver_front()) { drain(); return; }
            current = std::move(next);  // keeps the prefetched pointer
            have = have_next;
        }
        while (!pending.empty())
            if (deliver_front()) { drain(); return; }
    }

    void link_motion_resolution()  // StabilizationFilter.cpp:57-58
    {
        static_cast<PathSmootherSettings&>(m_Settings).motion_resolution = m_Settings.motion_resolution;
        static_cast<FrameTrackerSettings&>(m_Settings).motion_resolution = m_Settings.motion_resolution;
    }

    static lvkb200_settings to_pod(const StabilizationFilterSettings& s)
    {
        lvkb200_settings p;
        lvkb200_settings_default(&p);
        p.detection_resolution_width = s.detection_resolution.width; p.detection_resolution_height = s.detection_resolution.height;
        p.detection_regions_width = s.detection_regions.width; p.detection_regions_height = s.detection_regions.height;
        p.force_detection = s.force_detection;
        p.max_feature_density = s.max_feature_density; p.min_feature_density = s.min_feature_density;
        p.accumulation_rate = s.accumulation_rate;
        p.motion_resolution_width = s.motion_resolution.width; p.motion_resolution_height = s.motion_resolution.height;
        p.track_local_motions = s.track_local_motions;
        p.temporal_smoothing = s.temporal_smoothing; p.local_smoothing = s.local_smoothing;
        p.min_motion_samples = s.min_motion_samples;
        p.acceptance_threshold = s.acceptance_threshold; p.uniformity_threshold = s.uniformity_threshold;
        p.predictive_samples = s.predictive_samples;
        p.corrective_limits_width = s.corrective_limits.width; p.corrective_limits_height = s.corrective_limits.height;
        p.smoothing_steps = s.smoothing_steps; p.response_rate = s.response_rate;
        for (int i = 0; i < 4; i++) p.background_colour[i] = s.background_colour[i];
        p.crop_to_stable_region = s.crop_to_stable_region; p.stabilize_output = s.stabilize_output;
        p.min_scene_quality = s.min_scene_quality; p.min_tracking_quality = s.min_tracking_quality;
        return p;
    }

    // FFI status codes become assert_handler calls (SURVEY §8b "Errors")
    static bool check(lvkb200_status st, const char* where)
    {
        if (st == LVKB200_OK) return true;
        context::assert_handler("lvkb200", where, lvkb200_last_error());
        return false;
    }
    static void install_assert_bridge()
    {
        // failed reference preconditions inside the library are reported through status codes + last_error();
        // check() forwards them, so no C callback is needed (and none may throw across the C boundary).
        lvkb200_set_assert_handler(nullptr);
    }

    lvkb200_stream* m_Stream = nullptr;
    lvkb200_result m_Result{};
    VideoFrame m_Scratch;
    int m_Device = 0;
    int m_LastWidth = 0, m_LastHeight = 0;
};

// ---- shared plumbing of the stateless per-frame filters (own lvkb200_stream = device scratch + CUDA stream) -----------
namespace detail
{
class DeviceFilterBase
{
protected:
    explicit DeviceFilterBase(int cuda_device)
    {
        lvkb200_set_assert_handler(nullptr);
        check(lvkb200_stream_create(cuda_device, nullptr, &m_Stream), "lvkb200_stream_create");
    }
    ~DeviceFilterBase() { if (m_Stream) lvkb200_stream_destroy(m_Stream); }
    DeviceFilterBase(const DeviceFilterBase&) = delete;
    DeviceFilterBase& operator=(const DeviceFilterBase&) = delete;
    static bool check(lvkb200_status st, const char* where)
    {
        if (st == LVKB200_OK) return true;
        context::assert_handler("lvkb200", where, lvkb200_last_error());
        return false;
    }
    lvkb200_stream* m_Stream = nullptr;
};
}  // namespace detail

// ---- Filters/DeblockingFilter.hpp:26-59 --------------------------------------------------------------------------------------
struct DeblockingFilterSettings
{
    uint32_t detection_levels = 3;  // Must be greater than 0
    uint32_t block_size = 16;       // Must be greater than 0
    uint32_t filter_size = 5;       // Must be odd
    float filter_scaling = 4;       // Smaller is stronger (1/x)
};

class DeblockingFilter final : public VideoFilter, public Configurable<DeblockingFilterSettings>, private detail::DeviceFilterBase
{
public:
    explicit DeblockingFilter(DeblockingFilterSettings settings = {}, int cuda_device = 0)
        : VideoFilter("Deblocking Filter"), detail::DeviceFilterBase(cuda_device)
    {
        configure(settings);
    }
    void configure(const DeblockingFilterSettings& settings) override  // DeblockingFilter.cpp:36-45
    {
        LVK_ASSERT(settings.block_size > 0);
        LVK_ASSERT(settings.filter_size >= 3);
        LVK_ASSERT(settings.filter_size % 2 == 1);
        LVK_ASSERT(settings.detection_levels > 0);
        LVK_ASSERT(settings.filter_scaling > 1.0f);
        m_Settings = settings;
    }
    void draw_influence(VideoFrame&) const {}  // debug overlay (DeblockingFilter.cpp:122-132): not on the hot path
    cvlite::Rect filter_region() const { return m_FilterRegion; }  // :136-139
    lvkb200_stream* native_handle() const { return m_Stream; }

private:
    void filter(VideoFrame&& input, VideoFrame& output) override  // DeblockingFilter.cpp:48-118: in place, then moved out
    {
        LVK_ASSERT(!input.empty());
        const lvkb200_deblock_settings pod{m_Settings.detection_levels, m_Settings.block_size, m_Settings.filter_size,
                                           m_Settings.filter_scaling};
        const int bs = static_cast<int>(m_Settings.block_size);
        m_FilterRegion.x = m_FilterRegion.y = 0;
        m_FilterRegion.width = input.cols / bs * bs; m_FilterRegion.height = input.rows / bs * bs;
        {
            const detail::Pixels px(input, true);
            check(lvkb200_deblock(m_Stream, &pod, px.data, px.step, input.cols, input.rows,
                                  static_cast<lvkb200_format>(input.format), px.space, px.data, px.step, px.space),
                  "DeblockingFilter::filter");
        }
        if (&input != &output) output = std::move(input);
    }
    void sync_device() override { lvkb200_stream_sync(m_Stream); }
    cvlite::Rect m_FilterRegion;
};

// ---- Filters/ScalingFilter.hpp:27-52 -------------------------------------------------------------------------------------------
struct ScalingFilterSettings
{
    cvlite::Size output_size = {1920, 1080};
    float sharpness = 0.8f;
    bool yuv_input = true;
};

class ScalingFilter final : public VideoFilter, public Configurable<ScalingFilterSettings>, private detail::DeviceFilterBase
{
public:
    explicit ScalingFilter(const ScalingFilterSettings& settings = {}, int cuda_device = 0)
        : VideoFilter("Scaling Filter"), detail::DeviceFilterBase(cuda_device)
    {
        configure(settings);
    }
    explicit ScalingFilter(const cvlite::Size& output_size, const float sharpness = 0.8f)
        : ScalingFilter(make_settings(output_size, sharpness)) {}
    void configure(const ScalingFilterSettings& settings) override  // ScalingFilter.cpp:41-48
    {
        LVK_ASSERT(settings.sharpness >= 0.0f && settings.sharpness <= 1.0f);
        LVK_ASSERT(settings.output_size.width > 0);
        LVK_ASSERT(settings.output_size.height > 0);
        m_Settings = settings;
    }
    lvkb200_stream* native_handle() const { return m_Stream; }

private:
    static ScalingFilterSettings make_settings(const cvlite::Size& size, float sharpness)
    {
        ScalingFilterSettings s;
        s.output_size = size; s.sharpness = sharpness;
        return s;
    }
    void filter(VideoFrame&& input, VideoFrame& output) override  // ScalingFilter.cpp:52-59
    {
        LVK_ASSERT(!input.empty());
        const lvkb200_scaling_settings pod{m_Settings.output_size.width, m_Settings.output_size.height, m_Settings.sharpness,
                                           m_Settings.yuv_input ? 1 : 0};
        // a host output is (re)created at the output size (dst.create, Image.cpp:182); a device output must already have it
        VideoFrame result;
        VideoFrame* dst = &result;
        if (&input != &output && detail::on_device(output) && output.cols == pod.output_width && output.rows == pod.output_height)
            dst = &output;
        else
            detail::allocate(result, pod.output_height, pod.output_width);
        bool ok;
        {
            const detail::Pixels src(input, false), out(*dst, true);
            ok = check(lvkb200_scaling_filter(m_Stream, &pod, src.data, src.step, input.cols, input.rows, src.space, out.data,
                                              out.step, out.space), "ScalingFilter::filter");
        }
        const uint64_t ts = input.timestamp;
        const VideoFrame::Format fmt = input.format;
        if (!ok) { output.release(); return; }
        if (dst != &output) output = std::move(result);
        output.timestamp = ts;  // ScalingFilter.cpp:58
        output.format = fmt;
    }
    void sync_device() override { lvkb200_stream_sync(m_Stream); }
};

// ---- Filters/CompositeFilter.hpp:27-75 -----------------------------------------------------------------------------------------
struct CompositeFilterSettings
{
    std::vector<std::shared_ptr<lvk::VideoFilter>> filter_chain;
    bool save_outputs = false;
};

class CompositeFilter final : public VideoFilter, public Configurable<CompositeFilterSettings>
{
public:
    explicit CompositeFilter(const CompositeFilterSettings& settings = {}) : VideoFilter("Composite Filter") { configure(settings); }
    CompositeFilter(const std::initializer_list<std::shared_ptr<lvk::VideoFilter>>& filter_chain,
                    const CompositeFilterSettings& settings = {})
        : VideoFilter("Composite Filter")
    {
        CompositeFilterSettings s;
        s.filter_chain = filter_chain; s.save_outputs = settings.save_outputs;
        configure(s);
    }
    void configure(const CompositeFilterSettings& settings) override  // CompositeFilter.cpp:46-55
    {
        m_Settings = settings;
        m_FilterOutputs.resize(settings.filter_chain.size());
        m_FilterRunState.resize(settings.filter_chain.size(), true);
        enable_all_filters();
    }
    const std::vector<std::shared_ptr<lvk::VideoFilter>>& filters() const { return m_Settings.filter_chain; }
    std::shared_ptr<lvk::VideoFilter> filters(const size_t index)
    {
        LVK_ASSERT(index < m_Settings.filter_chain.size());
        return m_Settings.filter_chain[index];
    }
    const std::vector<Frame>& outputs() const { return m_FilterOutputs; }
    const VideoFrame& outputs(const size_t index) { LVK_ASSERT(index < m_FilterOutputs.size()); return m_FilterOutputs[index]; }
    bool is_filter_enabled(const size_t index) { LVK_ASSERT(index < m_FilterRunState.size()); return m_FilterRunState[index]; }
    void disable_filter(const size_t index) { LVK_ASSERT(index < m_FilterRunState.size()); m_FilterRunState[index] = false; }
    void enable_filter(const size_t index) { LVK_ASSERT(index < m_FilterRunState.size()); m_FilterRunState[index] = true; }
    void enable_all_filters() { for (size_t i = 0; i < m_FilterRunState.size(); i++) m_FilterRunState[i] = true; }
    size_t filter_count() const { return m_Settings.filter_chain.size(); }

private:
    // CompositeFilter.cpp:58-88: every enabled filter consumes the previous one's output; an empty frame (a filter
    // that is still buffering, e.g. the stabilizer's look-ahead queue) ends the pass; with save_outputs each stage's
    // result is also kept for outputs().
    void filter(VideoFrame&& input, VideoFrame& output) override
    {
        LVK_ASSERT(!input.empty());
        VideoFrame carried = std::move(input);
        const size_t stages = m_Settings.filter_chain.size();
        for (size_t stage = 0; stage < stages && !carried.empty(); stage++)
        {
            if (!m_FilterRunState[stage]) continue;
            VideoFrame& produced = m_FilterOutputs[stage];
            m_Settings.filter_chain[stage]->apply(std::move(carried), produced);
            if (m_Settings.save_outputs)
                carried = produced;             // keep the stage's result, hand a copy on
            else
                carried = std::move(produced);
        }
        output = std::move(carried);
    }
    std::vector<bool> m_FilterRunState;
    std::vector<Frame> m_FilterOutputs;
};

}  // namespace lvk
