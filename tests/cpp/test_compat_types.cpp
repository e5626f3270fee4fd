// CPU-only behaviour of the lvk-compat value types (no stream is created, so no GPU is needed):
// lvk::Time / lvk::Stopwatch (Timing/Stopwatch.cpp:27-166), lvk::Unique (Utility/Unique.tpp), lvk::VideoFrame copy
// semantics (reference-counted shallow copies like cv::UMat, Data/VideoFrame.cpp:37-44; clone() is deep).
#include <cstdio>
#include <string>

#include "../../livevisionkit_b200/compat/lvk/lvk.hpp"

#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main()
{
    using lvk::Time;
    CHECK(Time::Milliseconds(2.5).microseconds() == 2500.0);
    CHECK(Time::Seconds(0.5).frequency() == 2.0);
    CHECK((Time::Milliseconds(3) - Time::Milliseconds(1)).milliseconds() == 2.0);
    CHECK((Time::Milliseconds(3) / 2.0).microseconds() == 1500.0);
    CHECK(Time(0).is_zero() && Time::Hours(1).hms() == "01:00:00" && Time::Timestep(50.0).milliseconds() == 20.0);
    CHECK(Time::Nanoseconds(5) < Time::Nanoseconds(6) && Time::Nanoseconds(7) >= Time::Nanoseconds(7));

    lvk::Stopwatch watch(3);
    CHECK(watch.average().is_zero() && watch.deviation().is_zero() && !watch.is_running() && !watch.is_paused());
    for (int i = 0; i < 5; i++)
    {
        watch.start();
        CHECK(watch.is_running());
        watch.wait_until(Time::Microseconds(200.0 * (i + 1)));
        const Time t = watch.stop();
        CHECK(t >= Time::Microseconds(200.0 * (i + 1)) && watch.elapsed() == t && !watch.is_running());
    }
    CHECK(watch.history().size() == 3 && watch.history().capacity() == 3 && watch.history().is_full());
    CHECK(watch.history().newest() >= Time::Microseconds(1000.0) && watch.history().oldest() >= Time::Microseconds(600.0));
    const Time avg = watch.average();
    CHECK(avg >= Time::Microseconds(800.0));
    double mad = 0;  // mean absolute deviation, Stopwatch.cpp:142-160
    for (const Time& t : watch.history()) mad += std::abs(t.nanoseconds() - avg.nanoseconds());
    CHECK(std::abs(watch.deviation().nanoseconds() - mad / 3.0) <= 2.0);
    watch.start();
    const Time p = watch.pause();
    CHECK(watch.is_paused() && watch.pause() == p);
    watch.set_history_size(1);
    CHECK(watch.history().size() == 1);
    watch.reset_history();
    CHECK(watch.history().is_empty());

    lvk::Unique<> a, b, c(a);
    lvk::Unique<> d(std::move(b));
    CHECK(a.uid() != c.uid() && d.uid() == b.uid() && a.uid() + 1 == b.uid());

    lvk::VideoFrame frame;
    frame.create(4, 6);
    frame.timestamp = 42; frame.format = lvk::VideoFrame::YUV;
    frame.data[7] = 9;
    lvk::VideoFrame shallow = frame;                 // shares the pixels (cv::UMat reference counting)
    lvk::VideoFrame deep = frame.clone();
    frame.data[7] = 11;
    CHECK(shallow.data == frame.data && shallow.data[7] == 11 && shallow.timestamp == 42 && shallow.format == lvk::VideoFrame::YUV);
    CHECK(deep.data != frame.data && deep.data[7] == 9 && deep.cols == 6 && deep.rows == 4);
    shallow.create(4, 6);                            // shared pixels are not reused by create()
    CHECK(shallow.data != frame.data);
    lvk::VideoFrame view = frame(lvk::cvlite::Rect(1, 1, 2, 2));
    CHECK(view.data == frame.data + frame.step + 3 && view.cols == 2 && view.step == frame.step);
    lvk::VideoFrame moved = std::move(frame);
    CHECK(frame.empty() && !moved.empty() && moved.data[7] == 11);
    // lvk::TickTimer (Timing/TickTimer.cpp:30-66): a Stopwatch that counts its laps; tick(timestep) stretches the lap
    lvk::TickTimer ticker(4);
    CHECK(ticker.tick_count() == 0 && ticker.delta().is_zero());
    ticker.start();
    const Time lap = ticker.tick(Time::Microseconds(300.0));
    CHECK(lap >= Time::Microseconds(300.0) && ticker.delta() == lap && ticker.tick_count() == 1 && ticker.is_running());
    ticker.tick();
    CHECK(ticker.tick_count() == 2 && ticker.history().size() == 2 && ticker.average() >= ticker.delta() / 2.0);
    ticker.reset_counter();
    CHECK(ticker.tick_count() == 0 && ticker.history().size() == 2);

    // the assertion family (Directives.hpp:46-95): a failed check reports the violated relation and carries on
    std::string reported;
    const auto previous_handler = lvk::context::assert_handler;
    lvk::context::assert_handler = [&](std::string, std::string, std::string assertion) { reported = assertion; };
    const float ratio = 1.5f;
    const int level = 7;
    LVK_ASSERT_01(ratio);
    CHECK(reported == "0 <= ratio <= 1");
    LVK_ASSERT_01_STRICT(ratio);
    CHECK(reported == "0 < ratio < 1");
    LVK_ASSERT_RANGE(level, 1, 5);
    CHECK(reported == "1 <= level <= 5");
    LVK_ASSERT_RANGE_STRICT(level, 1, 7);
    CHECK(reported == "1 < level < 7");
    reported.clear();
    LVK_ASSERT_IF(level > 10, ratio < 1.0f);  // the condition does not hold: nothing is checked
    LVK_ASSERT_RANGE(level, 1, 7);
    LVK_ASSERT_01(0.25f);
    CHECK(reported.empty());
    LVK_ASSERT_IF(level > 5, ratio < 1.0f);
    CHECK(reported == "ratio < 1.0f");
    lvk::context::assert_handler = previous_handler;

    std::printf("compat types ok\n");
    return 0;
}
