// C-ABI implementation (include/lvkb200.h).  Host C++ only: orchestration + calls into the kernel launchers.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <atomic>
#include <mutex>

#include "common.hpp"
#include "host_math.hpp"
#include "stream_impl.hpp"

namespace lvkb200
{

std::string& last_error()
{
    thread_local std::string err;
    return err;
}

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
}

static lvkb200_assert_handler g_assert_handler = nullptr;
static std::atomic<uint64_t> g_launches{0};
// Launches recorded while THIS thread captures a CUDA graph are not executed: they are tallied per thread and credited on
// every replay of the graph instead (other threads' streams keep counting into the process-wide total meanwhile).
static thread_local bool t_capturing = false;
static thread_local int t_captured = 0;
void count_launches(int n)
{
    if (t_capturing) { t_captured += n; return; }
    g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed);
}
void begin_launch_capture() { t_capturing = true; t_captured = 0; }
int end_launch_capture() { t_capturing = false; return t_captured; }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

void report_assert(const char* file, const char* function, const char* assertion)
{
    const char* base = std::strrchr(file, '/');
    set_error("assertion failed: %s in %s (%s)", assertion, function, base ? base + 1 : file);
    if (g_assert_handler) g_assert_handler(base ? base + 1 : file, function, assertion);
}

}  // namespace lvkb200

using namespace lvkb200;

extern "C" {

int lvkb200_abi_version(void) { return LVKB200_ABI_VERSION; }

int lvkb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* lvkb200_last_error(void) { return last_error().c_str(); }

uint64_t lvkb200_kernel_launch_count(void) { return launch_count(); }

void lvkb200_set_remap_exact(int exact) { set_remap_exact(exact); }
int lvkb200_remap_exact(void) { return remap_exact(); }

const char* lvkb200_status_string(lvkb200_status s)
{
    switch (s)
    {
        case LVKB200_OK: return "ok";
        case LVKB200_ERR_INVALID: return "invalid argument / failed precondition";
        case LVKB200_ERR_CUDA: return "CUDA error";
        case LVKB200_ERR_NO_DEVICE: return "no CUDA device";
        case LVKB200_ERR_NO_MODEL: return "no motion model";
        case LVKB200_ERR_CAPACITY: return "buffer too small";
    }
    return "unknown";
}

void lvkb200_set_assert_handler(lvkb200_assert_handler handler) { g_assert_handler = handler; }

void lvkb200_settings_default(lvkb200_settings* s)
{
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    // FeatureDetectorSettings — Vision/FeatureDetector.hpp:28-37
    s->detection_resolution_width = 256; s->detection_resolution_height = 256;
    s->detection_regions_width = 2; s->detection_regions_height = 2;
    s->force_detection = 0;
    s->max_feature_density = 0.20f;
    s->min_feature_density = 0.05f;
    s->accumulation_rate = 2.0f;
    // FrameTrackerSettings — Vision/FrameTracker.hpp:31-44 (motion_resolution overridden by StabilizationFilterSettings)
    s->motion_resolution_width = 2; s->motion_resolution_height = 2;
    s->track_local_motions = 1;
    s->temporal_smoothing = 1.0f;
    s->local_smoothing = 20.0f;
    s->min_motion_samples = 75;
    s->acceptance_threshold = 8.0f;
    s->uniformity_threshold = 0.20f;
    // PathSmootherSettings — Vision/PathSmoother.hpp:29-39
    s->predictive_samples = 10;
    s->corrective_limits_width = 0.1f; s->corrective_limits_height = 0.1f;
    s->smoothing_steps = 20.0f;
    s->response_rate = 0.04f;
    // StabilizationFilterSettings — Filters/StabilizationFilter.hpp:28-39
    s->background_colour[0] = 255; s->background_colour[1] = 0; s->background_colour[2] = 255; s->background_colour[3] = 0;
    s->crop_to_stable_region = 0;
    s->stabilize_output = 1;
    s->min_scene_quality = 0.8f;
    s->min_tracking_quality = 0.3f;
}

void lvkb200_settings_obs_homography(lvkb200_settings* s)
{
    if (!s) return;
    lvkb200_settings_default(s);
    // Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:269-280
    s->detection_resolution_width = 480; s->detection_resolution_height = 270;
    s->detection_regions_width = 2; s->detection_regions_height = 1;
    s->max_feature_density = 0.12f;
    s->min_feature_density = 0.04f;
    s->accumulation_rate = 3.0f;
    s->track_local_motions = 0;
    s->acceptance_threshold = 3.0f;
    s->motion_resolution_width = 2; s->motion_resolution_height = 2;
}

void lvkb200_settings_obs_field(lvkb200_settings* s)
{
    if (!s) return;
    lvkb200_settings_default(s);
    // Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:257-268 ("Vector Field" subsystem)
    s->detection_resolution_width = 480; s->detection_resolution_height = 270;
    s->acceptance_threshold = 10.0f;
    s->track_local_motions = 1;
    s->motion_resolution_width = 16; s->motion_resolution_height = 16;
    s->detection_regions_width = 2; s->detection_regions_height = 2;
    s->max_feature_density = 0.12f;
    s->min_feature_density = 0.06f;
    s->accumulation_rate = 3.0f;
}

// ---------------------------------------------------------------------------------------------------------------------

lvkb200_status lvkb200_stream_create(int device, const lvkb200_settings* settings, lvkb200_stream** out)
{
    LVKB_REQUIRE(out != nullptr);
    *out = nullptr;
    int n = lvkb200_device_count();
    if (n <= 0 || device < 0 || device >= n)
    {
        set_error("no usable CUDA device (requested %d, found %d) — there is no CPU fallback", device, n);
        return LVKB200_ERR_NO_DEVICE;
    }
    LVKB_CUDA(cudaSetDevice(device));
    std::unique_ptr<lvkb200_stream> s(new lvkb200_stream());
    s->device = device;
    // The tracking chain is a string of small latency-bound kernels on the frame's critical path; the remap of the
    // previous output (own stream, default priority) fills the machine: the chain's CTAs go first whenever slots free.
    int prio_least = 0, prio_greatest = 0;
    LVKB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    LVKB_CUDA(cudaStreamCreateWithPriority(&s->cs, cudaStreamNonBlocking, prio_greatest));
    lvkb200_settings def;
    lvkb200_settings_default(&def);
    LVKB_TRY(s->configure(settings ? *settings : def));
    *out = s.release();
    return LVKB200_OK;
}

void lvkb200_stream_destroy(lvkb200_stream* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    // A remap that is still held back would be launched into the caller's output buffer NOW - which the caller may
    // already have freed (e.g. a garbage-collected tensor): it is discarded, never launched.  A caller that wants the
    // last device output completes it with lvkb200_stream_sync before destroying the stream.
    s->pending.active = false;
    s->sync_all();
    s->release();
    if (s->cs) cudaStreamDestroy(s->cs);
    delete s;
}

lvkb200_status lvkb200_stream_configure(lvkb200_stream* s, const lvkb200_settings* settings)
{
    LVKB_REQUIRE(s != nullptr && settings != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->configure(*settings);
}

lvkb200_status lvkb200_stream_get_settings(const lvkb200_stream* s, lvkb200_settings* out)
{
    LVKB_REQUIRE(s != nullptr && out != nullptr);
    *out = s->settings;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_restart(lvkb200_stream* s)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->restart();
}

lvkb200_status lvkb200_stream_reset_context(lvkb200_stream* s)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->reset_context();
}

int lvkb200_stream_ready(const lvkb200_stream* s) { return s ? (s->ready() ? 1 : 0) : 0; }

uint64_t lvkb200_stream_frame_delay(const lvkb200_stream* s) { return s ? s->settings.predictive_samples : 0; }

lvkb200_status lvkb200_stream_stable_region(const lvkb200_stream* s, int frame_width, int frame_height, int* x, int* y,
                                            int* width, int* height)
{
    LVKB_REQUIRE(s != nullptr && x && y && width && height);
    s->stable_region(frame_width, frame_height, x, y, width, height);
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_sync(lvkb200_stream* s)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->sync_all();
}

lvkb200_status lvkb200_device_synchronize(void)
{
    LVKB_CUDA(cudaDeviceSynchronize());
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_event_record(lvkb200_stream* s, int index)
{
    LVKB_REQUIRE(s != nullptr && index >= 0 && index < LVKB200_EVENT_SLOTS);
    LVKB_CUDA(cudaSetDevice(s->device));
    if (!s->user_events[index]) LVKB_CUDA(cudaEventCreate(&s->user_events[index]));
    LVKB_TRY(s->join_remap(s->cs));  // the event covers the output remaps queued so far (they run on their own stream)
    LVKB_CUDA(cudaEventRecord(s->user_events[index], s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_event_elapsed_ms(lvkb200_stream* s, int start_index, int stop_index, float* ms)
{
    LVKB_REQUIRE(s != nullptr && ms != nullptr);
    LVKB_REQUIRE(start_index >= 0 && start_index < LVKB200_EVENT_SLOTS && stop_index >= 0 &&
                 stop_index < LVKB200_EVENT_SLOTS);
    LVKB_REQUIRE(s->user_events[start_index] != nullptr && s->user_events[stop_index] != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    LVKB_CUDA(cudaEventSynchronize(s->user_events[stop_index]));
    LVKB_CUDA(cudaEventElapsedTime(ms, s->user_events[start_index], s->user_events[stop_index]));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_submit(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                     lvkb200_format format, uint64_t timestamp, lvkb200_memspace frame_space,
                                     void* out, size_t out_pitch, lvkb200_memspace out_space, lvkb200_result* res)
{
    LVKB_REQUIRE(s != nullptr && frame != nullptr && res != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->submit(frame, pitch, width, height, format, timestamp, frame_space, out, out_pitch, out_space, res);
}

lvkb200_status lvkb200_stream_prefetch(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->prefetch(frame, pitch, width, height);
}

lvkb200_status lvkb200_stream_prefetch_frame(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                             lvkb200_format format, lvkb200_memspace frame_space)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_REQUIRE(format == LVKB200_BGR || format == LVKB200_RGB || format == LVKB200_YUV);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->prefetch(frame, pitch, width, height, format, frame_space);
}

lvkb200_status lvkb200_stream_submit_async(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                           lvkb200_format format, uint64_t timestamp, lvkb200_memspace frame_space,
                                           void* out, size_t out_pitch, lvkb200_memspace out_space,
                                           lvkb200_result* res, uint64_t* ticket)
{
    LVKB_REQUIRE(s != nullptr && frame != nullptr && res != nullptr && ticket != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    s->deferred_output = true;
    s->last_ticket = 0;
    const lvkb200_status st = s->submit(frame, pitch, width, height, format, timestamp, frame_space, out, out_pitch,
                                        out_space, res);
    s->deferred_output = false;
    *ticket = (st == LVKB200_OK && res->has_output) ? s->last_ticket : 0;
    return st;
}

lvkb200_status lvkb200_stream_submit_batch(lvkb200_stream* s, const void* const* frames, size_t pitch, int width, int height,
                                           lvkb200_format format, const uint64_t* timestamps, lvkb200_memspace frame_space,
                                           void* const* outs, size_t out_pitch, lvkb200_memspace out_space, int count,
                                           lvkb200_result* results)
{
    LVKB_REQUIRE(s != nullptr && frames != nullptr && outs != nullptr && results != nullptr && count >= 0);
    LVKB_CUDA(cudaSetDevice(s->device));
    const bool announce = format == LVKB200_BGR || format == LVKB200_RGB || format == LVKB200_YUV;
    // host frames and host outputs: the whole sequence runs pipelined (VideoFilter::stream's three threads): upload of
    // frame i+1 and download of output i-1 overlap frame i, two outputs in flight; all outputs have landed on return
    const bool pipelined = frame_space == LVKB200_MEM_HOST && out_space == LVKB200_MEM_HOST && announce;
    uint64_t in_flight[3] = {0, 0, 0};
    int n_flight = 0;
    lvkb200_status st = LVKB200_OK;
    for (int i = 0; i < count && st == LVKB200_OK; i++)
    {
        LVKB_REQUIRE(frames[i] != nullptr);
        // frame i+1 is announced before frame i is submitted: its copy into the ring, detection image and pyramid are
        // queued behind frame i's tracking chain (the input thread of VideoFilter::stream running one frame ahead)
        if (announce && i + 1 < count && frames[i + 1] != nullptr)
            LVKB_TRY(s->prefetch(frames[i + 1], pitch, width, height, format, frame_space));
        s->deferred_output = pipelined;
        s->last_ticket = 0;
        st = s->submit(frames[i], pitch, width, height, format, timestamps ? timestamps[i] : static_cast<uint64_t>(i),
                       frame_space, outs[i], out_pitch, out_space, &results[i]);
        s->deferred_output = false;
        if (st == LVKB200_OK && pipelined && results[i].has_output && s->last_ticket)
        {
            in_flight[n_flight++] = s->last_ticket;
            if (n_flight > 2)
            {
                st = s->wait_output(in_flight[0]);
                in_flight[0] = in_flight[1]; in_flight[1] = in_flight[2];
                n_flight = 2;
            }
        }
    }
    for (int k = 0; k < n_flight; k++)
    {
        const lvkb200_status w = s->wait_output(in_flight[k]);
        if (st == LVKB200_OK) st = w;
    }
    return st;
}

lvkb200_status lvkb200_stream_wait_output(lvkb200_stream* s, uint64_t ticket)
{
    LVKB_REQUIRE(s != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->wait_output(ticket);
}

lvkb200_status lvkb200_stream_debug_fetch(lvkb200_stream* s, lvkb200_debug_item which, void* buffer, size_t capacity,
                                          size_t* size)
{
    LVKB_REQUIRE(s != nullptr && size != nullptr);
    return s->debug_fetch(which, buffer, capacity, size);
}

lvkb200_status lvkb200_stream_stage_totals_us(lvkb200_stream* s, double totals[LVKB200_STAGE_COUNT],
                                              uint64_t counts[LVKB200_STAGE_COUNT], int reset)
{
    LVKB_REQUIRE(s != nullptr && totals != nullptr && counts != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->stage_totals(totals, counts, reset != 0);
}

lvkb200_status lvkb200_stream_stage_times_us(lvkb200_stream* s, float times[LVKB200_STAGE_COUNT])
{
    LVKB_REQUIRE(s != nullptr && times != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    return s->stage_times(times);
}

// ---- stage-level entry points ---------------------------------------------------------------------------------------

lvkb200_status lvkb200_remap_homography(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                        lvkb200_memspace src_space, void* dst, size_t dst_pitch,
                                        lvkb200_memspace dst_space, const double t_inv[9],
                                        const uint8_t background[3], int yuv_input)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr && t_inv != nullptr && background != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0);  // Image.cpp:94
    LVKB_CUDA(cudaSetDevice(s->device));
    RemapParams p{};
    LVKB_TRY(s->stage_frame_in(src, src_pitch, width, height, 3, src_space, &p.src, &p.src_pitch));
    LVKB_TRY(s->stage_frame_out(dst, dst_pitch, width, height, 3, dst_space, &p.dst, &p.dst_pitch));
    p.width = width; p.height = height; p.yuv = yuv_input != 0;
    std::memcpy(p.bg, background, 3);
    float t[9];
    for (int k = 0; k < 9; k++) t[k] = static_cast<float>(t_inv[k]);  // cv::Vec4f(t.at<double>()) — Image.cpp:133-135
    LVKB_CUDA(launch_remap_homography(s->cs, p, t));
    return s->finish_frame_out(dst, dst_pitch, width, height, 3, dst_space);
}

lvkb200_status lvkb200_remap_mesh(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                  lvkb200_memspace src_space, void* dst, size_t dst_pitch, lvkb200_memspace dst_space,
                                  const float* offsets, int mesh_cols, int mesh_rows, const uint8_t background[3],
                                  int yuv_input)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr && offsets != nullptr && background != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0 && mesh_cols >= 2 && mesh_rows >= 2);
    LVKB_CUDA(cudaSetDevice(s->device));
    RemapParams p{};
    LVKB_TRY(s->stage_frame_in(src, src_pitch, width, height, 3, src_space, &p.src, &p.src_pitch));
    LVKB_TRY(s->stage_frame_out(dst, dst_pitch, width, height, 3, dst_space, &p.dst, &p.dst_pitch));
    p.width = width; p.height = height; p.yuv = yuv_input != 0;
    std::memcpy(p.bg, background, 3);
    const float* dmesh = nullptr;
    LVKB_TRY(s->upload_mesh(offsets, mesh_cols, mesh_rows, &dmesh));
    LVKB_CUDA(launch_remap_mesh(s->cs, p, dmesh, mesh_cols, mesh_rows));
    return s->finish_frame_out(dst, dst_pitch, width, height, 3, dst_space);
}

lvkb200_status lvkb200_warp_mesh_apply(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                       lvkb200_memspace src_space, void* dst, size_t dst_pitch,
                                       lvkb200_memspace dst_space, const float* offsets, int mesh_cols, int mesh_rows,
                                       const uint8_t background[3], int yuv_input, double t_inv_out[9])
{
    LVKB_REQUIRE(s != nullptr && offsets != nullptr);
    if (mesh_cols == 2 && mesh_rows == 2)
    {
        double t[9];
        LVKB_REQUIRE(mesh2x2_to_transform(offsets, width, height, t));
        if (t_inv_out) std::memcpy(t_inv_out, t, sizeof(t));
        return lvkb200_remap_homography(s, src, src_pitch, width, height, src_space, dst, dst_pitch, dst_space, t,
                                        background, yuv_input);
    }
    if (t_inv_out) std::memset(t_inv_out, 0, 9 * sizeof(double));
    return lvkb200_remap_mesh(s, src, src_pitch, width, height, src_space, dst, dst_pitch, dst_space, offsets,
                              mesh_cols, mesh_rows, background, yuv_input);
}

// ---- remaining stage-level entry points ------------------------------------------------------------------------------

lvkb200_status lvkb200_stream_set_profiling(lvkb200_stream* s, int enable)
{
    LVKB_REQUIRE(s != nullptr);
    s->profile_stages = enable != 0;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_set_debug_capture(lvkb200_stream* s, int enable)
{
    LVKB_REQUIRE(s != nullptr);
    s->debug_capture = enable != 0;
    return LVKB200_OK;
}

lvkb200_status lvkb200_detection_image(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                       lvkb200_format format, lvkb200_memspace space, uint8_t* det_out, int det_w,
                                       int det_h)
{
    LVKB_REQUIRE(s != nullptr && frame != nullptr && det_out != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0 && det_w > 0 && det_h > 0);
    LVKB_CUDA(cudaSetDevice(s->device));
    const int ch = (format == LVKB200_GRAY) ? 1 : ((format == LVKB200_BGRA || format == LVKB200_RGBA) ? 4 : 3);
    const uint8_t* dsrc = nullptr;
    size_t dpitch = 0;
    LVKB_TRY(s->stage_frame_in(frame, pitch, width, height, ch, space, &dsrc, &dpitch));
    IngestPlan plan;  // stage-level calls use a private plan so the stream's own geometry cache is untouched
    lvkb200_status st = plan.prepare(width, height, det_w, det_h, s->cs);
    DeviceBuffer ddet;
    PinnedBuffer hdet;
    const size_t det_pitch = (static_cast<size_t>(det_w) + 15) / 16 * 16;
    if (st == LVKB200_OK && ddet.ensure(det_pitch * det_h) != cudaSuccess) st = LVKB200_ERR_CUDA;
    if (st == LVKB200_OK && hdet.ensure(static_cast<size_t>(det_w) * det_h) != cudaSuccess) st = LVKB200_ERR_CUDA;
    if (st == LVKB200_OK) st = plan.launch(s->cs, dsrc, dpitch, format, ddet.as<uint8_t>(), det_pitch);
    if (st == LVKB200_OK)
    {
        cudaError_t e = cudaMemcpy2DAsync(hdet.ptr, det_w, ddet.ptr, det_pitch, det_w, det_h, cudaMemcpyDeviceToHost, s->cs);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->cs);
        if (e != cudaSuccess) { set_error("detection_image: %s", cudaGetErrorString(e)); st = LVKB200_ERR_CUDA; }
        else std::memcpy(det_out, hdet.ptr, static_cast<size_t>(det_w) * det_h);
    }
    cudaStreamSynchronize(s->cs);
    plan.release(); ddet.release(); hdet.release();
    return st;
}

// Uploads a host gray image into a temporary device buffer (pitch aligned to 16).
static lvkb200_status upload_gray(lvkb200_stream* s, const uint8_t* img, int w, int h, DeviceBuffer& buf, size_t* pitch)
{
    *pitch = (static_cast<size_t>(w) + 15) / 16 * 16;
    LVKB_CUDA(buf.ensure(*pitch * h));
    LVKB_CUDA(cudaMemcpy2DAsync(buf.ptr, *pitch, img, w, w, h, cudaMemcpyHostToDevice, s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_fast_detect(lvkb200_stream* s, const uint8_t* image, int width, int height, int roi_x,
                                   int roi_y, int roi_w, int roi_h, int threshold, lvkb200_keypoint* keypoints,
                                   int capacity, int* count)
{
    LVKB_REQUIRE(s != nullptr && image != nullptr && count != nullptr);
    LVKB_REQUIRE(capacity == 0 || keypoints != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    DeviceBuffer dimg;
    size_t pitch = 0;
    FastDetector det;
    std::vector<std::vector<FastPoint>> pts;
    lvkb200_status st = upload_gray(s, image, width, height, dimg, &pitch);
    if (st == LVKB200_OK) st = det.prepare(width, height);
    const FastRegion rg{roi_x, roi_y, roi_w, roi_h, threshold};
    if (st == LVKB200_OK) st = det.launch(s->cs, dimg.as<uint8_t>(), pitch, &rg, 1);
    if (st == LVKB200_OK) st = det.fetch(pts);
    cudaStreamSynchronize(s->cs);
    if (st == LVKB200_OK)
    {
        *count = static_cast<int>(pts[0].size());
        const int n = std::min(*count, capacity);
        for (int i = 0; i < n; i++)
            keypoints[i] = {static_cast<float>(pts[0][i].x), static_cast<float>(pts[0][i].y),
                            static_cast<float>(pts[0][i].score), -1};
        if (*count > capacity) st = LVKB200_ERR_CAPACITY;
    }
    det.release(); dimg.release();
    return st;
}

lvkb200_status lvkb200_lk_track(lvkb200_stream* s, const uint8_t* prev, const uint8_t* next, int width, int height,
                                const float* points, int count, int call_index, float* matched, uint8_t* status)
{
    LVKB_REQUIRE(s != nullptr && prev != nullptr && next != nullptr);
    LVKB_REQUIRE(count == 0 || (points != nullptr && matched != nullptr && status != nullptr));
    LVKB_CUDA(cudaSetDevice(s->device));
    if (count == 0) return LVKB200_OK;
    DeviceBuffer dprev, dnext, dp, dq, dst;
    LkPyramid pp, pn;
    size_t pitch = 0;
    lvkb200_status st = upload_gray(s, prev, width, height, dprev, &pitch);
    if (st == LVKB200_OK) st = upload_gray(s, next, width, height, dnext, &pitch);
    if (st == LVKB200_OK) st = pp.prepare(width, height, s->cs);
    if (st == LVKB200_OK) st = pn.prepare(width, height, s->cs);
    if (st == LVKB200_OK) st = pp.build(s->cs, dprev.as<uint8_t>(), pitch);
    if (st == LVKB200_OK) st = pn.build(s->cs, dnext.as<uint8_t>(), pitch);
    auto cuda_ok = [&](cudaError_t e) { if (e != cudaSuccess && st == LVKB200_OK) { set_error("lk_track: %s", cudaGetErrorString(e)); st = LVKB200_ERR_CUDA; } };
    const int padded = (count + 3) / 4 * 4;  // the kernel moves points in groups of four
    if (st == LVKB200_OK)
    {
        cuda_ok(dp.ensure(sizeof(float2) * padded));
        cuda_ok(dq.ensure(sizeof(float2) * padded));
        cuda_ok(dst.ensure(padded));
    }
    if (st == LVKB200_OK) cuda_ok(cudaMemsetAsync(dp.ptr, 0, sizeof(float2) * padded, s->cs));
    if (st == LVKB200_OK) cuda_ok(cudaMemcpyAsync(dp.ptr, points, sizeof(float2) * count, cudaMemcpyHostToDevice, s->cs));
    DeviceBuffer dprm;
    TrackParams hprm{};
    hprm.n = count;
    hprm.lk_epsilon_sq = lk_epsilon_for_call(std::max(call_index, 0));
    if (st == LVKB200_OK) cuda_ok(dprm.ensure(sizeof(TrackParams)));
    if (st == LVKB200_OK) cuda_ok(cudaMemcpyAsync(dprm.ptr, &hprm, sizeof(hprm), cudaMemcpyHostToDevice, s->cs));
    if (st == LVKB200_OK)
    {
        LkIo io{};
        io.pts_in = dp.as<float2>();
        io.prm_in = dprm.as<TrackParams>();
        io.next = dq.as<float2>();
        io.status = dst.as<uint8_t>();
        st = lk_track(s->cs, pp, pn, padded, io);
    }
    if (st == LVKB200_OK) cuda_ok(cudaMemcpyAsync(matched, dq.ptr, sizeof(float2) * count, cudaMemcpyDeviceToHost, s->cs));
    if (st == LVKB200_OK) cuda_ok(cudaMemcpyAsync(status, dst.ptr, count, cudaMemcpyDeviceToHost, s->cs));
    cuda_ok(cudaStreamSynchronize(s->cs));
    pp.release(); pn.release(); dprev.release(); dnext.release(); dp.release(); dq.release(); dst.release();
    dprm.release();
    return st;
}

lvkb200_status lvkb200_find_homography(lvkb200_stream* s, const float* src_points, const float* dst_points, int count,
                                       float threshold, double h_out[9], uint8_t* mask)
{
    LVKB_REQUIRE(s != nullptr && src_points != nullptr && dst_points != nullptr && h_out != nullptr && mask != nullptr);
    LVKB_REQUIRE(count >= 4);  // FrameTracker.cpp:335
    LVKB_CUDA(cudaSetDevice(s->device));
    std::vector<float> a(src_points, src_points + 2 * static_cast<size_t>(count));
    std::vector<float> b(dst_points, dst_points + 2 * static_cast<size_t>(count));
    std::vector<uint8_t> m;
    bool found = false;
    LVKB_TRY(s->run_homography(a, b, threshold, 0, h_out, m, &found));
    std::memcpy(mask, m.data(), count);
    if (!found)
    {
        set_error("find_homography: no model (degenerate correspondences)");
        return LVKB200_ERR_NO_MODEL;
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_estimate_affine_partial(lvkb200_stream* s, const float* src_points, const float* dst_points,
                                               int count, float threshold, double h_out[9], uint8_t* mask)
{
    LVKB_REQUIRE(s != nullptr && src_points != nullptr && dst_points != nullptr && h_out != nullptr && mask != nullptr);
    LVKB_REQUIRE(count >= 4);  // FrameTracker.cpp:335
    LVKB_CUDA(cudaSetDevice(s->device));
    std::vector<float> a(src_points, src_points + 2 * static_cast<size_t>(count));
    std::vector<float> b(dst_points, dst_points + 2 * static_cast<size_t>(count));
    std::vector<uint8_t> m;
    bool found = false;
    LVKB_TRY(s->run_homography(a, b, threshold, 1, h_out, m, &found));
    std::memcpy(mask, m.data(), count);
    if (!found)
    {
        set_error("estimate_affine_partial: no model (degenerate correspondences)");
        return LVKB200_ERR_NO_MODEL;
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_estimate_local_motions(lvkb200_stream* s, const float* tracked, const float* matched, int count,
                                              float* mesh_state, float* offsets_out, uint8_t* mask)
{
    LVKB_REQUIRE(s != nullptr && tracked != nullptr && matched != nullptr && mesh_state != nullptr &&
                 offsets_out != nullptr && mask != nullptr);
    LVKB_CUDA(cudaSetDevice(s->device));
    const size_t elems = static_cast<size_t>(2) * s->settings.motion_resolution_width * s->settings.motion_resolution_height;
    std::vector<float> a(tracked, tracked + 2 * static_cast<size_t>(count));
    std::vector<float> b(matched, matched + 2 * static_cast<size_t>(count));
    Mesh offsets;
    std::vector<uint8_t> m;
    // >= 64 unknowns: k_mesh_cgls on the device; the default 2x2 mesh: host solver (host_mesh.hpp)
    LVKB_TRY(s->run_local_motions(a, b, mesh_state, offsets, m));
    std::memcpy(offsets_out, offsets.data(), sizeof(float) * elems);
    std::memcpy(mask, m.data(), count);
    return LVKB200_OK;
}

// ---- lvk::DeblockingFilter ------------------------------------------------------------------------------------------

void lvkb200_deblock_settings_default(lvkb200_deblock_settings* s)
{
    if (!s) return;
    // DeblockingFilterSettings — Filters/DeblockingFilter.hpp:28-31
    s->detection_levels = 3;
    s->block_size = 16;
    s->filter_size = 5;
    s->filter_scaling = 4.0f;
}

lvkb200_status lvkb200_deblock(lvkb200_stream* s, const lvkb200_deblock_settings* settings, const void* frame,
                               size_t pitch, int width, int height, lvkb200_format format,
                               lvkb200_memspace frame_space, void* out, size_t out_pitch, lvkb200_memspace out_space)
{
    LVKB_REQUIRE(s != nullptr && settings != nullptr && frame != nullptr && out != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0);  // LVK_ASSERT(!input.empty()) — DeblockingFilter.cpp:50
    LVKB_REQUIRE(format == LVKB200_BGR || format == LVKB200_RGB || format == LVKB200_YUV);
    LVKB_CUDA(cudaSetDevice(s->device));
    LVKB_TRY(deblock_validate(*settings));
    const size_t row = static_cast<size_t>(width) * 3;
    LVKB_REQUIRE(pitch >= row && out_pitch >= row);
    const uint8_t* din = nullptr;
    size_t din_pitch = 0;
    LVKB_TRY(s->stage_frame_in(frame, pitch, width, height, 3, frame_space, &din, &din_pitch));
    // the kernels work in place on 4-byte aligned rows: straight in the caller's device buffer when it qualifies,
    // otherwise in the stream's staging buffer
    const bool direct = out_space == LVKB200_MEM_DEVICE && (out_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    uint8_t* work = static_cast<uint8_t*>(out);
    size_t work_pitch = out_pitch;
    if (!direct)
    {
        work_pitch = (row + 15) / 16 * 16;
        LVKB_CUDA(s->stage_out.ensure(work_pitch * height));
        work = s->stage_out.as<uint8_t>();
    }
    if (work != din)
        LVKB_CUDA(cudaMemcpy2DAsync(work, work_pitch, din, din_pitch, row, height, cudaMemcpyDeviceToDevice, s->cs));
    LVKB_TRY(s->deblock_stage.prepare(width, height, *settings, s->cs));
    LVKB_TRY(s->deblock_stage.launch(s->cs, work, work_pitch, format));
    if (!direct)
        LVKB_CUDA(cudaMemcpy2DAsync(out, out_pitch, work, work_pitch, row, height,
                                    out_space == LVKB200_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, s->cs));
    if (out_space == LVKB200_MEM_HOST || frame_space == LVKB200_MEM_HOST) LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_set_deblocking(lvkb200_stream* s, const lvkb200_deblock_settings* settings)
{
    LVKB_REQUIRE(s != nullptr);
    if (!settings)
    {
        s->deblock_enabled = false;
        return LVKB200_OK;
    }
    LVKB_TRY(deblock_validate(*settings));
    s->deblock_settings = *settings;
    s->deblock_enabled = true;
    return LVKB200_OK;
}

// ---- lvk::ScalingFilter ---------------------------------------------------------------------------------------------

void lvkb200_scaling_settings_default(lvkb200_scaling_settings* s)
{
    if (!s) return;
    // ScalingFilterSettings — Filters/ScalingFilter.hpp:29-31
    s->output_width = 1920;
    s->output_height = 1080;
    s->sharpness = 0.8f;
    s->yuv_input = 1;
}

// Image.cpp:227: the kernel argument, computed in float
static float rcas_kernel_sharpness(float sharpness) { return std::exp2(-2.0f * (1.0f - sharpness)); }

// upscale of a device frame into a device frame on s->cs (Image.cpp:155-201)
static lvkb200_status upscale_device(lvkb200_stream* s, const uint8_t* src, size_t src_pitch, int width, int height,
                                     uint8_t* dst, size_t dst_pitch, int dst_width, int dst_height, bool yuv)
{
    if (dst_width == width && dst_height == height)  // src.copyTo(dst) — Image.cpp:162-166
    {
        LVKB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, static_cast<size_t>(width) * 3, height,
                                    cudaMemcpyDeviceToDevice, s->cs));
        return LVKB200_OK;
    }
    RemapParams p{};
    p.src = src; p.src_pitch = src_pitch; p.dst = dst; p.dst_pitch = dst_pitch;
    p.width = width; p.height = height; p.dst_width = dst_width; p.dst_height = dst_height; p.yuv = yuv;
    LVKB_CUDA(launch_upscale(s->cs, p));
    return LVKB200_OK;
}

lvkb200_status lvkb200_upscale(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                               lvkb200_memspace src_space, void* dst, size_t dst_pitch, int dst_width, int dst_height,
                               lvkb200_memspace dst_space, int yuv_input)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0);                          // Image.cpp:158
    LVKB_REQUIRE(dst_width >= width && dst_height >= height);      // Image.cpp:157
    LVKB_CUDA(cudaSetDevice(s->device));
    const uint8_t* din = nullptr;
    size_t din_pitch = 0;
    uint8_t* dout = nullptr;
    size_t dout_pitch = 0;
    LVKB_TRY(s->stage_frame_in(src, src_pitch, width, height, 3, src_space, &din, &din_pitch));
    LVKB_TRY(s->stage_frame_out(dst, dst_pitch, dst_width, dst_height, 3, dst_space, &dout, &dout_pitch));
    LVKB_TRY(upscale_device(s, din, din_pitch, width, height, dout, dout_pitch, dst_width, dst_height, yuv_input != 0));
    LVKB_TRY(s->finish_frame_out(dst, dst_pitch, dst_width, dst_height, 3, dst_space));
    if (src_space == LVKB200_MEM_HOST && dst_space != LVKB200_MEM_HOST) LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_sharpen(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                               lvkb200_memspace src_space, void* dst, size_t dst_pitch, lvkb200_memspace dst_space,
                               float sharpness)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0);                // Image.cpp:207
    LVKB_REQUIRE(sharpness >= 0.0f && sharpness <= 1.0f);  // LVK_ASSERT_01 — Image.cpp:209
    LVKB_CUDA(cudaSetDevice(s->device));
    const uint8_t* din = nullptr;
    size_t din_pitch = 0;
    uint8_t* dout = nullptr;
    size_t dout_pitch = 0;
    LVKB_TRY(s->stage_frame_in(src, src_pitch, width, height, 3, src_space, &din, &din_pitch));
    LVKB_TRY(s->stage_frame_out(dst, dst_pitch, width, height, 3, dst_space, &dout, &dout_pitch));
    if (din == dout)
    {
        // in place on a device frame: the taps must see the unsharpened input -> sharpen a scratch copy of it
        const size_t row = static_cast<size_t>(width) * 3, sp = (row + 15) / 16 * 16;
        LVKB_CUDA(s->scaling_scratch.ensure(sp * height));
        LVKB_CUDA(cudaMemcpy2DAsync(s->scaling_scratch.ptr, sp, din, din_pitch, row, height, cudaMemcpyDeviceToDevice, s->cs));
        din = s->scaling_scratch.as<uint8_t>();
        din_pitch = sp;
    }
    LVKB_CUDA(launch_rcas(s->cs, din, din_pitch, dout, dout_pitch, width, height, rcas_kernel_sharpness(sharpness)));
    LVKB_TRY(s->finish_frame_out(dst, dst_pitch, width, height, 3, dst_space));
    if (src_space == LVKB200_MEM_HOST && dst_space != LVKB200_MEM_HOST) LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_scaling_filter(lvkb200_stream* s, const lvkb200_scaling_settings* settings, const void* frame,
                                      size_t pitch, int width, int height, lvkb200_memspace frame_space, void* out,
                                      size_t out_pitch, lvkb200_memspace out_space)
{
    LVKB_REQUIRE(s != nullptr && settings != nullptr && frame != nullptr && out != nullptr);
    LVKB_REQUIRE(width > 0 && height > 0);  // LVK_ASSERT(!input.empty()) — ScalingFilter.cpp:54
    // ScalingFilter::configure — ScalingFilter.cpp:43-45
    LVKB_REQUIRE(settings->sharpness >= 0.0f && settings->sharpness <= 1.0f);
    LVKB_REQUIRE(settings->output_width > 0);
    LVKB_REQUIRE(settings->output_height > 0);
    const int ow = settings->output_width, oh = settings->output_height;
    LVKB_REQUIRE(ow >= width && oh >= height);  // lvk::upscale — Image.cpp:157
    LVKB_CUDA(cudaSetDevice(s->device));
    const uint8_t* din = nullptr;
    size_t din_pitch = 0;
    uint8_t* dout = nullptr;
    size_t dout_pitch = 0;
    LVKB_TRY(s->stage_frame_in(frame, pitch, width, height, 3, frame_space, &din, &din_pitch));
    LVKB_TRY(s->stage_frame_out(out, out_pitch, ow, oh, 3, out_space, &dout, &dout_pitch));
    // the upscaled frame stays in the stream's device scratch (16-byte aligned rows for the RCAS tile loads)
    const size_t sp = (static_cast<size_t>(ow) * 3 + 15) / 16 * 16;
    LVKB_CUDA(s->scaling_scratch.ensure(sp * oh));
    uint8_t* mid = s->scaling_scratch.as<uint8_t>();
    LVKB_TRY(upscale_device(s, din, din_pitch, width, height, mid, sp, ow, oh, settings->yuv_input != 0));
    LVKB_CUDA(launch_rcas(s->cs, mid, sp, dout, dout_pitch, ow, oh, rcas_kernel_sharpness(settings->sharpness)));
    LVKB_TRY(s->finish_frame_out(out, out_pitch, ow, oh, 3, out_space));
    if (frame_space == LVKB200_MEM_HOST && out_space != LVKB200_MEM_HOST) LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

}  // extern "C"
