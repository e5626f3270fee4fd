// lvkb200_stream: per-frame orchestration of the stabilization path (host C++), state and scratch.
#include "stream_impl.hpp"

#include <algorithm>
#include <cmath>

#include "host_math.hpp"

using namespace lvkb200;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

lvkb200_status lvkb200_stream::configure(const lvkb200_settings& s)
{
    // Preconditions of StabilizationFilter::configure (StabilizationFilter.cpp:44-45), FrameTracker::configure
    // (FrameTracker.cpp:59-65), FeatureDetector::configure (FeatureDetector.cpp:50-57), PathSmoother::configure
    // (PathSmoother.cpp:38-44).
    LVKB_REQUIRE(s.min_tracking_quality >= 0.0f && s.min_tracking_quality <= 1.0f);
    LVKB_REQUIRE(s.min_scene_quality >= 0.0f && s.min_scene_quality <= 1.0f);
    LVKB_REQUIRE(s.motion_resolution_width >= 2 && s.motion_resolution_height >= 2);
    LVKB_REQUIRE(s.acceptance_threshold >= 0.0f);
    LVKB_REQUIRE(s.temporal_smoothing >= 0.0f);
    LVKB_REQUIRE(s.local_smoothing >= 0.0f);
    LVKB_REQUIRE(s.min_motion_samples >= 4);
    LVKB_REQUIRE(s.uniformity_threshold >= 0.0f && s.uniformity_threshold <= 1.0f);
    LVKB_REQUIRE(s.detection_regions_width > 0 && s.detection_regions_height > 0);
    LVKB_REQUIRE(s.detection_regions_height <= s.detection_resolution_height);
    LVKB_REQUIRE(s.detection_regions_width <= s.detection_resolution_width);
    LVKB_REQUIRE(s.min_feature_density <= s.max_feature_density);
    LVKB_REQUIRE(s.min_feature_density > 0.0f);
    LVKB_REQUIRE(s.accumulation_rate > 0.0f);
    LVKB_REQUIRE(s.max_feature_density >= 0.0f && s.max_feature_density <= 1.0f);
    LVKB_REQUIRE(s.min_feature_density >= 0.0f && s.min_feature_density <= 1.0f);
    LVKB_REQUIRE(s.corrective_limits_width >= 0.0f && s.corrective_limits_width <= 1.0f);
    LVKB_REQUIRE(s.corrective_limits_height >= 0.0f && s.corrective_limits_height <= 1.0f);
    LVKB_REQUIRE(s.predictive_samples > 0);
    LVKB_REQUIRE(s.smoothing_steps > 0.0f);
    LVKB_REQUIRE(s.response_rate >= 0.0f && s.response_rate <= 1.0f);
    settings = s;
    configured = true;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::restart() { return LVKB200_OK; }
lvkb200_status lvkb200_stream::reset_context() { return LVKB200_OK; }
bool lvkb200_stream::ready() const { return false; }

void lvkb200_stream::stable_region(int fw, int fh, int* x, int* y, int* w, int* h) const
{
    // StabilizationFilter::stable_region (StabilizationFilter.cpp:199-): scene margins scaled to the frame.
    const float thc = 1.0f * settings.corrective_limits_width, tvc = 1.0f * settings.corrective_limits_height;
    const float mx = thc / 2, my = tvc / 2, mw = 1.0f - thc, mh = 1.0f - tvc;
    *x = static_cast<int>(std::lrintf(mx * static_cast<float>(fw)));
    *y = static_cast<int>(std::lrintf(my * static_cast<float>(fh)));
    *w = static_cast<int>(std::lrintf(mw * static_cast<float>(fw)));
    *h = static_cast<int>(std::lrintf(mh * static_cast<float>(fh)));
}

lvkb200_status lvkb200_stream::submit(const void*, size_t, int, int, lvkb200_format, uint64_t, lvkb200_memspace, void*,
                                      size_t, lvkb200_memspace, lvkb200_result*)
{
    set_error("submit: not implemented yet");
    return LVKB200_ERR_INVALID;
}

lvkb200_status lvkb200_stream::debug_fetch(lvkb200_debug_item, void*, size_t, size_t* size)
{
    *size = 0;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::stage_times(float* times)
{
    for (int i = 0; i < LVKB200_STAGE_COUNT; i++) times[i] = 0.0f;
    return LVKB200_OK;
}

void lvkb200_stream::release()
{
    stage_in.release();
    stage_out.release();
    mesh_dev.release();
    mesh_pinned.release();
    for (auto& e : user_events)
    {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
}

lvkb200_status lvkb200_stream::stage_frame_in(const void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                              const uint8_t** dptr, size_t* dpitch)
{
    const size_t row = static_cast<size_t>(w) * ch;
    LVKB_REQUIRE(pitch >= row);
    if (space == LVKB200_MEM_DEVICE)
    {
        *dptr = static_cast<const uint8_t*>(p);
        *dpitch = pitch;
        return LVKB200_OK;
    }
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(stage_in.ensure(dp * h));
    LVKB_CUDA(cudaMemcpy2DAsync(stage_in.ptr, dp, p, pitch, row, h, cudaMemcpyHostToDevice, cs));
    *dptr = stage_in.as<uint8_t>();
    *dpitch = dp;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::stage_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                               uint8_t** dptr, size_t* dpitch)
{
    const size_t row = static_cast<size_t>(w) * ch;
    LVKB_REQUIRE(pitch >= row);
    if (space == LVKB200_MEM_DEVICE)
    {
        *dptr = static_cast<uint8_t*>(p);
        *dpitch = pitch;
        return LVKB200_OK;
    }
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(stage_out.ensure(dp * h));
    *dptr = stage_out.as<uint8_t>();
    *dpitch = dp;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::finish_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space)
{
    if (space == LVKB200_MEM_DEVICE) return LVKB200_OK;
    const size_t row = static_cast<size_t>(w) * ch;
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(cudaMemcpy2DAsync(p, pitch, stage_out.ptr, dp, row, h, cudaMemcpyDeviceToHost, cs));
    LVKB_CUDA(cudaStreamSynchronize(cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::upload_mesh(const float* offsets, int cols, int rows, const float** dmesh)
{
    const size_t bytes = sizeof(float) * 2 * static_cast<size_t>(cols) * rows;
    LVKB_CUDA(mesh_dev.ensure(bytes));
    LVKB_CUDA(mesh_pinned.ensure(bytes));
    // the pinned staging copy may still be in flight from a previous call on this stream
    LVKB_CUDA(cudaStreamSynchronize(cs));
    std::memcpy(mesh_pinned.ptr, offsets, bytes);
    LVKB_CUDA(cudaMemcpyAsync(mesh_dev.ptr, mesh_pinned.ptr, bytes, cudaMemcpyHostToDevice, cs));
    *dmesh = mesh_dev.as<float>();
    return LVKB200_OK;
}
