// lvkb200_stream definition (opaque to C callers).
#pragma once

#include <cstring>
#include <vector>

#include "common.hpp"
#include "stream.hpp"

struct lvkb200_stream
{
    int device = 0;
    cudaStream_t cs = nullptr;
    lvkb200_settings settings{};
    bool configured = false;

    // scratch used by the stage-level entry points when the caller hands host memory
    lvkb200::DeviceBuffer stage_in, stage_out, mesh_dev;
    lvkb200::PinnedBuffer mesh_pinned;
    cudaEvent_t user_events[LVKB200_EVENT_SLOTS] = {};

    lvkb200_status configure(const lvkb200_settings& s);
    lvkb200_status restart();
    lvkb200_status reset_context();
    bool ready() const;
    void stable_region(int fw, int fh, int* x, int* y, int* w, int* h) const;
    lvkb200_status submit(const void* frame, size_t pitch, int width, int height, lvkb200_format format,
                          uint64_t timestamp, lvkb200_memspace frame_space, void* out, size_t out_pitch,
                          lvkb200_memspace out_space, lvkb200_result* res);
    lvkb200_status debug_fetch(lvkb200_debug_item which, void* buffer, size_t capacity, size_t* size);
    lvkb200_status stage_times(float* times);
    void release();

    // Makes a device view of a caller frame: device memory is used in place, host memory is copied
    // (async, on this stream) into stage_in with a 16-byte aligned pitch.
    lvkb200_status stage_frame_in(const void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                  const uint8_t** dptr, size_t* dpitch);
    // Device destination for a caller frame (caller's own memory if it is device memory).
    lvkb200_status stage_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                   uint8_t** dptr, size_t* dpitch);
    // Copies the staged output back to host memory and waits (host destination only).
    lvkb200_status finish_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space);
    lvkb200_status upload_mesh(const float* offsets, int cols, int rows, const float** dmesh);
};
