// Internal helpers shared by the C-ABI implementation and the kernel launchers.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/lvkb200.h"

namespace lvkb200
{

// Thread-local last-error text behind lvkb200_last_error().
std::string& last_error();
void set_error(const char* fmt, ...);

// Reports a failed reference precondition through the installed assert handler
// (lvk::context::assert_handler equivalent, Directives.hpp:37-44) and records it.
void report_assert(const char* file, const char* function, const char* assertion);

#define LVKB_CUDA(call)                                                                                               \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t err__ = (call);                                                                                   \
        if (err__ != cudaSuccess)                                                                                     \
        {                                                                                                             \
            ::lvkb200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__));            \
            return LVKB200_ERR_CUDA;                                                                                  \
        }                                                                                                             \
    } while (0)

#define LVKB_REQUIRE(cond)                                                                                            \
    do                                                                                                                \
    {                                                                                                                 \
        if (!(cond))                                                                                                  \
        {                                                                                                             \
            ::lvkb200::report_assert(__FILE__, __func__, #cond);                                                      \
            return LVKB200_ERR_INVALID;                                                                               \
        }                                                                                                             \
    } while (0)

#define LVKB_TRY(expr)                                                                                                \
    do                                                                                                                \
    {                                                                                                                 \
        lvkb200_status st__ = (expr);                                                                                 \
        if (st__ != LVKB200_OK) return st__;                                                                          \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Per-frame scalars of the tracking chain (LK -> swap-erase compaction -> RANSAC).  They live in DEVICE memory and
// are refreshed by one small H2D copy, so the chain's kernels have frame-independent arguments and the whole chain
// can be replayed as one CUDA graph.
struct TrackParams
{
    int n;                  // points handed to the optical flow this frame
    int model;              // 0 = homography (cv::findHomography), 1 = 4-dof similarity (cv::estimateAffinePartial2D)
    double lk_epsilon_sq;   // stopping epsilon of this calc() call (see lk_epsilon_for_call)
    float threshold_sq;     // acceptance threshold^2 of the motion estimator
    float reserved2;
};

// Process-wide count of kernels launched by this library (bench.py reports it as gpu_launches).
void count_launches(int n);
uint64_t launch_count();
// Around a stream capture on the calling thread: launches are tallied (returned by end_) instead of counted.
void begin_launch_capture();
int end_launch_capture();

// ---- kernel launchers (one per reference stage) --------------------------------------------------------------------

struct RemapParams
{
    const uint8_t* src;  // device, packed 8UC3
    size_t src_pitch;
    uint8_t* dst;  // device, packed 8UC3
    size_t dst_pitch;
    int width, height;  // source size (== destination size for the remaps)
    uint8_t bg[3];
    bool yuv;
    int dst_width = 0, dst_height = 0;  // launch_upscale only
};

// easu_remap_homography (FSR.cl:407-452). t = dst->src transform narrowed to float (Image.cpp:133-135).
cudaError_t launch_remap_homography(cudaStream_t cs, const RemapParams& p, const float t[9]);
// easu_remap (FSR.cl:362-403) with the WarpMesh::apply upsample (WarpMesh.cpp:190-191) fused in:
// mesh = device pointer to rows*cols float2 normalized offsets.
cudaError_t launch_remap_mesh(cudaStream_t cs, const RemapParams& p, const float* mesh, int mesh_cols, int mesh_rows);

// lvk::upscale — easu_scale (FSR.cl:326-358, Image.cpp:155-201): p.width x p.height -> p.dst_width x p.dst_height.
cudaError_t launch_upscale(cudaStream_t cs, const RemapParams& p);
// Arithmetic build of the EASU launchers above: 0 = contract (remap_fast.cu, default), 1 = exact (remap.cu).
void set_remap_exact(int exact);
int remap_exact();
// lvk::sharpen — rcas (FSR.cl:460-535, Image.cpp:205-233), out of place; kernel_sharpness = exp2(-2 (1 - sharpness)).
cudaError_t launch_rcas(cudaStream_t cs, const uint8_t* src, size_t src_pitch, uint8_t* dst, size_t dst_pitch, int width,
                        int height, float kernel_sharpness);

}  // namespace lvkb200
