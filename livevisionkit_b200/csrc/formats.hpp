// Frame ingest / egress: OBS planar, semi-planar and packed YUV layouts <-> the packed 8UC3 frames the filters work on.
// Replaces Modules/OBS-Plugin/Interop/FrameIngest.cpp:479-723 (I4XXIngest, NV12Ingest, P422Ingest, P444Ingest,
// DirectIngest): the reference uploads the raw planes and then runs cv::resize / cv::merge / cv::mixChannels on the
// OpenCL device; here ONE kernel per direction reads the planes and writes the packed frame (or the reverse).
#pragma once

#include <vector>

#include "common.hpp"
#include "stream.hpp"

namespace lvkb200
{

// One tap pair of cv::resize(INTER_LINEAR) on 8-bit data (imgproc/src/resize.cpp, resizeGeneric_ + HResizeLinear /
// VResizeLinear<uchar>): source index and the two 11-bit fixed-point weights.
struct LinearTap
{
    int ofs;
    short w0, w1;
};

// Tables for upsampling `src` samples to `dst` samples along one axis.  Horizontal tables clamp the weights at the
// borders (fx = 0), vertical ones keep the weights and clip the row index at fetch time -- as OpenCV does.
std::vector<LinearTap> linear_taps(int src, int dst, bool horizontal);

struct PlaneRef
{
    uint8_t* base;  // first sample of the component
    size_t pitch;   // bytes between rows
    int xstride;    // bytes between horizontally adjacent samples of this component
};

struct FormatPlan
{
    int width = 0, height = 0;
    int chroma_w = 0, chroma_h = 0;
    DeviceBuffer xtab, ytab;  // LinearTap tables (chroma -> frame), only when the chroma planes are subsampled
    cudaError_t prepare(int w, int h, int cw, int ch, cudaStream_t cs);
};

// planes -> packed 8UC3 {Y, U, V} (FrameIngest::to_ocl).  Chroma planes of cw x ch samples are upsampled with
// cv::resize(INTER_LINEAR) arithmetic when (cw, ch) != (w, h).
cudaError_t launch_planes_to_packed(cudaStream_t cs, const FormatPlan& plan, PlaneRef y, PlaneRef u, PlaneRef v,
                                    uint8_t* dst, size_t dst_pitch);
// packed 8UC3 {Y, U, V} -> planes (FrameIngest::to_obs).  sub_x / sub_y in {1, 2}: chroma subsampling factors;
// 2x2 -> cv::resize(INTER_AREA): (a+b+c+d+2)>>2 for single-channel chroma planes, round-half-even of the mean when
// the chroma plane is the interleaved 2-channel one (NV12: OpenCV's vector path does not cover 2 channels);
// 2x1 -> saturate_cast<uchar>((a+b)*0.5f) (round half even).  alpha != nullptr: also writes 255 there (AYUV).
cudaError_t launch_packed_to_planes(cudaStream_t cs, const uint8_t* src, size_t src_pitch, int w, int h, int sub_x,
                                    int sub_y, PlaneRef y, PlaneRef u, PlaneRef v, const PlaneRef* alpha,
                                    bool interleaved_chroma);

}  // namespace lvkb200
