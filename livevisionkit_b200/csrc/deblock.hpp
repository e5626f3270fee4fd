// DeblockingFilter launcher state: per-geometry interpolation tables and the small work images.
#pragma once

#include "common.hpp"
#include "stream.hpp"

namespace lvkb200
{

struct DeblockPlan
{
    // geometry + settings the tables were built for
    int w = 0, h = 0;
    lvkb200_deblock_settings settings{};
    int bs = 0, sc = 0;      // macroblock size, integer down-scaling factor
    int rw = 0, rh = 0;      // m_FilterRegion: the largest whole-macroblock region (DeblockingFilter.cpp:67-68)
    int ex = 0, ey = 0;      // macroblock_extent
    int sw = 0, sh = 0;      // size of the down-scaled (smooth) image
    size_t small_pitch = 0;  // bytes, multiple of 4
    DeviceBuffer d_small, d_median, d_keep, d_x8, d_y8, d_xf, d_yf, d_levels;

    lvkb200_status prepare(int width, int height, const lvkb200_deblock_settings& s, cudaStream_t cs);
    // Filters `frame` (packed 8UC3, device memory) IN PLACE, stream-ordered on `cs`.
    lvkb200_status launch(cudaStream_t cs, uint8_t* frame, size_t pitch, lvkb200_format format);
    void release();
};

// DeblockingFilter::configure preconditions (DeblockingFilter.cpp:38-42) + what this build supports.
lvkb200_status deblock_validate(const lvkb200_deblock_settings& s);

}  // namespace lvkb200
