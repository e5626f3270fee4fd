// What `#include <LiveVisionKit.hpp>` resolves to when the reference's VideoEditor / OBS-Plugin sources are built against
// this repo: the lvk-compat header with the reference's own `struct VideoFrame : cv::UMat` (a mock opencv2/ stands in
// for OpenCV in the CPU tests).  Test infrastructure; INTEGRATION.md shows the same one-line redirect for a real build.
#pragma once
#define LVK_COMPAT_USE_OPENCV
#include "../../../livevisionkit_b200/compat/lvk/lvk.hpp"
