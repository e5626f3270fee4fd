// What `#include <LiveVisionKit.hpp>` resolves to when the reference's VideoEditor / OBS-Plugin sources are built against
// this repo: the OpenCV umbrella the reference's header pulls in (a mock opencv2/ stands in for OpenCV in the CPU tests)
// and the lvk-compat header with the reference's own `struct VideoFrame : cv::UMat`.  Test infrastructure;
// oracle/ref_build/build_lvk_editor.sh generates the same redirect (plus the reference's own Logger / CSVLogger, taken in
// place) for the whole-module build, and INTEGRATION.md shows it for a real build.
#pragma once
#define LVK_COMPAT_USE_OPENCV
#include <opencv2/opencv.hpp>
#include <opencv2/core/ocl.hpp>
#include "../../../livevisionkit_b200/compat/lvk/lvk.hpp"
