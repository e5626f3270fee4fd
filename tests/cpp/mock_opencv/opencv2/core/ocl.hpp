// MOCK of <opencv2/core/ocl.hpp> (see ../core.hpp): no OpenCL runtime behind the mock.
#pragma once
namespace cv { namespace ocl {
inline bool haveOpenCL() { return false; }
inline bool useOpenCL() { return false; }
inline void setUseOpenCL(bool) {}
inline void finish() {}
} }
