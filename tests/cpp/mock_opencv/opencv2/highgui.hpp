// MOCK of <opencv2/highgui.hpp> (see core.hpp in this directory): there is no display; the calls exist so that the
// reference's VideoEditor sources compile unchanged.  Test infrastructure only.
#pragma once

#include <string>

#include "core.hpp"

namespace cv
{
enum WindowFlags { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1, WINDOW_KEEPRATIO = 0 };
inline void namedWindow(const std::string&, int = WINDOW_AUTOSIZE) {}
inline void imshow(const std::string&, const UMat&) {}
inline int pollKey() { return -1; }
inline int waitKey(int = 0) { return -1; }
inline void destroyAllWindows() {}
}  // namespace cv
