// MOCK of <opencv2/opencv.hpp> (see core.hpp in this directory).
#pragma once
#include "core.hpp"
#include "highgui.hpp"
#include "videoio.hpp"
