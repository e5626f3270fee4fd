// MOCK of cv::VideoCapture (see core.hpp in this directory): read() hands out frames a test pushed in.
#pragma once

#include <deque>

#include "core.hpp"

namespace cv
{
class VideoCapture
{
public:
    virtual ~VideoCapture() = default;
    virtual bool isOpened() const { return true; }
    virtual bool read(UMat& image)
    {
        if (m_Frames.empty()) return false;
        image = m_Frames.front();
        m_Frames.pop_front();
        return true;
    }
    void push(const UMat& frame) { m_Frames.push_back(frame); }
private:
    std::deque<UMat> m_Frames;
};
}  // namespace cv
