// MOCK of <opencv2/videoio.hpp> (see core.hpp in this directory).  cv::VideoCapture hands out frames a test pushed in,
// or reads the RAW CLIP container below; cv::VideoWriter writes the same container.  That is enough to build and RUN
// the reference's VideoEditor CLI (Modules/VideoEditor, compiled unchanged against lvk-compat) without FFmpeg:
//   "LVKRAW1 <width> <height> <fps> <frames, 10 digits>\n" followed by frames x (height x width x 3) bytes, packed BGR.
// Test infrastructure only.
#pragma once

#include <algorithm>
#include <cstdio>
#include <deque>
#include <fstream>
#include <string>
#include <vector>

#include "core.hpp"

namespace cv
{

enum VideoCaptureAPIs { CAP_ANY = 0, CAP_FFMPEG = 1900 };
enum VideoCaptureProperties
{
    CAP_PROP_POS_FRAMES = 1, CAP_PROP_FRAME_WIDTH = 3, CAP_PROP_FRAME_HEIGHT = 4, CAP_PROP_FPS = 5, CAP_PROP_FOURCC = 6,
    CAP_PROP_FRAME_COUNT = 7, CAP_PROP_HW_ACCELERATION = 50, CAP_PROP_HW_ACCELERATION_USE_OPENCL = 52
};
enum VideoWriterProperties { VIDEOWRITER_PROP_HW_ACCELERATION = 6, VIDEOWRITER_PROP_HW_ACCELERATION_USE_OPENCL = 8 };

class VideoCapture
{
public:
    VideoCapture() = default;
    explicit VideoCapture(const std::string& filename, int = CAP_ANY, const std::vector<int>& = {}) { open(filename); }
    explicit VideoCapture(int /*device index*/) : m_Opened(false), m_Queue(false) {}  // no capture devices in the mock
    virtual ~VideoCapture() = default;

    bool open(const std::string& filename)
    {
        m_Queue = false;
        m_File = std::make_shared<std::ifstream>(filename, std::ios::binary);
        std::string magic;
        long long frames = 0;
        if (!(*m_File >> magic >> m_Width >> m_Height >> m_Fps >> frames) || magic != "LVKRAW1" || m_Width <= 0 || m_Height <= 0)
            return m_Opened = false;
        m_File->get();  // the newline that ends the header
        m_FrameCount = frames;
        m_Position = 0;
        return m_Opened = true;
    }
    virtual bool isOpened() const { return m_Opened; }
    virtual bool read(UMat& image)
    {
        if (m_Queue)
        {
            if (m_Frames.empty()) return false;
            image = m_Frames.front();
            m_Frames.pop_front();
            return true;
        }
        if (!m_Opened || m_Position >= m_FrameCount) return false;
        UMat frame(m_Height, m_Width, CV_8UC3);
        Mat pixels = frame.getMat(ACCESS_WRITE);
        m_File->read(reinterpret_cast<char*>(pixels.data), static_cast<std::streamsize>(pixels.step) * m_Height);
        if (!*m_File) return false;
        m_Position++;
        image = frame;
        return true;
    }
    double get(int property) const
    {
        switch (property)
        {
        case CAP_PROP_POS_FRAMES: return static_cast<double>(m_Position);
        case CAP_PROP_FRAME_WIDTH: return m_Width;
        case CAP_PROP_FRAME_HEIGHT: return m_Height;
        case CAP_PROP_FPS: return m_Fps;
        case CAP_PROP_FRAME_COUNT: return static_cast<double>(m_FrameCount);
        default: return 0.0;
        }
    }
    void push(const UMat& frame) { m_Frames.push_back(frame); }  // test hook: frames handed out by read()

private:
    bool m_Opened = true, m_Queue = true;
    std::deque<UMat> m_Frames;
    std::shared_ptr<std::ifstream> m_File;
    int m_Width = 0, m_Height = 0;
    double m_Fps = 0.0;
    long long m_FrameCount = 0, m_Position = 0;
};

class VideoWriter
{
public:
    VideoWriter() = default;
    VideoWriter(const std::string& filename, int, int, double fps, Size frame_size, const std::vector<int>& = {})
    {
        open(filename, fps, frame_size);
    }
    static int fourcc(char a, char b, char c, char d) { return (a & 255) | ((b & 255) << 8) | ((c & 255) << 16) | ((d & 255) << 24); }
    bool open(const std::string& filename, double fps, Size frame_size)
    {
        m_State = std::make_shared<State>();
        m_State->file.open(filename, std::ios::binary | std::ios::trunc);
        m_State->size = frame_size;
        m_State->fps = fps;
        write_header();
        return isOpened();
    }
    bool isOpened() const { return m_State && m_State->file.good(); }
    void write(const UMat& image)
    {
        if (!isOpened() || image.empty()) return;
        const Mat pixels = image.getMat(ACCESS_READ);
        const size_t row = static_cast<size_t>(image.cols) * 3;
        for (int y = 0; y < image.rows; y++)
            m_State->file.write(reinterpret_cast<const char*>(pixels.data + static_cast<size_t>(y) * pixels.step),
                                static_cast<std::streamsize>(row));
        m_State->frames++;
    }
    void release() { m_State.reset(); }

private:
    struct State
    {
        std::ofstream file;
        Size size;
        double fps = 0.0;
        long long frames = 0;
        ~State()  // the frame count is only known at the end: rewrite the fixed-width header
        {
            if (!file.is_open()) return;
            file.seekp(0);
            char header[96];
            const int n = std::snprintf(header, sizeof(header), "LVKRAW1 %d %d %.6f %010lld\n", size.width, size.height, fps, frames);
            file.write(header, n);
            file.close();
        }
    };
    void write_header()
    {
        char header[96];
        const int n = std::snprintf(header, sizeof(header), "LVKRAW1 %d %d %.6f %010lld\n", m_State->size.width,
                                    m_State->size.height, m_State->fps, 0LL);
        m_State->file.write(header, n);
    }
    std::shared_ptr<State> m_State;
};

}  // namespace cv
