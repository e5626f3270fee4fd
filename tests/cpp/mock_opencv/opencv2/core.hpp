// MOCK of the small part of <opencv2/core.hpp> that lvk-compat (-DLVK_COMPAT_USE_OPENCV) and the reference's call sites
// touch: geometry types, cv::Mat / cv::UMat with host storage and reference counting, cv::format.  Test infrastructure
// only (tests/test_compat_cpu.py): it exists so that the boundary can be COMPILED against the reference's real
// signatures in an image without OpenCV; nothing here is used by the product.
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32FC2 13
#define CV_MAT_CN(flags) ((((flags) >> 3) & 511) + 1)

namespace cv
{

// geometry as in OpenCV: class templates with the usual aliases (the library's Functions/Drawing.hpp takes cv::Rect_<T>)
template <typename T> struct Point_
{
    T x = 0, y = 0;
    Point_() = default;
    Point_(T px, T py) : x(px), y(py) {}
    Point_ operator+(const Point_& o) const { return {static_cast<T>(x + o.x), static_cast<T>(y + o.y)}; }
};
template <typename T> struct Size_
{
    T width = 0, height = 0;
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
using Point = Point_<int>;
using Point2f = Point_<float>;
using Size = Size_<int>;
using Size2f = Size_<float>;
struct Scalar { double val[4] = {0, 0, 0, 0}; Scalar() = default;
                Scalar(double a, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
                double operator[](int i) const { return val[i]; } double& operator[](int i) { return val[i]; } };
template <typename T> struct Rect_
{
    T x = 0, y = 0, width = 0, height = 0;
    Rect_() = default;
    Rect_(T px, T py, T w, T h) : x(px), y(py), width(w), height(h) {}
    Point_<T> tl() const { return {x, y}; }
    Point_<T> br() const { return {static_cast<T>(x + width), static_cast<T>(y + height)}; }
    Size_<T> size() const { return {width, height}; }
};
using Rect = Rect_<int>;
using Rect2f = Rect_<float>;

enum AccessFlag { ACCESS_READ = 1 << 24, ACCESS_WRITE = 1 << 25, ACCESS_RW = 3 << 24 };
enum UMatUsageFlags { USAGE_DEFAULT = 0, USAGE_ALLOCATE_HOST_MEMORY = 1 << 0, USAGE_ALLOCATE_DEVICE_MEMORY = 1 << 1,
                      USAGE_ALLOCATE_SHARED_MEMORY = 1 << 2 };

struct Mat
{
    uint8_t* data = nullptr;
    size_t step = 0;
    int rows = 0, cols = 0, flags = 0;
    std::shared_ptr<std::vector<uint8_t>> storage;
    bool empty() const { return data == nullptr || rows <= 0 || cols <= 0; }
};

class UMat
{
public:
    int rows = 0, cols = 0, flags = CV_8UC3;
    size_t step = 0;

    UMat() = default;
    explicit UMat(UMatUsageFlags) {}  // (the mock has one kind of memory)
    UMat(int r, int c, int type) { create(r, c, type); }
    UMat(const UMat&) = default;             // reference-counted: copies share the pixels
    UMat(UMat&& o) noexcept { *this = std::move(o); }
    UMat& operator=(const UMat&) = default;
    UMat& operator=(UMat&& o) noexcept
    {
        if (this != &o)
        {
            rows = o.rows; cols = o.cols; flags = o.flags; step = o.step; m_Storage = std::move(o.m_Storage); m_Offset = o.m_Offset;
            o.release();
        }
        return *this;
    }
    virtual ~UMat() = default;

    void create(int r, int c, int type)
    {
        const size_t pitch = static_cast<size_t>(c) * static_cast<size_t>(CV_MAT_CN(type));
        if (m_Storage && m_Storage.use_count() == 1 && r == rows && c == cols && type == flags && step == pitch) return;
        m_Storage = std::make_shared<std::vector<uint8_t>>(pitch * static_cast<size_t>(r), 0);
        rows = r; cols = c; flags = type; step = pitch; m_Offset = 0;
    }
    void create(Size size, int type) { create(size.height, size.width, type); }
    void release() { m_Storage.reset(); rows = cols = 0; step = 0; m_Offset = 0; }
    bool empty() const { return !m_Storage || rows <= 0 || cols <= 0; }
    Size size() const { return {cols, rows}; }
    int type() const { return flags; }
    int channels() const { return CV_MAT_CN(flags); }
    Mat getMat(int) const
    {
        Mat m;
        m.storage = m_Storage;
        m.data = m_Storage ? m_Storage->data() + m_Offset : nullptr;
        m.step = step; m.rows = rows; m.cols = cols; m.flags = flags;
        return m;
    }
    UMat clone() const
    {
        UMat c;
        copyTo(c);
        return c;
    }
    void copyTo(UMat& dst) const
    {
        if (empty()) { dst.release(); return; }
        dst.create(rows, cols, flags);
        const size_t row = static_cast<size_t>(cols) * static_cast<size_t>(channels());
        for (int y = 0; y < rows; y++)
            std::memcpy(dst.m_Storage->data() + dst.m_Offset + static_cast<size_t>(y) * dst.step,
                        m_Storage->data() + m_Offset + static_cast<size_t>(y) * step, row);
    }
    UMat operator()(const Rect& roi) const
    {
        UMat v = *this;
        v.m_Offset = m_Offset + static_cast<size_t>(roi.y) * step + static_cast<size_t>(roi.x) * static_cast<size_t>(channels());
        v.rows = roi.height; v.cols = roi.width;
        return v;
    }

private:
    std::shared_ptr<std::vector<uint8_t>> m_Storage;
    size_t m_Offset = 0;
};

inline std::string format(const char* fmt, ...)
{
    char buffer[512];
    va_list args;
    va_start(args, fmt);
    std::vsnprintf(buffer, sizeof(buffer), fmt, args);
    va_end(args);
    return buffer;
}

}  // namespace cv
