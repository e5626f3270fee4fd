/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Host side for the reference's FSR.cl compiled on the CPU (oracle/ref_build/opencl_c_shim.hpp): launches the
 * reference's __kernel functions over the same NDRange and with the same arguments as the reference's host code does,
 *   LiveVisionKit/Functions/Image.cpp:28-81   lvk::remap(src, dst, offset_map, background)   -> easu_remap
 *   LiveVisionKit/Functions/Image.cpp:85-151  lvk::remap(src, dst, homography, background, inverted) -> easu_remap_homography
 *   LiveVisionKit/Functions/Image.cpp:155-201 lvk::upscale                                   -> easu_scale
 *   LiveVisionKit/Functions/Image.cpp:205-233 lvk::sharpen                                   -> rcas
 *   LiveVisionKit/Functions/OpenCL/Kernels.cpp:49-72 optimal_groups (8x8 groups, global size rounded up)
 * This translation unit is appended to the kernel text by build_ref.sh: the kernels live in namespaces fsr_bgr
 * (program built without flags) and fsr_yuv (program built with -D YUV_INPUT), Image.cpp:37-38.
 * The exported symbols mirror oracle/easu_ref.c's so tests can swap one for the other.
 */
#include <thread>
#include <vector>

namespace
{
template <class F> void parallel_group_rows(int group_rows, int threads, F fn)
{
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads > group_rows) threads = group_rows;
    if (threads <= 1) { fn(0, group_rows); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        const int g0 = (int)((long long)group_rows * t / threads), g1 = (int)((long long)group_rows * (t + 1) / threads);
        pool.emplace_back([=] { fn(g0, g1); });
    }
    for (auto& th : pool) th.join();
}
inline int groups(int n) { return (n + 7) / 8; }
} // namespace

extern "C" {

/* t = the dst->src homography (row-major doubles), narrowed to float by cv::Vec4f as Image.cpp:133-135 does. */
void ref_easu_remap_homography(const uint8_t* src, int src_step, int rows, int cols, uint8_t* dst, int dst_step,
                               const double t[9], const uint8_t bg[3], int yuv, int threads)
{
    const float4 r1((float)t[0], (float)t[1], (float)t[2], 0.0f), r2((float)t[3], (float)t[4], (float)t[5], 0.0f),
        r3((float)t[6], (float)t[7], (float)t[8], 0.0f);
    const uchar4 background(bg[0], bg[1], bg[2], 0);
    const int4 bounds(0, 0, cols, rows); /* dst is not an ROI: offset (0,0), Image.cpp:118-119,132 */
    parallel_group_rows(groups(rows), threads, [&](int g0, int g1) {
        cl_run_groups(g0, g1, groups(cols), [&] {
            if (yuv) fsr_yuv::easu_remap_homography((uchar*)src, src_step, 0, rows, cols, dst, dst_step, 0, bounds, r1, r2, r3, background);
            else fsr_bgr::easu_remap_homography((uchar*)src, src_step, 0, rows, cols, dst, dst_step, 0, bounds, r1, r2, r3, background);
        });
    });
}

/* map = CV_32FC2 offsets; dst takes the map's size (Image.cpp:52). */
void ref_easu_remap_map(const uint8_t* src, int src_step, int src_rows, int src_cols, uint8_t* dst, int dst_step,
                        const float* map, int map_step, int map_rows, int map_cols, const uint8_t bg[3], int yuv, int threads)
{
    const uchar4 background(bg[0], bg[1], bg[2], 0);
    const int4 bounds(0, 0, map_cols, map_rows);
    parallel_group_rows(groups(map_rows), threads, [&](int g0, int g1) {
        cl_run_groups(g0, g1, groups(map_cols), [&] {
            if (yuv) fsr_yuv::easu_remap((uchar*)src, src_step, 0, src_rows, src_cols, dst, dst_step, 0, bounds, (uchar*)map, map_step, 0, background);
            else fsr_bgr::easu_remap((uchar*)src, src_step, 0, src_rows, src_cols, dst, dst_step, 0, bounds, (uchar*)map, map_step, 0, background);
        });
    });
}

void ref_easu_scale(const uint8_t* src, int src_step, int src_rows, int src_cols, uint8_t* dst, int dst_step, int dst_rows,
                    int dst_cols, int yuv, int threads)
{
    if (dst_rows == src_rows && dst_cols == src_cols) /* Image.cpp:162-166 */
    {
        for (int y = 0; y < src_rows; y++) std::memcpy(dst + (size_t)y * dst_step, src + (size_t)y * src_step, 3 * (size_t)src_cols);
        return;
    }
    const float2 rscale((float)src_cols / (float)dst_cols, (float)src_rows / (float)dst_rows); /* Image.cpp:191-194 */
    parallel_group_rows(groups(dst_rows), threads, [&](int g0, int g1) {
        cl_run_groups(g0, g1, groups(dst_cols), [&] {
            if (yuv) fsr_yuv::easu_scale((uchar*)src, src_step, 0, src_rows, src_cols, dst, dst_step, 0, dst_rows, dst_cols, rscale);
            else fsr_bgr::easu_scale((uchar*)src, src_step, 0, src_rows, src_cols, dst, dst_step, 0, dst_rows, dst_cols, rscale);
        });
    });
}

/* kernel_sharpness = exp2(-2 (1 - s)) (Image.cpp:227).  The kernel's border test (FSR.cl:481) lets the NDRange's
 * padding items copy pixels that lie outside the image; the reference's buffers absorb that, so the kernel runs here on
 * padded copies (8 extra columns and rows) and the image region is copied back.  src and dst are distinct (the in-place
 * call of ScalingFilter.cpp:57 is a data race in the reference and has no defined result). */
void ref_rcas(const uint8_t* src, int src_step, int rows, int cols, uint8_t* dst, int dst_step, float kernel_sharpness, int threads)
{
    const int pstep = 3 * (cols + 8) + 16, prows = rows + 9;
    std::vector<uint8_t> ps((size_t)pstep * prows, 0), pd((size_t)pstep * prows, 0);
    for (int y = 0; y < rows; y++) std::memcpy(ps.data() + (size_t)y * pstep, src + (size_t)y * src_step, 3 * (size_t)cols);
    parallel_group_rows(groups(rows), threads, [&](int g0, int g1) {
        cl_run_groups(g0, g1, groups(cols), [&] {
            fsr_bgr::rcas(ps.data(), pstep, 0, rows, cols, pd.data(), pstep, 0, kernel_sharpness);
        });
    });
    for (int y = 0; y < rows; y++) std::memcpy(dst + (size_t)y * dst_step, pd.data() + (size_t)y * pstep, 3 * (size_t)cols);
}

float ref_rcas_kernel_sharpness(float sharpness) { return std::exp2(-2.0f * (1.0f - sharpness)); } /* Image.cpp:227 */

} // extern "C"
