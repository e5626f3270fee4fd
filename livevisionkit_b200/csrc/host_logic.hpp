// Host-side, order-dependent logic of the stabilization path (no pixels are touched here):
//   FeatureGrid   — FeatureDetector's suppression grid / region bookkeeping  (Vision/FeatureDetector.cpp:48-214,
//                   Data/SpatialMap.tpp:241-265,588-625, Math/VirtualGrid.cpp:85-91,180-203)
//   PathSmoother  — Vision/PathSmoother.cpp:36-145 over a StreamBuffer ring (Data/StreamBuffer.tpp:37-252)
//   mesh helpers  — WarpMesh::set_to(H) / crop_in / clamp (Math/WarpMesh.cpp:333-342,379-427)
// float/double types follow the reference expression by expression: these decide control flow and feature order,
// which the parity contract requires to be exact.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/lvkb200.h"
#include "fast.hpp"

namespace lvkb200
{

// cv::saturate_cast<int>(float) == cvRound (round half to even); used by Size2f->Size and Rect2f->Rect.
inline int cv_round(float v) { return static_cast<int>(std::lrintf(v)); }

struct VGrid
{
    int cols = 1, rows = 1;
    float ax = 0, ay = 0, aw = 1, ah = 1, kw = 1, kh = 1;
    void set(int c, int r, float x, float y, float w, float h)
    {
        cols = c; rows = r; ax = x; ay = y; aw = w; ah = h;
        kw = w / static_cast<float>(c);
        kh = h / static_cast<float>(r);
    }
    bool test_point(float x, float y) const { return ax <= x && x < ax + aw && ay <= y && y < ay + ah; }  // Rect2f::contains
    void key_of(float x, float y, size_t& kx, size_t& ky) const
    {
        kx = static_cast<size_t>((x - ax) / kw);
        ky = static_cast<size_t>((y - ay) / kh);
    }
};

struct Feature
{
    float x, y;
    float response;
    int class_id;
};

// lvk::step (Functions/Math.tpp:133-142)
template <typename V>
inline V step_to(V current, V target, V amount)
{
    if (current > target) return std::max<V>(current - amount, target);
    return std::min<V>(current + amount, target);
}

constexpr int FAST_MIN_THRESHOLD = 10, FAST_MAX_THRESHOLD = 250, FAST_THRESHOLD_STEP = 5, FAST_FEATURE_TOLERANCE = 150;

class FeatureGrid
{
public:
    struct Region
    {
        float bx, by, bw, bh;  // cv::Rect2f bounds
        int threshold;
        size_t load;
    };

    void configure(const lvkb200_settings& s)  // FeatureDetector::configure
    {
        const float W = static_cast<float>(s.detection_resolution_width), H = static_cast<float>(s.detection_resolution_height);
        // cv::Size2f(resolution) * max_feature_density -> cv::Size (rounded)
        const int gc = cv_round(W * s.max_feature_density), gr = cv_round(H * s.max_feature_density);
        const bool grid_changed = gc != grid.cols || gr != grid.rows || cells.empty();
        grid.set(gc, gr, 0.f, 0.f, W, H);
        if (grid_changed)
        {
            cells.assign(static_cast<size_t>(gc) * gr, EMPTY);
            cell_keys.clear();
        }
        const bool regions_changed = s.detection_regions_width != region_grid.cols || s.detection_regions_height != region_grid.rows;
        region_grid.set(s.detection_regions_width, s.detection_regions_height, 0.f, 0.f, W, H);
        // construct_detection_regions() always rebuilds (threshold = FAST_MIN_THRESHOLD, load = 0)
        (void)regions_changed;
        regions.clear();
        for (int r = 0; r < region_grid.rows; r++)
            for (int c = 0; c < region_grid.cols; c++)
                regions.push_back({static_cast<float>(c) * region_grid.kw, static_cast<float>(r) * region_grid.kh,
                                   region_grid.kw, region_grid.kh, FAST_MIN_THRESHOLD, 0});
        const size_t max_features = static_cast<size_t>(gc) * gr;
        const float max_regions = static_cast<float>(regions.size());
        const float max_region_features = static_cast<float>(max_features) / max_regions;
        const float density_ratio = s.min_feature_density / s.max_feature_density;
        min_feature_load = static_cast<size_t>(max_region_features * density_ratio);
        fast_feature_target = static_cast<size_t>(s.accumulation_rate * max_region_features);
        force_detection = s.force_detection != 0;
        features.reserve(max_features);
    }

    size_t max_feature_capacity() const { return static_cast<size_t>(grid.cols) * grid.rows; }

    void reset()  // FeatureDetector::reset
    {
        clear_grid();
        for (auto& r : regions) r.load = 0;
    }

    // First half of detect(): which regions run FAST this frame, with which threshold.
    void plan_detection(std::vector<FastRegion>& out, std::vector<int>& region_index) const
    {
        out.clear();
        region_index.clear();
        for (size_t i = 0; i < regions.size(); i++)
        {
            const Region& r = regions[i];
            if (force_detection || r.load <= min_feature_load)
            {
                // frame(bounds): cv::Rect2f -> cv::Rect rounds every member
                out.push_back({cv_round(r.bx), cv_round(r.by), cv_round(r.bw), cv_round(r.bh), r.threshold});
                region_index.push_back(static_cast<int>(i));
            }
        }
    }

    // Second half of detect() (FeatureDetector.cpp:120-177): feed the FAST results region by region, in region order.
    // Returns the distribution quality; `out` receives [propagated..., new...].
    float finish_detection(const std::vector<int>& region_index, const std::vector<std::vector<FastPoint>>& fast,
                           std::vector<Feature>& out, std::vector<int>& fast_counts)
    {
        fast_counts.assign(regions.size(), -1);
        size_t k = 0;
        for (size_t i = 0; i < regions.size(); i++)
        {
            Region& rg = regions[i];
            if (k < region_index.size() && region_index[k] == static_cast<int>(i))
            {
                const auto& pts = fast[k++];
                for (const FastPoint& p : pts)
                {
                    Feature f{static_cast<float>(p.x) + rg.bx, static_cast<float>(p.y) + rg.by,
                              static_cast<float>(p.score), 0};
                    size_t kx, ky;
                    grid.key_of(f.x, f.y, kx, ky);
                    size_t& link = cells[ky * grid.cols + kx];
                    if (link == EMPTY)
                    {
                        link = features.size();
                        cell_keys.push_back(static_cast<uint32_t>(ky * grid.cols + kx));
                        features.push_back(f);
                    }
                    else
                    {
                        Feature& mx = features[link];
                        if (f.response > mx.response && mx.class_id <= 0) mx = f;
                    }
                }
                fast_counts[i] = static_cast<int>(pts.size());
                if (pts.size() > fast_feature_target + FAST_FEATURE_TOLERANCE)
                    rg.threshold = step_to<int>(rg.threshold, FAST_MAX_THRESHOLD, FAST_THRESHOLD_STEP);
                else if (pts.size() < fast_feature_target - FAST_FEATURE_TOLERANCE)  // size_t arithmetic, may wrap
                    rg.threshold = step_to<int>(rg.threshold, FAST_MIN_THRESHOLD, FAST_THRESHOLD_STEP);
            }
            rg.load = 0;
        }
        out.swap(features);
        features.clear();
        const float quality = distribution_quality();
        clear_grid();
        return quality;
    }

    void propagate(const std::vector<Feature>& feats)  // FeatureDetector::propagate
    {
        for (const Feature& f : feats)
        {
            if (!grid.test_point(f.x, f.y)) continue;
            size_t kx, ky;
            grid.key_of(f.x, f.y, kx, ky);
            size_t& link = cells[ky * grid.cols + kx];
            if (link == EMPTY)
            {
                link = features.size();
                cell_keys.push_back(static_cast<uint32_t>(ky * grid.cols + kx));
                size_t rx, ry;
                region_grid.key_of(f.x, f.y, rx, ry);
                regions[ry * region_grid.cols + rx].load++;
                features.push_back(f);
            }
            else
            {
                Feature& mx = features[link];
                if (f.response > mx.response && f.class_id >= mx.class_id) mx = f;
            }
        }
    }

    size_t pending() const { return features.size(); }
    std::vector<Region> regions;
    VGrid grid, region_grid;

private:
    static constexpr size_t EMPTY = static_cast<size_t>(-1);

    void clear_grid()
    {
        for (uint32_t k : cell_keys) cells[k] = EMPTY;
        cell_keys.clear();
        features.clear();
    }

    float distribution_quality() const  // SpatialMap::distribution_quality over the cell KEYS in insertion order
    {
        const size_t n = cell_keys.size();
        if (n == 0) return 1.0f;
        constexpr int sectors = 4;
        if (grid.cols <= sectors || grid.rows <= sectors)
            return static_cast<float>(n) / static_cast<float>(cells.size());
        VGrid sg;
        sg.set(sectors, sectors, 0.f, 0.f, static_cast<float>(grid.cols), static_cast<float>(grid.rows));
        size_t buckets[sectors * sectors] = {};
        const size_t ideal = static_cast<size_t>(static_cast<float>(n) / static_cast<float>(sectors * sectors));
        float excess = 0.0f;
        for (uint32_t key : cell_keys)
        {
            const float kx = static_cast<float>(key % grid.cols), ky = static_cast<float>(key / grid.cols);
            if (sg.test_point(kx, ky))
            {
                size_t sx, sy;
                sg.key_of(kx, ky, sx, sy);
                if (++buckets[sy * sectors + sx] > ideal) excess += 1.0f;
            }
        }
        return 1.0f - (excess / static_cast<float>(n - ideal));
    }

    std::vector<size_t> cells;        // SpatialMap<size_t> m_Map
    std::vector<uint32_t> cell_keys;  // m_Data keys in insertion order
    std::vector<Feature> features;    // m_Features
    size_t min_feature_load = 0, fast_feature_target = 0;
    bool force_detection = false;
};

// ---- WarpMesh helpers on row-major [r][c][xy] float arrays ----------------------------------------------------------

using Mesh = std::vector<float>;

// cv::perspectiveTransform on one float point with a double matrix (Homography::transform, Math/Homography.cpp:125-130)
inline void homography_transform_f(const double H[9], float x, float y, float& ox, float& oy)
{
    const double dx = x, dy = y;
    double w = dx * H[6] + dy * H[7] + H[8];
    if (std::fabs(w) > 2.220446049250313e-16)
    {
        w = 1.0 / w;
        ox = static_cast<float>((dx * H[0] + dy * H[1] + H[2]) * w);
        oy = static_cast<float>((dx * H[3] + dy * H[4] + H[5]) * w);
    }
    else
        ox = oy = 0.f;
}

// WarpMesh::set_to(const Homography&, const cv::Size2f& motion_scale) — Math/WarpMesh.cpp:333-342
inline void mesh_set_to_homography(const double H[9], float scale_w, float scale_h, int cols, int rows, Mesh& out)
{
    out.resize(static_cast<size_t>(cols) * rows * 2);
    const float csx = scale_w / static_cast<float>(cols - 1), csy = scale_h / static_cast<float>(rows - 1);
    const float nfx = 1.0f / scale_w, nfy = 1.0f / scale_h;
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++)
        {
            const float px = static_cast<float>(c) * csx, py = static_cast<float>(r) * csy;
            float tx, ty;
            homography_transform_f(H, px, py, tx, ty);
            out[(static_cast<size_t>(r) * cols + c) * 2] = (px - tx) * nfx;
            out[(static_cast<size_t>(r) * cols + c) * 2 + 1] = (py - ty) * nfy;
        }
}

// WarpMesh::crop_in on an identity mesh — Math/WarpMesh.cpp:379-390
inline void mesh_crop_in(Mesh& m, int cols, int rows, float rx, float ry, float rw, float rh)
{
    const float csx = (rw - 1.0f) / static_cast<float>(cols - 1), csy = (rh - 1.0f) / static_cast<float>(rows - 1);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++)
        {
            m[(static_cast<size_t>(r) * cols + c) * 2] += static_cast<float>(c) * csx + rx;
            m[(static_cast<size_t>(r) * cols + c) * 2 + 1] += static_cast<float>(r) * csy + ry;
        }
}

// cv::getGaussianKernel(n, sigma, CV_32F): float32(w / sum w), w_i = exp(-(i - (n-1)/2)^2 / (2 sigma^2)) in double
inline void gaussian_kernel_f32(int n, double sigma, std::vector<float>& out)
{
    std::vector<double> w(n);
    const double sigmaX = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2X = -0.5 / (sigmaX * sigmaX);
    double sum = 0;
    for (int i = 0; i < n; i++)
    {
        const double x = i - (n - 1) * 0.5;
        w[i] = std::exp(scale2X * x * x);
        sum += w[i];
    }
    sum = 1.0 / sum;
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = static_cast<float>(w[i] * sum);
}

class PathSmoother
{
public:
    void configure(const lvkb200_settings& s)  // PathSmoother::configure
    {
        const int cols = s.motion_resolution_width, rows = s.motion_resolution_height;
        const size_t elems = static_cast<size_t>(cols) * rows * 2;
        if (cols != mcols || rows != mrows)
        {
            mcols = cols; mrows = rows;
            // m_Trajectory.fill(resolution): clear + pad_back identity up to the CURRENT capacity
            for (auto& m : traj) m.assign(elems, 0.f);
            trace.assign(elems, 0.f);
            position.assign(elems, 0.f);
        }
        const size_t window = 2 * static_cast<size_t>(s.predictive_samples) + 1;
        if (traj.size() != window)
        {
            // resize() keeps the newest elements; pad_front() fills the front with identity meshes
            std::vector<Mesh> keep;
            if (traj.size() > window) keep.assign(traj.end() - window, traj.end());
            else keep = traj;
            traj.assign(window - keep.size(), Mesh(elems, 0.f));
            traj.insert(traj.end(), keep.begin(), keep.end());
            const size_t centre = (window - 1) / 2;
            position = traj[0];
            for (size_t i = 1; i <= centre; i++)
                for (size_t k = 0; k < elems; k++) position[k] += traj[i][k];
            base_smoothing = static_cast<double>(window) / 12.0;
        }
        // crop<float>({1,1}, corrective_limits) — Functions/Math.tpp:218-233
        const float thc = 1.0f * s.corrective_limits_width, tvc = 1.0f * s.corrective_limits_height;
        margin_x = thc / 2; margin_y = tvc / 2; margin_w = 1.0f - thc; margin_h = 1.0f - tvc;
        scene_crop.assign(elems, 0.f);
        mesh_crop_in(scene_crop, cols, rows, margin_x, margin_y, margin_w, margin_h);
        smoothing_steps = s.smoothing_steps;
        response_rate = s.response_rate;
    }

    void restart()  // PathSmoother::restart
    {
        for (auto& m : traj) std::fill(m.begin(), m.end(), 0.f);
        std::fill(position.begin(), position.end(), 0.f);
        std::fill(trace.begin(), trace.end(), 0.f);
    }

    // PathSmoother::next — returns the path correction
    void next(const Mesh& motion, Mesh& correction)
    {
        const size_t elems = motion.size();
        for (size_t k = 0; k < elems; k++) position[k] -= traj[0][k];
        // ring push: the oldest slot is recycled
        std::rotate(traj.begin(), traj.begin() + 1, traj.end());
        traj.back() = motion;
        const size_t centre = (traj.size() - 1) / 2;
        for (size_t k = 0; k < elems; k++) position[k] += traj[centre][k];

        gaussian_kernel_f32(static_cast<int>(traj.size()), base_smoothing + smoothing_factor, filter);

        float weight = 1.0f;
        trace = traj[0];
        for (size_t i = 1; i < traj.size(); i++)
        {
            weight -= filter[i - 1];
            for (size_t k = 0; k < elems; k++) trace[k] = traj[i][k] * weight + trace[k];  // cv::scaleAdd
        }
        correction.resize(elems);
        for (size_t k = 0; k < elems; k++) correction[k] = trace[k] - position[k];

        float max_drift = 0.0f;
        for (size_t k = 0; k < elems; k += 2)
        {
            const float xd = std::fabs(correction[k]) / margin_x, yd = std::fabs(correction[k + 1]) / margin_y;
            max_drift = std::max(max_drift, xd);
            max_drift = std::max(max_drift, yd);
        }
        if (max_drift > 1.0f)
        {
            for (size_t k = 0; k < elems; k += 2)
            {
                correction[k] = std::min(std::max(correction[k], -margin_x), margin_x);
                correction[k + 1] = std::min(std::max(correction[k + 1], -margin_y), margin_y);
            }
            max_drift = 1.0f;
        }
        // hysteresis<double>(drift, 0.3, smoothing_steps, 0.7, 0.0) then exp_moving_average<double>(.., float rate)
        const double d = max_drift;
        double target = d;
        if (d >= 0.7) target = 0.0;
        else if (d <= 0.3) target = static_cast<double>(smoothing_steps);
        smoothing_factor = smoothing_factor + response_rate * (target - smoothing_factor);
        last_drift = max_drift;
    }

    Mesh scene_crop;
    float margin_x = 0.05f, margin_y = 0.05f, margin_w = 0.9f, margin_h = 0.9f;
    double smoothing_factor = 0.0;
    float last_drift = 0.f;

private:
    int mcols = 0, mrows = 0;
    std::vector<Mesh> traj;
    Mesh trace, position;
    std::vector<float> filter;
    double base_smoothing = 0.0;
    float smoothing_steps = 20.f, response_rate = 0.04f;
};

}  // namespace lvkb200
