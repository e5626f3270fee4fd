// TEST-ONLY shim: exposes the product's host-side logic headers (livevisionkit_b200/csrc/host_*.hpp) to Python so
// that tests/test_hostlogic_cpu.py can drive them without a GPU and compare them with the oracle.  It is compiled by
// the test itself (g++), never shipped, never loaded by the product.
#include <cstring>
#include <vector>

#include "../livevisionkit_b200/csrc/host_logic.hpp"
#include "../livevisionkit_b200/csrc/host_math.hpp"
#include "../livevisionkit_b200/csrc/host_mesh.hpp"

using namespace lvkb200;

struct Shim
{
    lvkb200_settings settings;
    FeatureGrid grid;
    PathSmoother smoother;
    MeshSolver solver;
    std::vector<Feature> features;
    std::vector<FastRegion> regions;
    std::vector<int> region_index;
    std::vector<int> fast_counts;
};

extern "C" {

void* shim_create(const lvkb200_settings* s)
{
    Shim* h = new Shim();
    h->settings = *s;
    h->grid.configure(*s);
    h->smoother.configure(*s);
    h->solver.configure(*s);
    return h;
}

void shim_destroy(void* p) { delete static_cast<Shim*>(p); }

void shim_grid_info(void* p, int* cols, int* rows, int* nregions)
{
    Shim* h = static_cast<Shim*>(p);
    *cols = h->grid.grid.cols;
    *rows = h->grid.grid.rows;
    *nregions = static_cast<int>(h->grid.regions.size());
}

// -> number of regions that want FAST this frame; rects/thresholds written as 5 ints each.
int shim_plan(void* p, int* out)
{
    Shim* h = static_cast<Shim*>(p);
    h->grid.plan_detection(h->regions, h->region_index);
    for (size_t i = 0; i < h->regions.size(); i++)
    {
        out[6 * i + 0] = h->region_index[i];
        out[6 * i + 1] = h->regions[i].x; out[6 * i + 2] = h->regions[i].y;
        out[6 * i + 3] = h->regions[i].w; out[6 * i + 4] = h->regions[i].h;
        out[6 * i + 5] = h->regions[i].threshold;
    }
    return static_cast<int>(h->regions.size());
}

// pts: concatenated (x, y, score) int triplets per planned region, counts[i] entries each.
float shim_finish(void* p, const int* pts, const int* counts, int* n_features)
{
    Shim* h = static_cast<Shim*>(p);
    std::vector<std::vector<FastPoint>> fast(h->regions.size());
    size_t off = 0;
    for (size_t i = 0; i < h->regions.size(); i++)
    {
        fast[i].resize(counts[i]);
        for (int k = 0; k < counts[i]; k++, off++)
            fast[i][k] = {static_cast<short>(pts[3 * off]), static_cast<short>(pts[3 * off + 1]), pts[3 * off + 2]};
    }
    const float q = h->grid.finish_detection(h->region_index, fast, h->features, h->fast_counts);
    *n_features = static_cast<int>(h->features.size());
    return q;
}

void shim_get_features(void* p, float* out)  // x, y, response, class_id
{
    Shim* h = static_cast<Shim*>(p);
    for (size_t i = 0; i < h->features.size(); i++)
    {
        out[4 * i] = h->features[i].x; out[4 * i + 1] = h->features[i].y;
        out[4 * i + 2] = h->features[i].response; out[4 * i + 3] = static_cast<float>(h->features[i].class_id);
    }
}

void shim_set_features(void* p, const float* in, int n)
{
    Shim* h = static_cast<Shim*>(p);
    h->features.resize(n);
    for (int i = 0; i < n; i++)
        h->features[i] = {in[4 * i], in[4 * i + 1], in[4 * i + 2], static_cast<int>(in[4 * i + 3])};
}

void shim_propagate(void* p) { Shim* h = static_cast<Shim*>(p); h->grid.propagate(h->features); }

void shim_region_state(void* p, int* thresholds, int* loads)
{
    Shim* h = static_cast<Shim*>(p);
    for (size_t i = 0; i < h->grid.regions.size(); i++)
    {
        thresholds[i] = h->grid.regions[i].threshold;
        loads[i] = static_cast<int>(h->grid.regions[i].load);
    }
}

void shim_smoother_next(void* p, const float* motion, int elems, float* correction, double* smoothing_factor, float* drift)
{
    Shim* h = static_cast<Shim*>(p);
    Mesh m(motion, motion + elems), c;
    h->smoother.next(m, c);
    std::memcpy(correction, c.data(), sizeof(float) * elems);
    *smoothing_factor = h->smoother.smoothing_factor;
    *drift = h->smoother.last_drift;
}

void shim_scene_crop(void* p, float* out, int elems)
{
    Shim* h = static_cast<Shim*>(p);
    std::memcpy(out, h->smoother.scene_crop.data(), sizeof(float) * elems);
}

void shim_gaussian(int n, double sigma, float* out)
{
    std::vector<float> k;
    gaussian_kernel_f32(n, sigma, k);
    std::memcpy(out, k.data(), sizeof(float) * n);
}

int shim_mesh_to_transform(const float* offsets, int w, int h, double* t) { return mesh2x2_to_transform(offsets, w, h, t) ? 1 : 0; }

void shim_set_to_homography(const double* H, float sw, float sh, int cols, int rows, float* out)
{
    Mesh m;
    mesh_set_to_homography(H, sw, sh, cols, rows, m);
    std::memcpy(out, m.data(), sizeof(float) * m.size());
}

int shim_local_motions(void* p, const float* tracked, const float* matched, int n, float* state, float* offsets, uint8_t* mask)
{
    Shim* h = static_cast<Shim*>(p);
    const size_t elems = static_cast<size_t>(2) * h->settings.motion_resolution_width * h->settings.motion_resolution_height;
    std::memcpy(h->solver.state().data(), state, sizeof(float) * elems);
    std::vector<float> a(tracked, tracked + 2 * n), b(matched, matched + 2 * n);
    Mesh off;
    std::vector<uint8_t> m;
    int iters = 0;
    h->solver.estimate(a, b, off, m, &iters);
    std::memcpy(state, h->solver.state().data(), sizeof(float) * elems);
    std::memcpy(offsets, off.data(), sizeof(float) * elems);
    std::memcpy(mask, m.data(), n);
    return iters;
}

}  // extern "C"
