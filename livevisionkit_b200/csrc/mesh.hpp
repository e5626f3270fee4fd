// K6c — device solver for FrameTracker::estimate_local_motions (LiveVisionKit/Vision/FrameTracker.cpp:200-321): the
// Eigen::LeastSquaresConjugateGradient solve of the motion-mesh system as ONE persistent CTA that follows the
// swap-erase compaction on the tracking stream.  Used for EVERY mesh that fits one CTA's shared memory: the OBS
// "Vector Field" preset (16x16 vertices -> 512 unknowns, ~135 CG iterations per frame = 2.7 ms in the host solver) and
// the library-default 2x2 mesh (8 unknowns), so no preset has a CPU solver on its path.  The host restatement
// (host_mesh.hpp) remains for meshes too large for a CTA and behind LVKB200_MESH_DEVICE_MIN=<unknowns> (debug knob).
#pragma once

#include <vector>

#include "common.hpp"
#include "ransac.hpp"
#include "stream.hpp"

namespace lvkb200
{

constexpr int MESH_DEVICE_MIN_UNKNOWNS = 8;
constexpr int MESH_CGLS_THREADS = 512;

// Static rows of the system that are not the (diagonal) temporal rows: the similarity constraints of
// generate_mesh_constraints (FrameTracker.cpp:380-457); exactly four non-zeros per row.
struct MeshStaticRows
{
    int mesh_cols = 0, mesh_rows = 0;  // vertices
    std::vector<int> col;              // 4 per row
    std::vector<float> val;            // 4 per row
    int rows() const { return static_cast<int>(col.size() / 4); }
};

// Header of the solver's result block (mirrored in mapped pinned host memory), followed by the mesh (2*cols*rows floats).
struct MeshSolveResult
{
    int solved;      // 0: fewer than min_samples correspondences -> nothing was estimated, the state is unchanged
    int iterations;  // CG iterations
    int n;           // correspondences the solver saw (after the device-side compaction)
    int pad;
};

struct MeshSolveParams
{
    float temporal_weight;  // ts
    float acceptance;       // inlier threshold on the L1 error
    float key_w, key_h;     // mesh_grid.key_size()
    int min_samples;        // settings.min_motion_samples
};

class MeshCgls
{
public:
    // Uploads the static system.  feature_capacity bounds the correspondences per frame.  False when the system does
    // not fit the CTA's shared memory (the caller then keeps the host solver).
    bool configure(const MeshStaticRows& sys, int feature_capacity, cudaStream_t cs, cudaError_t* err);
    bool ready() const { return n_unknowns > 0; }
    int unknowns() const { return n_unknowns; }
    cudaError_t reset_state(cudaStream_t cs);                    // FrameTracker::restart: mesh <- 0
    cudaError_t set_state(cudaStream_t cs, const float* mesh);  // stage-level entry point only (synchronous copy)
    // Enqueues the solve on cs: correspondences d_src -> d_dst (*d_n of them, in the reference's order), inlier mask to
    // d_mask; the result header + mesh go to the device block and its host mirror, and, with `out.host` set, the used
    // part of the tracking result block (LK matches + status, mask) is copied to ITS host mirror as well.
    cudaError_t launch(cudaStream_t cs, const MeshSolveParams& prm, const float2* d_src, const float2* d_dst,
                       const int* d_n, const TrackParams* d_params, uint8_t* d_mask, const TrackOutCopy& out);
    // After the stream has been synchronized.
    const MeshSolveResult& result() const { return *h_out.as<MeshSolveResult>(); }
    const float* mesh() const { return reinterpret_cast<const float*>(h_out.as<uint8_t>() + sizeof(MeshSolveResult)); }
    void release();

private:
    int n_unknowns = 0, n_sim = 0, mesh_cols = 0, mesh_rows = 0, capacity = 0, csc_nnz = 0;
    size_t smem_bytes = 0;
    int slots = -1;  // >= 0: k_mesh_cgls2 (1024 threads, chains) in that slot variant; -1: k_mesh_cgls
    DeviceBuffer d_static, d_state, d_out;
    PinnedBuffer h_out;
};

}  // namespace lvkb200
