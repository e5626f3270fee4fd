// lvkb200_stream definition (opaque to C callers): one video stream == one lvk::StabilizationFilter instance.
#pragma once

#include <cstring>
#include <vector>

#include "common.hpp"
#include "deblock.hpp"
#include "fast.hpp"
#include "formats.hpp"
#include "host_logic.hpp"
#include "host_mesh.hpp"
#include "ingest.hpp"
#include "lk.hpp"
#include "mesh.hpp"
#include "ransac.hpp"
#include "stream.hpp"

struct lvkb200_stream
{
    int device = 0;
    cudaStream_t cs = nullptr;
    lvkb200_settings settings{};
    bool configured = false;
    bool debug_capture = false;

    // ---- scratch used by the stage-level entry points when the caller hands host memory
    lvkb200::DeviceBuffer stage_in, stage_out, mesh_dev;
    lvkb200::DeviceBuffer scaling_scratch;  // ScalingFilter: the upscaled frame between EASU and RCAS
    lvkb200::PinnedBuffer mesh_pinned;
    cudaEvent_t user_events[LVKB200_EVENT_SLOTS] = {};

    // ---- per-frame host scratch of track(): cleared, never freed, so a steady-state frame allocates nothing
    struct TrackScratch
    {
        std::vector<lvkb200::FastRegion> regions;
        std::vector<int> region_index;
        std::vector<std::vector<lvkb200::FastPoint>> fast_points;
        std::vector<float> tracked, matched;
        std::vector<uint8_t> status, inliers;
    } scratch;

    // ---- DeblockingFilter chained in front of the stabilizer (lvkb200_stream_set_deblocking) / stand-alone
    lvkb200::DeblockPlan deblock, deblock_stage;
    bool deblock_enabled = false;
    lvkb200_deblock_settings deblock_settings{};

    // ---- FrameIngest (OBS plane layouts <-> packed frames): chroma tap tables + plane / frame scratch
    lvkb200::FormatPlan format_plan;
    lvkb200::DeviceBuffer planes_in, planes_out, obs_frame_in, obs_frame_out;
    // pipelined OBS-layout path (lvkb200_stream_prefetch_obs / _submit_obs_async): plane staging per prefetch slot and per
    // output slot, and the ingest's geometry tables prepared on the copy-in stream
    lvkb200::DeviceBuffer planes_in_async[2], planes_out_async[2];
    lvkb200::FormatPlan format_plan_in;

    // ---- device-side stages
    lvkb200::IngestPlan ingest;
    lvkb200::FastDetector fast;
    // Three pyramids rotate: pyr[cur] = current frame, pyr[prev_pyr()] = previous frame, pyr[next_pyr()] = the NEXT
    // frame's, built ahead of time when the caller announced it (pre_ingest)
    static constexpr int PYR_COUNT = 3;
    lvkb200::LkPyramid pyr[PYR_COUNT];
    int cur = 0;
    int prev_pyr() const { return (cur + PYR_COUNT - 1) % PYR_COUNT; }
    int next_pyr() const { return (cur + 1) % PYR_COUNT; }
    cudaEvent_t track_done = nullptr;  // recorded on cs behind the tracking chain: the host waits for THIS, not for cs
    lvkb200::DeviceBuffer d_det;
    size_t det_pitch = 0;
    lvkb200::DeviceBuffer d_pts_prev, d_src, d_dst, d_models, d_scores;
    lvkb200::DeviceBuffer d_perm, d_removed, d_count, d_params;
    // results of the tracking chain, contiguous so that ONE device->host copy brings them back:
    // [matched float2 x cap | status u8 x cap | mask u8 x cap | pad | RansacResult]
    lvkb200::DeviceBuffer d_track_out;
    lvkb200::PinnedBuffer h_track_out, h_pts_prev, h_src, h_dst, h_det, h_count, h_params;
    size_t off_status = 0, off_mask = 0, off_result = 0, track_out_bytes = 0;
    int point_capacity = 0;
    float2* d_pts_next() const { return d_track_out.as<float2>(); }
    uint8_t* d_status() const { return d_track_out.as<uint8_t>() + off_status; }
    uint8_t* d_mask() const { return d_track_out.as<uint8_t>() + off_mask; }
    lvkb200::RansacResult* d_result() const
    {
        return reinterpret_cast<lvkb200::RansacResult*>(d_track_out.as<uint8_t>() + off_result);
    }
    // The tracking chain (H2D params+points -> LK -> swap-erase -> RANSAC -> one D2H) replayed as ONE CUDA graph per
    // (frame parity, estimator): ~10 API calls and their inter-kernel gaps become one launch.
    cudaGraphExec_t track_graph[2][2] = {};
    bool use_graphs = true;
    bool profile_stages = false;  // per-stage CUDA events (eager path); off by default
    void destroy_graphs();
    lvkb200_status enqueue_tracking(const std::vector<float>& pts, bool global, float threshold, int model);
    lvkb200::LkPack lk_pack{};  // this frame's parameters + points: passed to the LK kernel by value
    lvkb200_status launch_lk(int parity, bool global, int n, bool with_events);
    lvkb200_status record_estimator_chain(bool with_events);

    // ---- StabilizationFilter / FrameTracker / FeatureDetector / PathSmoother host state
    lvkb200::FeatureGrid grid;
    lvkb200::PathSmoother smoother;
    lvkb200::MeshSolver mesh_solver;
    // K6c: every mesh (>= MESH_DEVICE_MIN_UNKNOWNS = 8 unknowns, the 2x2 default included) is solved by one CTA on the
    // tracking stream (mesh.cu)
    lvkb200::MeshCgls mesh_device;
    unsigned mesh_device_generation = ~0u;  // mesh_solver.generation() the device copy of the static rows was made from
    int mesh_device_capacity = 0;
    bool mesh_device_unfit = false;  // the system does not fit one CTA's shared memory: host solver
    int mesh_device_min_unknowns = lvkb200::MESH_DEVICE_MIN_UNKNOWNS;  // LVKB200_MESH_DEVICE_MIN overrides (tuning knob)
    bool mesh_on_device() const { return settings.track_local_motions && !mesh_device_unfit &&
                                         mesh_solver.unknowns() >= mesh_device_min_unknowns; }
    lvkb200_status prepare_mesh_device();  // (re)uploads the static rows when the settings or the capacity changed
    lvkb200_status launch_mesh_device();   // compaction + solve behind LK on cs
    int last_mesh_iterations = 0;
    std::vector<lvkb200::Feature> features;  // m_TrackedFeatures
    bool frame_initialized = false;
    int lk_calls = 0;  // calc() calls made on this tracker's cv::SparsePyrLKOpticalFlow equivalent
    int det_w = 0, det_h = 0;
    float tracking_stability = 0.f, scene_quality = 0.f, trust_factor = 0.f;

    // ---- m_FrameQueue: device-resident ring of full-resolution frames
    struct QueuedFrame
    {
        lvkb200::DeviceBuffer buf;
        size_t pitch = 0;
        int w = 0, h = 0;
        lvkb200_format format = LVKB200_UNKNOWN;
        uint64_t timestamp = 0;
    };
    std::vector<QueuedFrame> ring;
    size_t ring_start = 0, ring_size = 0;
    cudaEvent_t input_copied = nullptr;  // the caller's (host) frame has been consumed

    // ---- pipelined operation (the GPU analogue of VideoFilter::stream's three threads, Filters/VideoFilter.cpp:62-209):
    // a copy-in stream uploads the NEXT host frame into a spare buffer while this frame is processed, and a copy-out
    // stream downloads the PREVIOUS output from one of two staging buffers.
    cudaStream_t cs_in = nullptr, cs_out = nullptr;
    // two spare buffers: the caller prefetches frame t+1 BEFORE submitting frame t, so two uploads are outstanding
    QueuedFrame prefetch_slot[2];
    // Look-ahead (lvkb200_stream_prefetch_frame): the announced frame's format is known, so its detection image and
    // pyramid are built on cs right behind the CURRENT frame's tracking chain — while the host digests that chain's
    // results the GPU would otherwise idle — and the next submit finds them ready (pre_ingest / track).
    struct Lookahead
    {
        bool announced = false;  // format known: eligible for pre_ingest
        lvkb200_format format = LVKB200_UNKNOWN;
        bool built = false;      // detection image + pyr[pyr_index] hold this frame
        int pyr_index = -1;
        bool deblocked = false;  // the chained deblocking stage already ran on the slot's buffer
    } lookahead[2];
    lvkb200_status pre_ingest();
    bool frame_prebuilt = false;  // the frame being submitted was adopted from a slot whose look-ahead was built
    const void* prefetched_ptr[2] = {nullptr, nullptr};
    cudaEvent_t prefetch_done[2] = {nullptr, nullptr};  // recorded on cs_in after each upload
    int prefetch_next = 0;
    // The output remap runs on its OWN stream: it depends only on a frame uploaded `predictive_samples` submits ago and
    // on the host-computed correction, so it overlaps the next frame's (latency-bound) tracking chain on `cs`.
    // Frame buffers rotate through `spare_buf`: the buffer the latest remap reads is parked there, so every buffer an
    // upload can target was last read by the remap BEFORE the latest one (wait_frame_buffers_free).
    cudaStream_t cs_remap = nullptr;
    cudaEvent_t remap_done[2] = {nullptr, nullptr};  // remap number n records remap_done[n & 1] on cs_remap
    uint64_t remaps_launched = 0;
    cudaEvent_t chain_point = nullptr;  // recorded on cs before a remap is queued: uploads on cs have been issued
    // recorded on cs BEFORE the frame's tracking chain is enqueued: a held-back remap that waits on it runs BESIDE the
    // chain instead of behind it.  Measured (tools/gpu_overlap_ab.sh): the remap then owns every register of every SM
    // and the chain's small kernels wait for its CTAs to retire - at 1080p (remap 39 us < chain 52 us) the step gets
    // longer (11.2k -> 9.9k fps), at 4K (remap 133 us >> chain) it gets shorter (5.0k -> 5.5k fps).  Hence by frame size;
    // LVKB200_REMAP_OVERLAP=0/1 forces either order.
    cudaEvent_t pre_chain = nullptr;
    bool pre_chain_valid = false, overlap_after_lk = false;
    int graph_kernels = 0;  // kernels inside the captured estimator graph (counted once per replay)
    int remap_overlap = -1;  // -1: by frame size (>= REMAP_OVERLAP_MIN_PIXELS)
    static constexpr long long REMAP_OVERLAP_MIN_PIXELS = 3000000;
    lvkb200::DeviceBuffer spare_buf;
    // A remap whose pixels nobody waits for inside submit (device output, pipelined host output) is held back and
    // launched behind the NEXT frame's LK + RANSAC graph, where the SMs are idle (see apply_mesh / flush_remap).
    struct PendingRemap
    {
        bool active = false;
        lvkb200::RemapParams p{};
        bool homography = true;
        float tf[9] = {};
        const float* dmesh = nullptr;
        bool async_host_out = false;
        int slot = 0;
        uint64_t ticket = 0;
        void* out = nullptr;
        size_t out_pitch = 0;
        lvkb200_memspace out_space = LVKB200_MEM_DEVICE;
        // pipelined OBS-layout output: instead of downloading the packed frame, the copy-out stream converts it back to
        // the planes of `egress_frame` (to_obs) and downloads those
        bool egress = false;
        lvkb200_obs_frame egress_frame{};
    } pending;
    // set by lvkb200_stream_submit_obs_async around submit(): the output of that submit leaves as OBS planes
    bool next_egress = false;
    lvkb200_obs_frame next_egress_frame{};
    // formats_api.cpp: packed device frame -> planes of `dst` (host memory), all on `stream`, staging slot `slot`
    lvkb200_status egress_planes(cudaStream_t stream, const uint8_t* packed, size_t pitch, int width, int height,
                                 lvkb200_format format, const lvkb200_obs_frame& dst, int slot);
    lvkb200_status flush_remap();
    // Allocates every still-empty frame buffer of the rotation (ring, prefetch slots, parked buffer) at once.  Lazily,
    // each of the first ~14 frames of a stream paid a cudaMalloc — and, in pipelined operation, a full sync_all in
    // front of it — exactly while a short measurement (or a live stream's first quarter second) was running.
    lvkb200_status ensure_frame_pool(size_t bytes);
    size_t frame_pool_bytes = 0;
    lvkb200_status wait_frame_buffers_free(cudaStream_t stream);
    lvkb200_status join_remap(cudaStream_t stream);  // makes `stream` wait for every remap queued so far
    lvkb200_status sync_all();                       // host waits for cs and cs_remap
    lvkb200::DeviceBuffer async_out[2];
    cudaEvent_t async_remap_done[2] = {}, async_out_done[2] = {};
    bool async_out_used[2] = {false, false};
    uint64_t async_tickets = 0;
    bool deferred_output = false;  // set for the duration of a submit_async call
    uint64_t last_ticket = 0;
    lvkb200_status prefetch(const void* frame, size_t pitch, int width, int height, lvkb200_format format = LVKB200_UNKNOWN,
                            lvkb200_memspace space = LVKB200_MEM_HOST);
    lvkb200_status wait_output(uint64_t ticket);
    lvkb200_status ensure_pipeline();

    // ---- per-stage CUDA events of the last submit (ingest, pyramid, fast, lk, estimate, remap)
    // Double-buffered by frame parity so that totals can be harvested two frames late without any extra sync.
    cudaEvent_t stage_ev[2][LVKB200_STAGE_COUNT][2] = {};
    bool stage_used[2][LVKB200_STAGE_COUNT] = {};
    int stage_parity = 0;
    double stage_total_us[LVKB200_STAGE_COUNT] = {};
    uint64_t stage_count[LVKB200_STAGE_COUNT] = {};
    void harvest_stage_times(int parity);
    lvkb200_status stage_totals(double* totals, uint64_t* counts, bool reset);

    // ---- host-side timeline of submit (LVKB200_HOST_TRACE=1): where the CPU thread spends its time per frame;
    // printed to stderr when the stream is destroyed.  Diagnostic only.
    enum HostPhase { HP_UPLOAD, HP_ENQ_DETECT, HP_WAIT_FAST, HP_DETECT, HP_ENQ_TRACK, HP_WAIT_TRACK, HP_POST,
                     HP_SMOOTH, HP_REMAP, HP_OUTSIDE, HP_COUNT };
    bool host_trace = false;
    double host_us[HP_COUNT] = {};
    uint64_t host_frames = 0;
    double host_mark = 0.0;
    void host_tick(int phase);  // adds the time since the previous tick to `phase`

    // ---- debug taps of the last submit
    std::vector<uint8_t> dbg_det, dbg_lk_status, dbg_inliers;
    std::vector<lvkb200_keypoint> dbg_detected, dbg_propagated;
    std::vector<float> dbg_lk_matched, dbg_tracked, dbg_matched, dbg_motion, dbg_correction;
    std::vector<int32_t> dbg_fast_counts;
    double dbg_h[9] = {}, dbg_t[9] = {};
    bool dbg_has_h = false, dbg_has_t = false;

    lvkb200_status configure(const lvkb200_settings& s);
    lvkb200_status restart();
    lvkb200_status reset_context();
    bool ready() const { return ring_size == ring.size() && !ring.empty(); }
    void stable_region(int fw, int fh, int* x, int* y, int* w, int* h) const;
    lvkb200_status submit(const void* frame, size_t pitch, int width, int height, lvkb200_format format,
                          uint64_t timestamp, lvkb200_memspace frame_space, void* out, size_t out_pitch,
                          lvkb200_memspace out_space, lvkb200_result* res);
    lvkb200_status debug_fetch(lvkb200_debug_item which, void* buffer, size_t capacity, size_t* size);
    lvkb200_status stage_times(float* times);
    void release();

    // FrameTracker::track on the frame in `slot`; fills motion (empty == nullopt).
    lvkb200_status track(const QueuedFrame& frame, lvkb200::Mesh& motion, bool* has_motion);
    lvkb200_status ensure_points(int n);
    lvkb200_status fetch_tracking(int n, bool with_model, std::vector<float>& matched, std::vector<uint8_t>& status,
                                  lvkb200::RansacResult* model, std::vector<uint8_t>& mask);
    lvkb200_status run_local_motions(const std::vector<float>& tracked, const std::vector<float>& matched,
                                     float* mesh_state, lvkb200::Mesh& offsets, std::vector<uint8_t>& mask);
    lvkb200_status run_homography(const std::vector<float>& tracked, const std::vector<float>& matched, float threshold, int model,
                                  double h[9], std::vector<uint8_t>& mask, bool* found);
    lvkb200_status apply_mesh(QueuedFrame& src, const lvkb200::Mesh& offsets, void* out, size_t out_pitch,
                              lvkb200_memspace out_space);
    void stage_begin(int stage, cudaStream_t on = nullptr);  // nullptr == cs
    void stage_end(int stage, cudaStream_t on = nullptr);

    // Makes a device view of a caller frame: device memory is used in place, host memory is copied
    // (async, on this stream) into stage_in with a 16-byte aligned pitch.
    lvkb200_status stage_frame_in(const void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                  const uint8_t** dptr, size_t* dpitch);
    // Device destination for a caller frame (caller's own memory if it is device memory).
    lvkb200_status stage_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                   uint8_t** dptr, size_t* dpitch);
    // Copies the staged output back to host memory and waits (host destination only).
    lvkb200_status finish_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space);
    lvkb200_status upload_mesh(const float* offsets, int cols, int rows, const float** dmesh);
};
