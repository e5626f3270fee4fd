// lvkb200_stream: per-frame orchestration of the stabilization path (host C++), state and scratch.
// Mirrors StabilizationFilter::filter (LiveVisionKit/Filters/StabilizationFilter.cpp:69-135) and
// FrameTracker::track (LiveVisionKit/Vision/FrameTracker.cpp:108-196); every pixel- or point-parallel step is a
// CUDA kernel (ingest.cu, lk.cu, fast.cu, ransac.cu, remap.cu), the order-dependent bookkeeping is host_logic.hpp.
#include "stream_impl.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "host_math.hpp"

using namespace lvkb200;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr uint64_t HOST_TRACE_SKIP = 40;

static inline double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void lvkb200_stream::host_tick(int phase)
{
    if (!host_trace) return;
    const double t = now_us();
    if (phase >= 0 && host_frames > HOST_TRACE_SKIP) host_us[phase] += t - host_mark;  // skip the start-up frames
    host_mark = t;
}

constexpr float QA_UPDATE_RATE = 0.1f;  // StabilizationFilter.cpp:29
constexpr float QA_BLEND_STEP = 0.05f;  // StabilizationFilter.cpp:30
constexpr float HOMOGRAPHY_DISTRIBUTION_THRESHOLD = 0.6f;  // FrameTracker.cpp:37

enum Stage { ST_INGEST = 0, ST_PYRAMID, ST_FAST, ST_LK, ST_ESTIMATE, ST_REMAP };
static_assert(ST_REMAP + 1 == LVKB200_STAGE_COUNT, "stage table out of step with lvkb200.h");

lvkb200_status lvkb200_stream::configure(const lvkb200_settings& s)
{
    // Preconditions of StabilizationFilter::configure (StabilizationFilter.cpp:44-45), FrameTracker::configure
    // (FrameTracker.cpp:59-65), FeatureDetector::configure (FeatureDetector.cpp:50-57), PathSmoother::configure
    // (PathSmoother.cpp:38-44).
    LVKB_REQUIRE(s.min_tracking_quality >= 0.0f && s.min_tracking_quality <= 1.0f);
    LVKB_REQUIRE(s.min_scene_quality >= 0.0f && s.min_scene_quality <= 1.0f);
    LVKB_REQUIRE(s.motion_resolution_width >= 2 && s.motion_resolution_height >= 2);
    LVKB_REQUIRE(s.acceptance_threshold >= 0.0f);
    LVKB_REQUIRE(s.temporal_smoothing >= 0.0f);
    LVKB_REQUIRE(s.local_smoothing >= 0.0f);
    LVKB_REQUIRE(s.min_motion_samples >= 4);
    LVKB_REQUIRE(s.uniformity_threshold >= 0.0f && s.uniformity_threshold <= 1.0f);
    LVKB_REQUIRE(s.detection_regions_width > 0 && s.detection_regions_height > 0);
    LVKB_REQUIRE(s.detection_regions_height <= s.detection_resolution_height);
    LVKB_REQUIRE(s.detection_regions_width <= s.detection_resolution_width);
    LVKB_REQUIRE(s.detection_regions_width * s.detection_regions_height <= FAST_MAX_REGIONS);
    LVKB_REQUIRE(s.min_feature_density <= s.max_feature_density);
    LVKB_REQUIRE(s.min_feature_density > 0.0f);
    LVKB_REQUIRE(s.accumulation_rate > 0.0f);
    LVKB_REQUIRE(s.max_feature_density >= 0.0f && s.max_feature_density <= 1.0f);
    LVKB_REQUIRE(s.min_feature_density >= 0.0f && s.min_feature_density <= 1.0f);
    LVKB_REQUIRE(s.corrective_limits_width >= 0.0f && s.corrective_limits_width <= 1.0f);
    LVKB_REQUIRE(s.corrective_limits_height >= 0.0f && s.corrective_limits_height <= 1.0f);
    LVKB_REQUIRE(s.predictive_samples > 0);
    LVKB_REQUIRE(s.smoothing_steps > 0.0f);
    LVKB_REQUIRE(s.response_rate >= 0.0f && s.response_rate <= 1.0f);

    // StabilizationFilter.cpp:51-52: disabling the stabilization resets the context
    if (configured && settings.stabilize_output && !s.stabilize_output) LVKB_TRY(reset_context());

    if (!configured)
    {
        host_trace = std::getenv("LVKB200_HOST_TRACE") != nullptr;
        if (const char* e = std::getenv("LVKB200_REMAP_OVERLAP")) remap_overlap = std::min(std::max(std::atoi(e), 0), 2);
        if (const char* e = std::getenv("LVKB200_MESH_DEVICE_MIN")) mesh_device_min_unknowns = std::atoi(e);
    }
    const bool det_changed = !configured || s.detection_resolution_width != settings.detection_resolution_width ||
                             s.detection_resolution_height != settings.detection_resolution_height;
    settings = s;

    smoother.configure(s);

    // m_FrameQueue.resize(time_delay + 1): StreamBuffer::resize keeps the newest frames (Data/StreamBuffer.tpp:206-221)
    const size_t capacity = static_cast<size_t>(s.predictive_samples) + 1;
    if (ring.size() != capacity)
    {
        LVKB_TRY(sync_all());
        std::vector<QueuedFrame> fresh(capacity);
        const size_t keep = std::min(ring_size, capacity);
        for (size_t i = 0; i < keep; i++)
            std::swap(fresh[i], ring[(ring_start + (ring_size - keep) + i) % ring.size()]);
        for (auto& f : ring) f.buf.release();
        ring.swap(fresh);
        frame_pool_bytes = 0;  // fresh slots are empty: the next submit fills the pool again
        ring_start = 0;
        ring_size = keep;
    }

    grid.configure(s);
    mesh_solver.configure(s);
    mesh_device_unfit = false;  // re-evaluated by prepare_mesh_device for the new mesh
    det_w = s.detection_resolution_width;
    det_h = s.detection_resolution_height;
    if (det_changed)
    {
        // FrameTracker.cpp:84-90 rescales the previous detection frame; we restart tracking instead (one motion
        // sample is skipped after a detection-resolution change; documented in DESIGN.md).
        features.clear();
        grid.reset();
        frame_initialized = false;
        for (auto& py : pyr) py.valid = false;
        for (auto& la : lookahead) la.built = false;
        if (cs) cudaStreamSynchronize(cs);
        destroy_graphs();   // the pyramids will be re-allocated for the new detection resolution
        point_capacity = 0;  // and the point capacity follows the new suppression grid
    }
    configured = true;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::restart()
{
    // StabilizationFilter::restart — StabilizationFilter.cpp:139-144
    // an output that was already handed out is completed first (its remap may still be held back): after restart()
    // nothing of this stream writes into a caller's buffer any more
    LVKB_TRY(flush_remap());
    scene_quality = 1.0f;
    ring_start = 0;
    ring_size = 0;
    return reset_context();
}

lvkb200_status lvkb200_stream::reset_context()
{
    // FrameTracker::restart (FrameTracker.cpp:97-104) + PathSmoother::restart (PathSmoother.cpp:139-145)
    tracking_stability = 0.0f;
    features.clear();
    grid.reset();
    frame_initialized = false;
    mesh_solver.restart();
    if (mesh_device.ready()) LVKB_CUDA(mesh_device.reset_state(cs));
    smoother.restart();
    return LVKB200_OK;
}

void lvkb200_stream::stable_region(int fw, int fh, int* x, int* y, int* w, int* h) const
{
    // StabilizationFilter::stable_region (StabilizationFilter.cpp:199-): scene margins scaled to the frame.
    *x = static_cast<int>(std::lrintf(smoother.margin_x * static_cast<float>(fw)));
    *y = static_cast<int>(std::lrintf(smoother.margin_y * static_cast<float>(fh)));
    *w = static_cast<int>(std::lrintf(smoother.margin_w * static_cast<float>(fw)));
    *h = static_cast<int>(std::lrintf(smoother.margin_h * static_cast<float>(fh)));
}

lvkb200_status lvkb200_stream::wait_frame_buffers_free(cudaStream_t stream)
{
    // The latest remap issued (launched or still held back) reads spare_buf only; every other frame buffer was last
    // read by the one before it or earlier.
    const uint64_t issued = remaps_launched + (pending.active ? 1 : 0);
    if (issued >= 2) LVKB_CUDA(cudaStreamWaitEvent(stream, remap_done[(issued - 2) & 1], 0));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::ensure_frame_pool(size_t bytes)
{
    if (bytes <= frame_pool_bytes) return LVKB200_OK;
    // only EMPTY buffers are touched: a buffer that holds a queued frame (of a smaller geometry) keeps it and grows the
    // old way, when its slot is overwritten
    for (QueuedFrame& q : ring)
        if (q.buf.capacity == 0) LVKB_CUDA(q.buf.ensure(bytes));
    for (QueuedFrame& p : prefetch_slot)
        if (p.buf.capacity == 0) LVKB_CUDA(p.buf.ensure(bytes));
    if (spare_buf.capacity == 0) LVKB_CUDA(spare_buf.ensure(bytes));
    frame_pool_bytes = bytes;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::join_remap(cudaStream_t stream)
{
    LVKB_TRY(flush_remap());
    if (remaps_launched >= 1) LVKB_CUDA(cudaStreamWaitEvent(stream, remap_done[(remaps_launched - 1) & 1], 0));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::sync_all()
{
    LVKB_TRY(flush_remap());
    if (cs) LVKB_CUDA(cudaStreamSynchronize(cs));
    if (cs_remap) LVKB_CUDA(cudaStreamSynchronize(cs_remap));
    return LVKB200_OK;
}

void lvkb200_stream::stage_begin(int stage, cudaStream_t on)
{
    // Only the roofline kernel (remap) is timed unconditionally; the other stages are timed when profiling is on.
    if (!profile_stages && stage != ST_REMAP) return;
    cudaEvent_t* ev = stage_ev[stage_parity][stage];
    if (!ev[0])
    {
        cudaEventCreate(&ev[0]);
        cudaEventCreate(&ev[1]);
    }
    cudaEventRecord(ev[0], on ? on : cs);
    stage_used[stage_parity][stage] = true;
}

void lvkb200_stream::stage_end(int stage, cudaStream_t on)
{
    if (!profile_stages && stage != ST_REMAP) return;
    cudaEventRecord(stage_ev[stage_parity][stage][1], on ? on : cs);
}

// Adds the (completed) stage durations recorded in slot `parity` to the running totals and frees the slot.
void lvkb200_stream::harvest_stage_times(int parity)
{
    for (int i = 0; i < LVKB200_STAGE_COUNT; i++)
    {
        if (!stage_used[parity][i]) continue;
        stage_used[parity][i] = false;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, stage_ev[parity][i][0], stage_ev[parity][i][1]) == cudaSuccess)
        {
            stage_total_us[i] += static_cast<double>(ms) * 1000.0;
            stage_count[i]++;
        }
        else
            cudaGetLastError();
    }
}

lvkb200_status lvkb200_stream::stage_totals(double* totals, uint64_t* counts, bool reset)
{
    LVKB_TRY(sync_all());
    harvest_stage_times(0);
    harvest_stage_times(1);
    for (int i = 0; i < LVKB200_STAGE_COUNT; i++)
    {
        totals[i] = stage_total_us[i];
        counts[i] = stage_count[i];
        if (reset)
        {
            stage_total_us[i] = 0.0;
            stage_count[i] = 0;
        }
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::stage_times(float* times)
{
    // durations of the LAST submit (slot of the previous parity), read without consuming them
    LVKB_TRY(sync_all());
    const int p = stage_parity;
    for (int i = 0; i < LVKB200_STAGE_COUNT; i++)
    {
        times[i] = 0.0f;
        if (stage_used[p][i] && stage_ev[p][i][0])
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, stage_ev[p][i][0], stage_ev[p][i][1]) == cudaSuccess) times[i] = ms * 1000.0f;
            else cudaGetLastError();
        }
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::ensure_points(int n)
{
    if (n <= point_capacity) return LVKB200_OK;
    // capacity = the detector's maximum feature count (so the tracking graph is captured once), rounded up
    const int want = std::max(n, static_cast<int>(grid.max_feature_capacity()));
    const int cap = (std::max(want, 256) + 63) / 64 * 64;
    LVKB_CUDA(cudaStreamSynchronize(cs));
    destroy_graphs();
    off_status = sizeof(float2) * static_cast<size_t>(cap);
    off_mask = off_status + static_cast<size_t>(cap);
    off_result = align_up(off_mask + static_cast<size_t>(cap), 16);
    track_out_bytes = align_up(off_result + sizeof(RansacResult), 16);  // moved as whole 16-byte words
    LVKB_CUDA(d_track_out.ensure(track_out_bytes));
    LVKB_CUDA(h_track_out.ensure(track_out_bytes));
    LVKB_CUDA(d_pts_prev.ensure(sizeof(float2) * cap));
    LVKB_CUDA(d_src.ensure(sizeof(float2) * cap));
    LVKB_CUDA(d_dst.ensure(sizeof(float2) * cap));
    LVKB_CUDA(d_models.ensure(sizeof(float) * 9 * RANSAC_HYPOTHESES));
    if (d_scores.capacity < sizeof(float) * (RANSAC_HYPOTHESES + 4))
    {
        // + the packed (score, index) arg-min key of the scoring pass, created "armed" (all ones); the refine kernel re-arms it
        LVKB_CUDA(d_scores.ensure(sizeof(float) * (RANSAC_HYPOTHESES + 4)));
        LVKB_CUDA(cudaMemsetAsync(d_scores.ptr, 0xFF, sizeof(float) * (RANSAC_HYPOTHESES + 4), cs));
    }
    LVKB_CUDA(d_perm.ensure(sizeof(int) * cap));
    LVKB_CUDA(d_removed.ensure(sizeof(int) * cap));
    LVKB_CUDA(d_count.ensure(sizeof(int)));
    LVKB_CUDA(d_params.ensure(align_up(sizeof(TrackParams), 16)));
    LVKB_CUDA(h_count.ensure(sizeof(int)));
    LVKB_CUDA(h_params.ensure(align_up(sizeof(TrackParams), 16)));
    LVKB_CUDA(h_pts_prev.ensure(sizeof(float2) * cap));
    LVKB_CUDA(h_src.ensure(sizeof(float2) * cap));
    LVKB_CUDA(h_dst.ensure(sizeof(float2) * cap));
    std::memset(h_pts_prev.ptr, 0, sizeof(float2) * cap);
    point_capacity = cap;
    return LVKB200_OK;
}

void lvkb200_stream::destroy_graphs()
{
    for (auto& row : track_graph)
        for (auto& g : row)
        {
            if (g) cudaGraphExecDestroy(g);
            g = nullptr;
        }
}

// The device work of the tracking chain, in stream order on `cs` (either captured into a graph or executed eagerly):
// LK (fetches the parameters + points from mapped pinned host memory itself) -> [swap-erase compaction -> RANSAC];
// LK and the estimator's last kernel also write their results into mapped pinned host memory, so the chain has no
// transfer steps of its own and never touches a copy engine.
lvkb200_status lvkb200_stream::launch_lk(int parity, bool global, int n, bool with_events)
{
    // LK is launched eagerly: the frame's parameters and points ride the launch as kernel arguments (lk_pack).
    uint8_t* hout = h_track_out.device_view<uint8_t>();
    LVKB_REQUIRE(hout != nullptr);
    LkIo io{};
    io.pts_copy = d_pts_prev.as<float2>();
    io.prm_copy = d_params.as<TrackParams>();
    io.next = d_pts_next();
    io.status = d_status();
    const bool inline_points = point_capacity <= LK_INLINE_POINTS;
    if (!inline_points)
    {
        // more points than the parameter space holds: the kernel reads them from mapped pinned host memory (slow path)
        std::memcpy(h_params.ptr, &lk_pack.prm, sizeof(TrackParams));
        io.pts_in = h_pts_prev.device_view<float2>();
        io.prm_in = h_params.device_view<TrackParams>();
    }
    if (with_events) stage_begin(ST_LK);
    LVKB_TRY(lk_track(cs, pyr[(parity + PYR_COUNT - 1) % PYR_COUNT], pyr[parity], (n + 3) / 4 * 4, io,
                      inline_points ? &lk_pack : nullptr));
    if (overlap_after_lk && pending.active && cs_remap)
    {
        // the held-back remap may start once LK has finished: it then runs beside the compaction / scoring / refine
        // kernels (one or a few CTAs, the machine is idle) but not beside LK (716 warps that want every SM)
        if (!pre_chain) LVKB_CUDA(cudaEventCreateWithFlags(&pre_chain, cudaEventDisableTiming));
        LVKB_CUDA(cudaEventRecord(pre_chain, cs));
        pre_chain_valid = true;
    }
    if (!global && mesh_on_device())
    {
        // swap-erase compaction -> k_mesh_cgls, which also delivers the LK results
        if (with_events) { stage_end(ST_LK); stage_begin(ST_ESTIMATE); }
        LVKB_TRY(launch_mesh_device());
        if (with_events) stage_end(ST_ESTIMATE);
        return LVKB200_OK;
    }
    if (!global)
    {
        // no estimator kernel follows (small meshes are solved on the host): one small CTA delivers the LK results
        TrackOutCopy out{};
        out.dev = d_track_out.as<uint8_t>();
        out.host = hout;
        out.off_status = static_cast<uint32_t>(off_status);
        LVKB_TRY(track_out_copy(cs, d_params.as<TrackParams>(), out));
    }
    if (with_events) stage_end(ST_LK);
    return LVKB200_OK;
}

// Local-motion estimator on the device (FrameTracker.cpp:200-321): fast_filter compaction, then the whole LSCG solve
// in one CTA; its last step copies mesh, mask and LK results into mapped pinned host memory.
lvkb200_status lvkb200_stream::prepare_mesh_device()
{
    if (mesh_device.ready() && mesh_device_generation == mesh_solver.generation() &&
        mesh_device_capacity == point_capacity)
        return LVKB200_OK;
    MeshStaticRows sys;
    sys.mesh_cols = mesh_solver.mesh_cols();
    sys.mesh_rows = mesh_solver.mesh_rows();
    mesh_solver.export_static(sys.col, sys.val);
    cudaError_t err = cudaSuccess;
    const bool had_state = mesh_device.ready() && mesh_device.unknowns() == mesh_solver.unknowns();
    if (!mesh_device.configure(sys, point_capacity, cs, &err))
    {
        LVKB_CUDA(err);
        mesh_device_unfit = true;  // too large for one CTA: the host solver takes over
        return LVKB200_OK;
    }
    // a fresh device solver starts from the host's current solution (zeros after a restart / resize)
    if (!had_state) LVKB_CUDA(mesh_device.set_state(cs, mesh_solver.state().data()));
    mesh_device_generation = mesh_solver.generation();
    mesh_device_capacity = point_capacity;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::launch_mesh_device()
{
    TrackOutCopy out{};
    out.dev = d_track_out.as<uint8_t>();
    out.host = h_track_out.device_view<uint8_t>();
    out.off_status = static_cast<uint32_t>(off_status);
    out.off_mask = static_cast<uint32_t>(off_mask);
    out.off_result = static_cast<uint32_t>(off_result);
    LVKB_REQUIRE(out.host != nullptr);
    const TrackParams* prm = d_params.as<TrackParams>();
    LVKB_TRY(compact_swap_erase(cs, d_pts_prev.as<float2>(), d_pts_next(), d_status(), prm, d_src.as<float2>(),
                                d_dst.as<float2>(), d_perm.as<int>(), d_removed.as<int>(), d_count.as<int>()));
    MeshSolveParams mp{};
    mp.temporal_weight = mesh_solver.temporal_weight();
    mp.acceptance = mesh_solver.acceptance_threshold();
    mp.key_w = mesh_solver.key_w();
    mp.key_h = mesh_solver.key_h();
    mp.min_samples = static_cast<int>(settings.min_motion_samples);
    LVKB_CUDA(mesh_device.launch(cs, mp, d_src.as<float2>(), d_dst.as<float2>(), d_count.as<int>(), prm, d_mask(), out));
    return LVKB200_OK;
}

// The estimator's device work in stream order on `cs` (captured into a graph or executed eagerly):
// swap-erase compaction -> hypotheses -> scores -> refine, whose last kernel also copies the chain's results into
// mapped pinned host memory: no transfer step of its own, no copy engine.
lvkb200_status lvkb200_stream::record_estimator_chain(bool with_events)
{
    const TrackParams* prm = d_params.as<TrackParams>();
    TrackOutCopy out{};
    out.dev = d_track_out.as<uint8_t>();
    out.host = h_track_out.device_view<uint8_t>();
    out.off_status = static_cast<uint32_t>(off_status);
    out.off_mask = static_cast<uint32_t>(off_mask);
    out.off_result = static_cast<uint32_t>(off_result);
    LVKB_REQUIRE(out.host != nullptr);
    // fast_filter + motion estimation chained on the device; the host replays the same erase order afterwards.
    // The model (homography / partial affine, FrameTracker.cpp:167-176) is selected by TrackParams::model.
    if (with_events) stage_begin(ST_ESTIMATE);
    LVKB_TRY(compact_swap_erase(cs, d_pts_prev.as<float2>(), d_pts_next(), d_status(), prm, d_src.as<float2>(),
                                d_dst.as<float2>(), d_perm.as<int>(), d_removed.as<int>(), d_count.as<int>()));
    LVKB_TRY(ransac_homography(cs, d_src.as<float2>(), d_dst.as<float2>(), d_count.as<int>(), prm,
                               d_models.as<float>(), d_scores.as<float>(), d_result(), d_mask(), out));
    if (with_events) stage_end(ST_ESTIMATE);
    return LVKB200_OK;
}

// Enqueues the tracking chain for this frame's points (previous -> current pyramid).
lvkb200_status lvkb200_stream::enqueue_tracking(const std::vector<float>& pts, bool global, float threshold, int model)
{
    const int n = static_cast<int>(pts.size() / 2);
    LVKB_TRY(ensure_points(n));
    if (!global && mesh_on_device()) LVKB_TRY(prepare_mesh_device());
    TrackParams& hp = lk_pack.prm;
    hp.n = n;
    hp.model = model;
    hp.lk_epsilon_sq = lk_epsilon_for_call(lk_calls);
    hp.threshold_sq = threshold * threshold;
    lk_calls = std::min(lk_calls + 1, 64);  // one m_OpticalTracker per FrameTracker: never reset (FrameTracker.cpp:41)
    lk_pack.inline_points = 1;
    if (point_capacity <= LK_INLINE_POINTS)
        std::memcpy(lk_pack.pts, pts.data(), sizeof(float) * pts.size());
    else
        std::memcpy(h_pts_prev.ptr, pts.data(), sizeof(float) * pts.size());

    LVKB_TRY(launch_lk(cur, global, n, profile_stages));
    if (!global) return LVKB200_OK;
    if (!use_graphs || profile_stages) return record_estimator_chain(profile_stages);

    // compaction + RANSAC (3 kernels) replay as ONE CUDA graph: one launch instead of three, no inter-kernel API gaps
    cudaGraphExec_t& exec = track_graph[0][1];
    if (!exec)
    {
        cudaGraph_t graph = nullptr;
        LVKB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        begin_launch_capture();
        const lvkb200_status st = record_estimator_chain(false);
        graph_kernels = end_launch_capture();  // credited per replay below
        const cudaError_t e = cudaStreamEndCapture(cs, &graph);
        if (st != LVKB200_OK || e != cudaSuccess || !graph)
        {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            use_graphs = false;  // fall back to eager launches for the rest of this stream's life
            return record_estimator_chain(false);
        }
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess)
        {
            exec = nullptr;
            cudaGetLastError();
            use_graphs = false;
            return record_estimator_chain(false);
        }
    }
    LVKB_CUDA(cudaGraphLaunch(exec, cs));
    count_launches(graph_kernels);  // compact, score (+ hypotheses), refine
    return LVKB200_OK;
}

// Waits for the tracking chain (its results already sit in pinned host memory: the copy is part of the chain).
lvkb200_status lvkb200_stream::fetch_tracking(int n, bool with_model, std::vector<float>& matched,
                                              std::vector<uint8_t>& status, RansacResult* model,
                                              std::vector<uint8_t>& mask)
{
    matched.resize(static_cast<size_t>(n) * 2);
    status.resize(n);
    if (n == 0) return LVKB200_OK;
    // the chain's last kernel has written the results into mapped pinned memory; work queued behind it on cs (the next
    // frame's pre_ingest) is not waited for
    LVKB_CUDA(track_done ? cudaEventSynchronize(track_done) : cudaStreamSynchronize(cs));
    const uint8_t* base = h_track_out.as<uint8_t>();
    std::memcpy(matched.data(), base, sizeof(float) * matched.size());
    std::memcpy(status.data(), base + off_status, n);
    if (with_model)
    {
        std::memcpy(model, base + off_result, sizeof(RansacResult));
        const int m = std::min(std::max(model->n, 0), n);
        mask.assign(base + off_mask, base + off_mask + m);
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::run_homography(const std::vector<float>& tracked, const std::vector<float>& matched,
                                              float threshold, int model, double h[9], std::vector<uint8_t>& mask,
                                              bool* found)
{
    const int n = static_cast<int>(tracked.size() / 2);
    LVKB_REQUIRE(n >= 4);  // FrameTracker.cpp:335
    LVKB_TRY(ensure_points(n));
    std::memcpy(h_src.ptr, tracked.data(), sizeof(float) * tracked.size());
    std::memcpy(h_dst.ptr, matched.data(), sizeof(float) * matched.size());
    *h_count.as<int>() = n;
    TrackParams* hp = h_params.as<TrackParams>();
    hp->n = n;
    hp->model = model;
    hp->lk_epsilon_sq = 0.0;
    hp->threshold_sq = threshold * threshold;
    LVKB_CUDA(cudaMemcpyAsync(d_params.ptr, h_params.ptr, sizeof(TrackParams), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_src.ptr, h_src.ptr, sizeof(float2) * n, cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_dst.ptr, h_dst.ptr, sizeof(float2) * n, cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_count.ptr, h_count.ptr, sizeof(int), cudaMemcpyHostToDevice, cs));
    LVKB_TRY(ransac_homography(cs, d_src.as<float2>(), d_dst.as<float2>(), d_count.as<int>(),
                               d_params.as<TrackParams>(), d_models.as<float>(), d_scores.as<float>(), d_result(),
                               d_mask()));
    LVKB_CUDA(cudaMemcpyAsync(h_track_out.ptr, d_track_out.ptr, track_out_bytes, cudaMemcpyDeviceToHost, cs));
    LVKB_CUDA(cudaStreamSynchronize(cs));
    const uint8_t* base = h_track_out.as<uint8_t>();
    RansacResult r;
    std::memcpy(&r, base + off_result, sizeof(r));
    *found = r.found != 0;
    mask.assign(base + off_mask, base + off_mask + n);
    std::memcpy(h, r.h, sizeof(double) * 9);
    return LVKB200_OK;
}

// Stage-level entry (parity tests): estimate_local_motions on caller-supplied correspondences and mesh state.
lvkb200_status lvkb200_stream::run_local_motions(const std::vector<float>& tracked, const std::vector<float>& matched,
                                                 float* mesh_state, Mesh& offsets, std::vector<uint8_t>& mask)
{
    const int n = static_cast<int>(tracked.size() / 2);
    const size_t elems = static_cast<size_t>(mesh_solver.unknowns());
    if (mesh_on_device())
    {
        LVKB_TRY(ensure_points(n));
        LVKB_TRY(prepare_mesh_device());
    }
    if (!mesh_on_device())  // small mesh, or one that does not fit a CTA: host solver on a scratch copy of the settings
    {
        MeshSolver solver;
        solver.configure(settings);
        std::memcpy(solver.state().data(), mesh_state, sizeof(float) * elems);
        solver.estimate(tracked, matched, offsets, mask, &last_mesh_iterations);
        std::memcpy(mesh_state, solver.state().data(), sizeof(float) * elems);
        return LVKB200_OK;
    }
    std::memcpy(h_src.ptr, tracked.data(), sizeof(float) * tracked.size());
    std::memcpy(h_dst.ptr, matched.data(), sizeof(float) * matched.size());
    *h_count.as<int>() = n;
    TrackParams* hp = h_params.as<TrackParams>();
    *hp = TrackParams{};
    hp->n = 0;  // no LK results to deliver
    LVKB_CUDA(mesh_device.set_state(cs, mesh_state));
    LVKB_CUDA(cudaMemcpyAsync(d_params.ptr, h_params.ptr, sizeof(TrackParams), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_src.ptr, h_src.ptr, sizeof(float2) * n, cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_dst.ptr, h_dst.ptr, sizeof(float2) * n, cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_count.ptr, h_count.ptr, sizeof(int), cudaMemcpyHostToDevice, cs));
    TrackOutCopy out{};
    out.dev = d_track_out.as<uint8_t>();
    out.host = h_track_out.device_view<uint8_t>();
    out.off_status = static_cast<uint32_t>(off_status);
    out.off_mask = static_cast<uint32_t>(off_mask);
    out.off_result = static_cast<uint32_t>(off_result);
    MeshSolveParams mp{};
    mp.temporal_weight = mesh_solver.temporal_weight();
    mp.acceptance = mesh_solver.acceptance_threshold();
    mp.key_w = mesh_solver.key_w();
    mp.key_h = mesh_solver.key_h();
    mp.min_samples = 0;
    LVKB_CUDA(mesh_device.launch(cs, mp, d_src.as<float2>(), d_dst.as<float2>(), d_count.as<int>(),
                                 d_params.as<TrackParams>(), d_mask(), out));
    LVKB_CUDA(cudaStreamSynchronize(cs));
    const MeshSolveResult& r = mesh_device.result();
    LVKB_REQUIRE(r.solved == 1 && r.n == n);
    last_mesh_iterations = r.iterations;
    std::memcpy(mesh_state, mesh_device.mesh(), sizeof(float) * elems);
    const uint8_t* m = h_track_out.as<uint8_t>() + off_mask;
    mask.assign(m, m + n);
    // offsets from the solution, without disturbing the stream's own tracker state
    std::vector<float> keep = mesh_solver.state();
    mesh_solver.adopt_state(mesh_state);
    mesh_solver.offsets_from_state(offsets);
    mesh_solver.adopt_state(keep.data());
    LVKB_CUDA(mesh_device.set_state(cs, keep.data()));
    return LVKB200_OK;
}

// lvk::fast_erase on parallel arrays (Functions/Container.tpp:31-39)
template <typename T>
static inline void fast_erase_n(std::vector<T>& v, size_t index, size_t width)
{
    const size_t last = v.size() / width - 1;
    for (size_t k = 0; k < width; k++) std::swap(v[index * width + k], v[last * width + k]);
    v.resize(last * width);
}

lvkb200_status lvkb200_stream::track(const QueuedFrame& frame, Mesh& motion, bool* has_motion)
{
    *has_motion = false;
    tracking_stability = 0.0f;  // FrameTracker.cpp:113

    // ---- advance time and import the next frame (FrameTracker.cpp:116-117; gray view StabilizationFilter.cpp:98)
    cur = next_pyr();
    det_pitch = align_up(static_cast<size_t>(det_w), 16);
    LVKB_CUDA(d_det.ensure(det_pitch * det_h));
    LVKB_TRY(ingest.prepare(frame.w, frame.h, det_w, det_h, cs));
    LVKB_TRY(fast.prepare(det_w, det_h));
    for (auto& py : pyr) LVKB_TRY(py.prepare(det_w, det_h, cs));

    // the previous submit may already have built this frame's detection image and pyramid (pre_ingest)
    const bool prebuilt = frame_prebuilt && pyr[cur].valid;
    frame_prebuilt = false;
    stage_begin(ST_INGEST);
    if (!prebuilt)
        LVKB_TRY(ingest.launch(cs, frame.buf.as<uint8_t>(), frame.pitch, frame.format, d_det.as<uint8_t>(), det_pitch));
    stage_end(ST_INGEST);
    // FAST reads the detection image only, the pyramid is needed by LK only: FAST goes first and the pyramid is
    // queued behind it, so it is built while the host waits for and digests the FAST keypoints.
    const bool can_track = frame_initialized && pyr[prev_pyr()].valid;
    // per-frame scratch lives in the stream (capacity is kept): a steady-state frame allocates nothing on the host
    std::vector<FastRegion>& regions = scratch.regions;
    std::vector<int>& region_index = scratch.region_index;
    std::vector<std::vector<FastPoint>>& fast_points = scratch.fast_points;
    regions.clear();
    region_index.clear();
    if (can_track)
    {
        grid.plan_detection(regions, region_index);  // FeatureDetector::detect (FrameTracker.cpp:127)
        stage_begin(ST_FAST);
        if (!regions.empty())
            LVKB_TRY(fast.launch(cs, d_det.as<uint8_t>(), det_pitch, regions.data(), static_cast<int>(regions.size())));
        stage_end(ST_FAST);
    }
    stage_begin(ST_PYRAMID);
    if (!prebuilt) LVKB_TRY(pyr[cur].build(cs, d_det.as<uint8_t>(), det_pitch));
    stage_end(ST_PYRAMID);
    host_tick(HP_ENQ_DETECT);

    if (debug_capture)
    {
        LVKB_CUDA(h_det.ensure(static_cast<size_t>(det_w) * det_h));
        LVKB_CUDA(cudaMemcpy2DAsync(h_det.ptr, det_w, d_det.ptr, det_pitch, det_w, det_h, cudaMemcpyDeviceToHost, cs));
        LVKB_CUDA(cudaStreamSynchronize(cs));
        dbg_det.assign(h_det.as<uint8_t>(), h_det.as<uint8_t>() + static_cast<size_t>(det_w) * det_h);
    }

    // ---- we need at least two frames (FrameTracker.cpp:120-124)
    if (!can_track)
    {
        frame_initialized = true;
        return LVKB200_OK;
    }

    if (!regions.empty()) LVKB_TRY(fast.fetch(fast_points));
    host_tick(HP_WAIT_FAST);
    const float distribution = grid.finish_detection(region_index, fast_points, features, dbg_fast_counts);
    if (debug_capture)
    {
        dbg_detected.resize(features.size());
        for (size_t i = 0; i < features.size(); i++)
            dbg_detected[i] = {features[i].x, features[i].y, features[i].response, features[i].class_id};
    }
    if (features.size() < settings.min_motion_samples || distribution < settings.uniformity_threshold)
    {
        features.clear();
        return LVKB200_OK;
    }

    // ---- sparse optical flow (FrameTracker.cpp:135-146)
    std::vector<float>&tracked = scratch.tracked, &matched = scratch.matched;
    std::vector<uint8_t>& status = scratch.status;
    tracked.resize(features.size() * 2);
    matched.clear();
    status.clear();
    for (size_t i = 0; i < features.size(); i++)
    {
        tracked[2 * i] = features[i].x;
        tracked[2 * i + 1] = features[i].y;
    }
    const int n_tracked = static_cast<int>(features.size());
    const bool global = !settings.track_local_motions;
    // FrameTracker.cpp:167-176: homography when the features are well distributed, partial affine otherwise
    const int model_kind = (distribution > HOMOGRAPHY_DISTRIBUTION_THRESHOLD) ? 0 : 1;
    host_tick(HP_DETECT);
    // 0: behind the chain, 1: beside the whole chain, 2: beside everything after LK
    const int overlap = remap_overlap >= 0 ? remap_overlap
                                           : (static_cast<long long>(frame.w) * frame.h >= REMAP_OVERLAP_MIN_PIXELS ? 1 : 0);
    overlap_after_lk = overlap == 2;
    if (overlap == 1 && pending.active && cs_remap)
    {
        // everything the held-back remap depends on (its parked source frame was uploaded >= frame_delay submits ago)
        // precedes this point of cs: the remap may start now, beside this frame's LK + RANSAC
        if (!pre_chain) LVKB_CUDA(cudaEventCreateWithFlags(&pre_chain, cudaEventDisableTiming));
        LVKB_CUDA(cudaEventRecord(pre_chain, cs));
        pre_chain_valid = true;
    }
    LVKB_TRY(enqueue_tracking(tracked, global, settings.acceptance_threshold, model_kind));
    if (!track_done) LVKB_CUDA(cudaEventCreateWithFlags(&track_done, cudaEventDisableTiming));
    LVKB_CUDA(cudaEventRecord(track_done, cs));
    LVKB_TRY(flush_remap());  // the previous output's remap runs beside LK + RANSAC (see apply_mesh)
    LVKB_TRY(pre_ingest());   // the announced next frame's detection image + pyramid, behind this frame's chain
    host_tick(HP_ENQ_TRACK);
    RansacResult model{};
    std::vector<uint8_t>& inliers = scratch.inliers;
    inliers.clear();
    LVKB_TRY(fetch_tracking(n_tracked, global, matched, status, &model, inliers));
    host_tick(HP_WAIT_TRACK);
    if (debug_capture)
    {
        dbg_lk_matched = matched;
        dbg_lk_status = status;
    }

    // ---- fast_filter(features, tracked, matched, status) (FrameTracker.cpp:149, Container.tpp:97-121)
    for (int k = static_cast<int>(status.size()) - 1; k >= 0; k--)
    {
        if (!status[k])
        {
            std::swap(features[k], features.back());
            features.pop_back();
            fast_erase_n(tracked, k, 2);
            fast_erase_n(matched, k, 2);
        }
    }
    if (matched.size() / 2 < settings.min_motion_samples)
    {
        features.clear();
        return LVKB200_OK;
    }
    if (debug_capture)
    {
        dbg_tracked = tracked;
        dbg_matched = matched;
    }

    // ---- motion estimation (FrameTracker.cpp:157-176)
    if (settings.track_local_motions && mesh_on_device())
    {
        // solved by k_mesh_cgls behind LK: the mesh and the mask are already in pinned host memory
        const MeshSolveResult& r = mesh_device.result();
        const size_t n_est = matched.size() / 2;
        // the device compaction must have produced exactly the host's survivor list
        LVKB_REQUIRE(r.solved == 1 && r.n == static_cast<int>(n_est));
        const uint8_t* m = h_track_out.as<uint8_t>() + off_mask;
        inliers.assign(m, m + n_est);
        mesh_solver.adopt_state(mesh_device.mesh());
        mesh_solver.offsets_from_state(motion);
        last_mesh_iterations = r.iterations;
    }
    else if (settings.track_local_motions)
    {
        stage_begin(ST_ESTIMATE);
        mesh_solver.estimate(tracked, matched, motion, inliers, &last_mesh_iterations);
        stage_end(ST_ESTIMATE);
    }
    else
    {
        // the device compaction must have produced exactly the host's survivor list
        LVKB_REQUIRE(model.n == static_cast<int>(matched.size() / 2) && inliers.size() == matched.size() / 2);
        if (!model.found)
        {
            // cv::findHomography returned an empty matrix: the reference asserts (Math/Homography.cpp:89-95).
            // Reported through the assert handler; the frame is then treated as "no motion".
            report_assert(__FILE__, "estimate_global_motion", "homography estimation found no model");
            features.clear();
            return LVKB200_OK;
        }
        std::memcpy(dbg_h, model.h, sizeof(dbg_h));
        dbg_has_h = true;
        mesh_set_to_homography(model.h, static_cast<float>(det_w), static_cast<float>(det_h),
                               settings.motion_resolution_width, settings.motion_resolution_height, motion);
    }
    if (debug_capture) dbg_inliers = inliers;

    // ---- tracking stability = inlier ratio (FrameTracker.cpp:179, Container.tpp:125-129)
    size_t inl = 0;
    for (uint8_t v : inliers) inl += (v == 1);
    tracking_stability = static_cast<float>(inl) / static_cast<float>(inliers.size());

    // ---- drop outliers, age inliers, propagate (FrameTracker.cpp:183-193)
    for (int i = static_cast<int>(inliers.size()) - 1; i >= 0; i--)
    {
        if (inliers[i])
        {
            features[i].class_id++;
            features[i].x = matched[2 * i];
            features[i].y = matched[2 * i + 1];
        }
        else
        {
            std::swap(features[i], features.back());
            features.pop_back();
        }
    }
    grid.propagate(features);
    if (debug_capture)
    {
        dbg_propagated.resize(features.size());
        for (size_t i = 0; i < features.size(); i++)
            dbg_propagated[i] = {features[i].x, features[i].y, features[i].response, features[i].class_id};
    }
    *has_motion = true;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::apply_mesh(QueuedFrame& src, const Mesh& offsets, void* out, size_t out_pitch,
                                          lvkb200_memspace out_space)
{
    // WarpMesh::apply — Math/WarpMesh.cpp:183-223
    LVKB_TRY(ensure_pipeline());
    LVKB_TRY(flush_remap());  // at most one remap is held back
    PendingRemap& pr = pending;
    RemapParams& p = pr.p;
    p = RemapParams{};
    p.src = src.buf.as<uint8_t>();
    p.src_pitch = src.pitch;
    p.width = src.w;
    p.height = src.h;
    p.yuv = src.format == LVKB200_YUV;  // Image.cpp:100
    for (int k = 0; k < 3; k++) p.bg[k] = static_cast<uint8_t>(settings.background_colour[k]);  // Image.cpp:136-141
    pr.async_host_out = deferred_output && out_space == LVKB200_MEM_HOST;
    pr.egress = pr.async_host_out && next_egress;
    if (pr.egress) pr.egress_frame = next_egress_frame;
    pr.slot = 0;
    pr.out = out;
    pr.out_pitch = out_pitch;
    pr.out_space = out_space;
    if (pr.async_host_out)
    {
        // pipelined output: remap into one of two device staging buffers; the copy-out stream downloads it while
        // the next frames are being processed (lvkb200_stream_wait_output waits for that download)
        last_ticket = ++async_tickets;
        pr.ticket = last_ticket;
        pr.slot = static_cast<int>(last_ticket & 1);
        p.dst_pitch = align_up(static_cast<size_t>(src.w) * 3, 16);
        LVKB_CUDA(async_out[pr.slot].ensure(p.dst_pitch * src.h));
        p.dst = async_out[pr.slot].as<uint8_t>();
    }
    else
        LVKB_TRY(stage_frame_out(out, out_pitch, src.w, src.h, 3, out_space, &p.dst, &p.dst_pitch));
    pr.homography = settings.motion_resolution_width == 2 && settings.motion_resolution_height == 2;
    pr.dmesh = nullptr;
    if (pr.homography)
    {
        double t[9];
        LVKB_REQUIRE(mesh2x2_to_transform(offsets.data(), src.w, src.h, t));
        std::memcpy(dbg_t, t, sizeof(t));
        dbg_has_t = true;
        for (int k = 0; k < 9; k++) pr.tf[k] = static_cast<float>(t[k]);
    }
    else
    {
        // the mesh staging buffers are shared with the previous remap: order this upload (on cs) behind it
        LVKB_TRY(join_remap(cs));
        LVKB_TRY(upload_mesh(offsets.data(), settings.motion_resolution_width, settings.motion_resolution_height, &pr.dmesh));
    }
    pr.active = true;
    std::swap(src.buf, spare_buf);  // park the buffer this remap reads; the slot gets the previously parked one
    // Hold the launch back when nobody waits for the pixels inside this call: the remap fills every SM for ~50 us, and
    // queued right now it would share them with the NEXT frame's ingest/pyramid/FAST (also issue-bound: both slow down).
    // Launched behind the next frame's LK + RANSAC graph instead, it runs while those few latency-bound warps leave
    // the machine idle.  Flushed by the next submit, lvkb200_stream_sync/_event_record/_wait_output.
    const bool hold = pr.homography && (out_space == LVKB200_MEM_DEVICE || pr.async_host_out) && !profile_stages &&
                      !debug_capture;
    if (hold) return LVKB200_OK;
    LVKB_TRY(flush_remap());
    if (out_space == LVKB200_MEM_DEVICE || pr.async_host_out) return LVKB200_OK;
    LVKB_TRY(join_remap(cs));
    return finish_frame_out(out, out_pitch, p.width, p.height, 3, out_space);
}

// Launches the held-back remap (if any) on cs_remap, and the download of its result when the output is host memory
// of a pipelined submit.
lvkb200_status lvkb200_stream::flush_remap()
{
    if (!pending.active) return LVKB200_OK;
    PendingRemap& pr = pending;
    pr.active = false;
    const RemapParams& p = pr.p;
    if (pr.async_host_out && async_out_used[pr.slot])
        LVKB_CUDA(cudaStreamWaitEvent(cs_remap, async_out_done[pr.slot], 0));  // the staging buffer's last download
    if (pre_chain_valid && pr.homography)
    {
        LVKB_CUDA(cudaStreamWaitEvent(cs_remap, pre_chain, 0));
    }
    else
    {
        // everything queued on cs so far (the source frame's upload, the mesh upload) precedes the remap
        LVKB_CUDA(cudaEventRecord(chain_point, cs));
        LVKB_CUDA(cudaStreamWaitEvent(cs_remap, chain_point, 0));
    }
    pre_chain_valid = false;
    stage_begin(ST_REMAP, cs_remap);
    if (pr.homography)
        LVKB_CUDA(launch_remap_homography(cs_remap, p, pr.tf));
    else
        LVKB_CUDA(launch_remap_mesh(cs_remap, p, pr.dmesh, settings.motion_resolution_width, settings.motion_resolution_height));
    stage_end(ST_REMAP, cs_remap);
    LVKB_CUDA(cudaEventRecord(remap_done[remaps_launched & 1], cs_remap));
    remaps_launched++;
    if (pr.async_host_out)
    {
        LVKB_CUDA(cudaEventRecord(async_remap_done[pr.slot], cs_remap));
        LVKB_CUDA(cudaStreamWaitEvent(cs_out, async_remap_done[pr.slot], 0));
        if (pr.egress)
            LVKB_TRY(egress_planes(cs_out, p.dst, p.dst_pitch, p.width, p.height, p.yuv ? LVKB200_YUV : LVKB200_BGR,
                                   pr.egress_frame, pr.slot));
        else
            LVKB_CUDA(cudaMemcpy2DAsync(pr.out, pr.out_pitch, p.dst, p.dst_pitch, static_cast<size_t>(p.width) * 3, p.height,
                                        cudaMemcpyDeviceToHost, cs_out));
        LVKB_CUDA(cudaEventRecord(async_out_done[pr.slot], cs_out));
        async_out_used[pr.slot] = true;
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::ensure_pipeline()
{
    if (cs_in) return LVKB200_OK;
    LVKB_CUDA(cudaStreamCreateWithFlags(&cs_in, cudaStreamNonBlocking));
    LVKB_CUDA(cudaStreamCreateWithFlags(&cs_out, cudaStreamNonBlocking));
    for (auto& e : prefetch_done) LVKB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    LVKB_CUDA(cudaStreamCreateWithFlags(&cs_remap, cudaStreamNonBlocking));
    LVKB_CUDA(cudaEventCreateWithFlags(&chain_point, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++)
    {
        LVKB_CUDA(cudaEventCreateWithFlags(&remap_done[k], cudaEventDisableTiming));
        LVKB_CUDA(cudaEventCreateWithFlags(&async_remap_done[k], cudaEventDisableTiming));
        LVKB_CUDA(cudaEventCreateWithFlags(&async_out_done[k], cudaEventDisableTiming));
    }
    return LVKB200_OK;
}

// Starts the upload (host memory) or the copy (device memory) of the NEXT input frame on the copy-in stream; the
// following submit of the same pointer adopts the buffer instead of copying.  With a known format the frame is also
// eligible for pre_ingest.
lvkb200_status lvkb200_stream::prefetch(const void* frame, size_t pitch, int width, int height, lvkb200_format format,
                                        lvkb200_memspace space)
{
    LVKB_REQUIRE(frame != nullptr && width > 0 && height > 0);
    const size_t row = static_cast<size_t>(width) * 3;
    LVKB_REQUIRE(pitch >= row);
    LVKB_TRY(ensure_pipeline());
    LVKB_TRY(ensure_frame_pool(align_up(row, 16) * height));
    const int k = prefetch_next;
    prefetch_next ^= 1;
    QueuedFrame& ps = prefetch_slot[k];
    ps.pitch = align_up(row, 16);
    ps.w = width;
    ps.h = height;
    if (ps.buf.capacity < ps.pitch * height)
    {
        LVKB_TRY(sync_all());  // the buffer being replaced may still be read by a queued remap
        LVKB_CUDA(ps.buf.ensure(ps.pitch * height));
    }
    // this buffer was a ring buffer until the last swap: wait for the last remap that could have read it
    LVKB_TRY(wait_frame_buffers_free(cs_in));
    LVKB_CUDA(cudaMemcpy2DAsync(ps.buf.ptr, ps.pitch, frame, pitch, row, height,
                                space == LVKB200_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs_in));
    LVKB_CUDA(cudaEventRecord(prefetch_done[k], cs_in));
    prefetched_ptr[k] = frame;
    lookahead[k] = Lookahead{};
    lookahead[k].announced = format == LVKB200_BGR || format == LVKB200_RGB || format == LVKB200_YUV;
    lookahead[k].format = format;
    return LVKB200_OK;
}

// Builds the detection image and the pyramid of the announced NEXT frame on cs, right behind the current frame's
// tracking chain (whose completion the host waits for through track_done, not through the stream): the ~20 us of
// ingest + pyramid run while the host digests the chain's results instead of at the head of the next submit.
// Everything here only moves work earlier on the same in-order stream; the next submit's results are unchanged.
lvkb200_status lvkb200_stream::pre_ingest()
{
    if (profile_stages || debug_capture || !settings.stabilize_output) return LVKB200_OK;
    for (int k = 0; k < 2; k++)
    {
        Lookahead& la = lookahead[k];
        QueuedFrame& ps = prefetch_slot[k];
        if (!la.announced || la.built || prefetched_ptr[k] == nullptr) continue;
        const QueuedFrame& now = ring[(ring_start + ring_size - 1) % ring.size()];
        if (ps.w != now.w || ps.h != now.h) continue;  // geometry change: the ordinary path re-plans the ingest
        LVKB_CUDA(cudaStreamWaitEvent(cs, prefetch_done[k], 0));
        if (deblock_enabled)
        {
            // CompositeFilter chain: the tracker sees the deblocked frame
            LVKB_TRY(deblock.prepare(ps.w, ps.h, deblock_settings, cs));
            LVKB_TRY(deblock.launch(cs, ps.buf.as<uint8_t>(), ps.pitch, la.format));
            la.deblocked = true;
        }
        LVKB_TRY(ingest.launch(cs, ps.buf.as<uint8_t>(), ps.pitch, la.format, d_det.as<uint8_t>(), det_pitch));
        la.pyr_index = next_pyr();
        LVKB_TRY(pyr[la.pyr_index].build(cs, d_det.as<uint8_t>(), det_pitch));
        la.built = true;
        break;
    }
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::wait_output(uint64_t ticket)
{
    if (ticket == 0 || !cs_out) return LVKB200_OK;
    LVKB_REQUIRE(ticket <= async_tickets);
    if (pending.active && pending.async_host_out && pending.ticket <= ticket) LVKB_TRY(flush_remap());  // still held back
    // waits for the most recent download recorded on that staging slot (>= the ticket's own download)
    LVKB_CUDA(cudaEventSynchronize(async_out_done[ticket & 1]));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::submit(const void* frame, size_t pitch, int width, int height, lvkb200_format format,
                                      uint64_t timestamp, lvkb200_memspace frame_space, void* out, size_t out_pitch,
                                      lvkb200_memspace out_space, lvkb200_result* res)
{
    std::memset(res, 0, sizeof(*res));
    res->out_format = LVKB200_UNKNOWN;
    LVKB_REQUIRE(format != LVKB200_UNKNOWN);  // StabilizationFilter.cpp:71
    LVKB_REQUIRE(width > 0 && height > 0);    // :72
    // Only packed 3-channel frames reach lvk::remap (Image.cpp:32,96).
    LVKB_REQUIRE(format == LVKB200_BGR || format == LVKB200_RGB || format == LVKB200_YUV);
    const size_t row = static_cast<size_t>(width) * 3;
    LVKB_REQUIRE(pitch >= row);
    if (host_trace)
    {
        host_frames++;
        host_tick(HP_OUTSIDE);  // time since the previous submit returned (the caller's own work, prefetch, wait_output)
    }
    stage_parity ^= 1;
    harvest_stage_times(stage_parity);  // this slot holds the events of two submits ago: long complete
    dbg_has_h = dbg_has_t = false;
    dbg_detected.clear(); dbg_propagated.clear(); dbg_lk_matched.clear(); dbg_lk_status.clear(); dbg_tracked.clear();
    dbg_matched.clear(); dbg_inliers.clear(); dbg_motion.clear(); dbg_correction.clear(); dbg_fast_counts.clear();

    // ---- m_FrameQueue.push(std::move(input)): the frame becomes resident in the device ring
    LVKB_TRY(ensure_frame_pool(align_up(row, 16) * height));
    const size_t cap = ring.size();
    size_t slot;
    if (ring_size == cap)
    {
        slot = ring_start;  // StreamBuffer::push on a full ring overwrites the oldest (StreamBuffer.tpp:37-84)
        ring_start = (ring_start + 1) % cap;
    }
    else
    {
        slot = (ring_start + ring_size) % cap;
        ring_size++;
    }
    QueuedFrame& q = ring[slot];
    int pk = -1;
    for (int k = 0; k < 2; k++)
        if (prefetched_ptr[k] == frame && prefetch_slot[k].w == width && prefetch_slot[k].h == height &&
            prefetch_slot[k].pitch == align_up(row, 16))
            pk = k;
    const bool prefetched = pk >= 0;
    bool deblocked = false;
    frame_prebuilt = false;
    if (prefetched)
    {
        // lvkb200_stream_prefetch already uploaded this frame into a spare buffer on the copy-in stream:
        // adopt that buffer (O(1) swap) and make this stream wait for the upload instead of copying again.
        std::swap(q.buf, prefetch_slot[pk].buf);
        q.pitch = prefetch_slot[pk].pitch;
        LVKB_CUDA(cudaStreamWaitEvent(cs, prefetch_done[pk], 0));
        prefetched_ptr[pk] = nullptr;
        // look-ahead work of the previous submit (pre_ingest): only valid for the format it was announced with and
        // for the pyramid slot this frame is about to take
        const Lookahead& la = lookahead[pk];
        const bool usable = la.announced && la.format == format;
        deblocked = usable && la.deblocked;
        frame_prebuilt = usable && la.built && la.pyr_index == next_pyr() && la.deblocked == deblock_enabled;
        LVKB_REQUIRE(!la.deblocked || usable);  // the slot's pixels were deblocked for another format: caller error
        lookahead[pk] = Lookahead{};
    }
    else
    {
        q.pitch = align_up(row, 16);
        LVKB_CUDA(q.buf.ensure(q.pitch * height));
        LVKB_TRY(wait_frame_buffers_free(cs));  // the remaps run on their own stream
        LVKB_CUDA(cudaMemcpy2DAsync(q.buf.ptr, q.pitch, frame, pitch, row, height,
                                    frame_space == LVKB200_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs));
    }
    q.w = width; q.h = height; q.format = format; q.timestamp = timestamp;
    if (deblock_enabled && !deblocked)
    {
        // CompositeFilter{Deblocking, Stabilization} (CompositeFilter.cpp:58-88): the first filter's output is the
        // second one's input — here the frame never leaves its ring slot
        LVKB_TRY(deblock.prepare(width, height, deblock_settings, cs));
        LVKB_TRY(deblock.launch(cs, q.buf.as<uint8_t>(), q.pitch, format));
    }
    // The caller keeps ownership of its buffer: host memory must have been consumed before we return.  A prefetched
    // host frame was uploaded on the copy-in stream: its upload event is waited for instead (almost always complete
    // already; without it the paths that never wait on `cs` could return while the copy is still reading the buffer).
    struct InputGuard
    {
        lvkb200_stream* s;
        bool host;
        cudaEvent_t upload;
        ~InputGuard()
        {
            if (host && s->input_copied) cudaEventSynchronize(s->input_copied);
            if (upload) cudaEventSynchronize(upload);
        }
    } input_guard{this, frame_space == LVKB200_MEM_HOST && !prefetched,
                  (frame_space == LVKB200_MEM_HOST && prefetched) ? prefetch_done[pk] : nullptr};
    if (frame_space == LVKB200_MEM_HOST && !prefetched)
    {
        if (!input_copied) LVKB_CUDA(cudaEventCreateWithFlags(&input_copied, cudaEventDisableTiming));
        LVKB_CUDA(cudaEventRecord(input_copied, cs));
    }

    if (!settings.stabilize_output)
    {
        // StabilizationFilter.cpp:77-95: only up-keep the delay
        LVKB_TRY(flush_remap());
        if (ready())
        {
            QueuedFrame& oldest = ring[ring_start];
            ring_start = (ring_start + 1) % cap;
            ring_size--;
            LVKB_REQUIRE(out != nullptr);
            if (settings.crop_to_stable_region)
            {
                const uint8_t black[3] = {0, 0, 0};  // WarpMesh::apply default background
                lvkb200_settings saved = settings;
                for (int k = 0; k < 3; k++) settings.background_colour[k] = black[k];
                const lvkb200_status st = apply_mesh(oldest, smoother.scene_crop, out, out_pitch, out_space);
                settings = saved;
                LVKB_TRY(st);
            }
            else
            {
                // the pass-through copy follows the same protocol as a remap (own stream, parked source buffer)
                LVKB_TRY(ensure_pipeline());
                LVKB_CUDA(cudaEventRecord(chain_point, cs));
                LVKB_CUDA(cudaStreamWaitEvent(cs_remap, chain_point, 0));
                LVKB_CUDA(cudaMemcpy2DAsync(out, out_pitch, oldest.buf.ptr, oldest.pitch, static_cast<size_t>(oldest.w) * 3,
                                            oldest.h,
                                            out_space == LVKB200_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                            cs_remap));
                LVKB_CUDA(cudaEventRecord(remap_done[remaps_launched & 1], cs_remap));
                remaps_launched++;
                std::swap(oldest.buf, spare_buf);
                if (out_space == LVKB200_MEM_HOST) LVKB_CUDA(cudaStreamSynchronize(cs_remap));
            }
            res->has_output = 1;
            res->out_timestamp = oldest.timestamp;
            res->out_format = oldest.format;
        }
        res->scene_quality = scene_quality;
        res->trust_factor = trust_factor;
        return LVKB200_OK;
    }

    // ---- track the motion of the incoming frame (StabilizationFilter.cpp:98-99)
    Mesh motion;
    bool has_motion = false;
    host_tick(HP_UPLOAD);
    LVKB_TRY(track(q, motion, &has_motion));
    LVKB_TRY(flush_remap());  // (frames on which no tracking chain ran)
    host_tick(HP_POST);
    const size_t elems = static_cast<size_t>(settings.motion_resolution_width) * settings.motion_resolution_height * 2;
    if (!has_motion) motion.assign(elems, 0.0f);  // m_NullMotion
    if (debug_capture) dbg_motion = motion;

    // ---- quality assurance (StabilizationFilter.cpp:101-115)
    const float tracking_quality = tracking_stability;
    scene_quality = scene_quality + QA_UPDATE_RATE * (tracking_quality - scene_quality);
    if (tracking_quality < settings.min_tracking_quality)
        trust_factor = 0.0f;
    else if (scene_quality < settings.min_scene_quality)
        trust_factor = step_to<float>(trust_factor, 0.0f, QA_BLEND_STEP);
    else
        trust_factor = step_to<float>(trust_factor, 1.0f, QA_BLEND_STEP);
    for (float& v : motion) v *= trust_factor;

    // ---- path smoothing (StabilizationFilter.cpp:121)
    Mesh correction;
    smoother.next(motion, correction);
    host_tick(HP_SMOOTH);

    if (ready())
    {
        QueuedFrame& next_frame = ring[ring_start];
        ring_start = (ring_start + 1) % cap;
        ring_size--;
        if (settings.crop_to_stable_region)
            for (size_t k = 0; k < elems; k++) correction[k] += smoother.scene_crop[k];
        if (debug_capture) dbg_correction = correction;
        LVKB_REQUIRE(out != nullptr);
        LVKB_TRY(apply_mesh(next_frame, correction, out, out_pitch, out_space));
        host_tick(HP_REMAP);
        res->has_output = 1;
        res->out_timestamp = next_frame.timestamp;
        res->out_format = next_frame.format;
    }
    else if (debug_capture)
        dbg_correction = correction;

    res->tracking_stability = tracking_stability;
    res->scene_quality = scene_quality;
    res->trust_factor = trust_factor;
    res->feature_count = static_cast<int32_t>(features.size());
    res->has_motion = has_motion ? 1 : 0;
    return LVKB200_OK;
}

template <typename T>
static lvkb200_status copy_out(const T* data, size_t count, void* buffer, size_t capacity, size_t* size)
{
    *size = count * sizeof(T);
    if (buffer && capacity > 0) std::memcpy(buffer, data, std::min(capacity, *size));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::debug_fetch(lvkb200_debug_item which, void* buffer, size_t capacity, size_t* size)
{
    *size = 0;
    switch (which)
    {
        case LVKB200_DBG_DETECTION_IMAGE: return copy_out(dbg_det.data(), dbg_det.size(), buffer, capacity, size);
        case LVKB200_DBG_DETECTED: return copy_out(dbg_detected.data(), dbg_detected.size(), buffer, capacity, size);
        case LVKB200_DBG_LK_MATCHED: return copy_out(dbg_lk_matched.data(), dbg_lk_matched.size(), buffer, capacity, size);
        case LVKB200_DBG_LK_STATUS: return copy_out(dbg_lk_status.data(), dbg_lk_status.size(), buffer, capacity, size);
        case LVKB200_DBG_TRACKED: return copy_out(dbg_tracked.data(), dbg_tracked.size(), buffer, capacity, size);
        case LVKB200_DBG_MATCHED: return copy_out(dbg_matched.data(), dbg_matched.size(), buffer, capacity, size);
        case LVKB200_DBG_INLIERS: return copy_out(dbg_inliers.data(), dbg_inliers.size(), buffer, capacity, size);
        case LVKB200_DBG_HOMOGRAPHY: return dbg_has_h ? copy_out(dbg_h, 9, buffer, capacity, size) : LVKB200_OK;
        case LVKB200_DBG_MOTION: return copy_out(dbg_motion.data(), dbg_motion.size(), buffer, capacity, size);
        case LVKB200_DBG_CORRECTION: return copy_out(dbg_correction.data(), dbg_correction.size(), buffer, capacity, size);
        case LVKB200_DBG_WARP_TRANSFORM: return dbg_has_t ? copy_out(dbg_t, 9, buffer, capacity, size) : LVKB200_OK;
        case LVKB200_DBG_PROPAGATED: return copy_out(dbg_propagated.data(), dbg_propagated.size(), buffer, capacity, size);
        case LVKB200_DBG_FAST_COUNTS: return copy_out(dbg_fast_counts.data(), dbg_fast_counts.size(), buffer, capacity, size);
        case LVKB200_DBG_MESH_ITERATIONS:
        {
            const int32_t it = last_mesh_iterations;
            return copy_out(&it, 1, buffer, capacity, size);
        }
    }
    return LVKB200_ERR_INVALID;
}

void lvkb200_stream::release()
{
    if (host_trace && host_frames > HOST_TRACE_SKIP)
    {
        static const char* names[HP_COUNT] = {"upload", "enq_detect", "wait_fast", "detect", "enq_track",
                                              "wait_track", "post", "smooth", "remap", "outside"};
        const double frames = static_cast<double>(host_frames - HOST_TRACE_SKIP);
        std::fprintf(stderr, "[lvkb200 host trace] %.0f frames, us/frame:", frames);
        double sum = 0.0;
        for (int i = 0; i < HP_COUNT; i++)
        {
            std::fprintf(stderr, " %s=%.1f", names[i], host_us[i] / frames);
            sum += host_us[i];
        }
        std::fprintf(stderr, " total=%.1f\n", sum / frames);
        host_frames = 0;
    }
    planes_in.release(); planes_out.release(); obs_frame_in.release(); obs_frame_out.release();
    format_plan.xtab.release(); format_plan.ytab.release();
    stage_in.release(); stage_out.release(); scaling_scratch.release(); mesh_dev.release(); mesh_pinned.release(); mesh_device.release();
    ingest.release(); fast.release(); for (auto& py : pyr) py.release(); d_det.release();
    if (track_done) { cudaEventDestroy(track_done); track_done = nullptr; }
    deblock.release(); deblock_stage.release();
    destroy_graphs();
    d_pts_prev.release(); d_src.release(); d_dst.release(); d_models.release(); d_scores.release();
    d_track_out.release(); h_track_out.release(); d_params.release(); h_params.release();
    h_pts_prev.release(); h_src.release(); h_dst.release();
    h_det.release(); h_count.release(); d_perm.release(); d_removed.release(); d_count.release();
    point_capacity = 0;
    for (auto& f : ring) f.buf.release();
    if (input_copied) cudaEventDestroy(input_copied);
    input_copied = nullptr;
    if (cs_in) cudaStreamSynchronize(cs_in);
    if (cs_out) cudaStreamSynchronize(cs_out);
    for (auto& ps : prefetch_slot) ps.buf.release();
    for (int k = 0; k < 2; k++) { planes_in_async[k].release(); planes_out_async[k].release(); }
    for (int k = 0; k < 2; k++)
    {
        async_out[k].release();
        if (async_remap_done[k]) cudaEventDestroy(async_remap_done[k]);
        if (async_out_done[k]) cudaEventDestroy(async_out_done[k]);
        async_remap_done[k] = async_out_done[k] = nullptr;
        async_out_used[k] = false;
    }
    for (auto& e : prefetch_done)
    {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    if (cs_remap) cudaStreamSynchronize(cs_remap);
    spare_buf.release();
    for (auto& e : remap_done)
    {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    remaps_launched = 0;
    pending.active = false;
    if (chain_point) cudaEventDestroy(chain_point);
    chain_point = nullptr;
    if (pre_chain) cudaEventDestroy(pre_chain);
    pre_chain = nullptr;
    pre_chain_valid = false;
    if (cs_in) cudaStreamDestroy(cs_in);
    if (cs_out) cudaStreamDestroy(cs_out);
    if (cs_remap) cudaStreamDestroy(cs_remap);
    cs_in = cs_out = cs_remap = nullptr;
    prefetched_ptr[0] = prefetched_ptr[1] = nullptr;
    for (auto& e : user_events)
    {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    for (auto& slot : stage_ev)
        for (auto& pair : slot)
            for (auto& e : pair)
            {
                if (e) cudaEventDestroy(e);
                e = nullptr;
            }
}

lvkb200_status lvkb200_stream::stage_frame_in(const void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                              const uint8_t** dptr, size_t* dpitch)
{
    const size_t row = static_cast<size_t>(w) * ch;
    LVKB_REQUIRE(pitch >= row);
    if (space == LVKB200_MEM_DEVICE)
    {
        *dptr = static_cast<const uint8_t*>(p);
        *dpitch = pitch;
        return LVKB200_OK;
    }
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(stage_in.ensure(dp * h));
    LVKB_CUDA(cudaMemcpy2DAsync(stage_in.ptr, dp, p, pitch, row, h, cudaMemcpyHostToDevice, cs));
    *dptr = stage_in.as<uint8_t>();
    *dpitch = dp;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::stage_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space,
                                               uint8_t** dptr, size_t* dpitch)
{
    const size_t row = static_cast<size_t>(w) * ch;
    LVKB_REQUIRE(pitch >= row);
    if (space == LVKB200_MEM_DEVICE)
    {
        *dptr = static_cast<uint8_t*>(p);
        *dpitch = pitch;
        return LVKB200_OK;
    }
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(stage_out.ensure(dp * h));
    *dptr = stage_out.as<uint8_t>();
    *dpitch = dp;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::finish_frame_out(void* p, size_t pitch, int w, int h, int ch, lvkb200_memspace space)
{
    if (space == LVKB200_MEM_DEVICE) return LVKB200_OK;
    const size_t row = static_cast<size_t>(w) * ch;
    const size_t dp = align_up(row, 16);
    LVKB_CUDA(cudaMemcpy2DAsync(p, pitch, stage_out.ptr, dp, row, h, cudaMemcpyDeviceToHost, cs));
    LVKB_CUDA(cudaStreamSynchronize(cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream::upload_mesh(const float* offsets, int cols, int rows, const float** dmesh)
{
    const size_t bytes = sizeof(float) * 2 * static_cast<size_t>(cols) * rows;
    LVKB_CUDA(mesh_dev.ensure(bytes));
    LVKB_CUDA(mesh_pinned.ensure(bytes));
    // the pinned staging copy may still be in flight from a previous call on this stream
    LVKB_CUDA(cudaStreamSynchronize(cs));
    std::memcpy(mesh_pinned.ptr, offsets, bytes);
    LVKB_CUDA(cudaMemcpyAsync(mesh_dev.ptr, mesh_pinned.ptr, bytes, cudaMemcpyHostToDevice, cs));
    *dmesh = mesh_dev.as<float>();
    return LVKB200_OK;
}
