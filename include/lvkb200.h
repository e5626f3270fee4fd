/*
 * lvkb200 — C-ABI of the B200-native LiveVisionKit stabilization path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/OpenCV types.  Each entry
 * point cites the reference interface it replaces (paths relative to /root/reference/LiveVisionKit).
 * The C++ mirror of the reference API (lvk::VideoFilter / lvk::StabilizationFilter, same names and
 * settings structs) that forwards to these calls lives in livevisionkit_b200/compat/ ;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - All functions return lvkb200_status; nothing throws across the boundary.  The compat layer
 *     converts non-OK into lvk::context::assert_handler calls (Directives.hpp:37-44).
 *   - A lvkb200_stream is one video stream == one lvk::StabilizationFilter instance: it owns one CUDA
 *     stream, the device-resident frame ring, the tracker state and all scratch.  It is externally
 *     synchronised (one caller thread at a time, like the reference, Filters/VideoFilter.cpp:108-152);
 *     distinct handles are fully concurrent.
 *   - Frames are packed 8-bit, 3 channels (the only layout that reaches lvk::remap, Functions/Image.cpp:32,96),
 *     addressed by pointer + pitch (bytes) either in host memory (pinned recommended) or device memory.
 *   - There is NO CPU fallback: every call needs a CUDA device and fails with LVKB200_ERR_NO_DEVICE /
 *     LVKB200_ERR_CUDA otherwise.
 */
#ifndef LVKB200_H
#define LVKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVKB200_ABI_VERSION 1

typedef enum lvkb200_status
{
    LVKB200_OK = 0,
    LVKB200_ERR_INVALID = 1,   /* bad argument / failed reference precondition (LVK_ASSERT equivalent) */
    LVKB200_ERR_CUDA = 2,      /* a CUDA call failed; see lvkb200_last_error() */
    LVKB200_ERR_NO_DEVICE = 3, /* no usable CUDA device */
    LVKB200_ERR_NO_MODEL = 4,  /* motion estimator found no model (reference: empty cv::Mat -> LVK_ASSERT, Math/Homography.cpp:89-95) */
    LVKB200_ERR_CAPACITY = 5   /* caller buffer too small */
} lvkb200_status;

/* lvk::VideoFrame::Format — Data/VideoFrame.hpp:27 (same numbering). */
typedef enum lvkb200_format
{
    LVKB200_BGR = 0, LVKB200_BGRA = 1, LVKB200_RGB = 2, LVKB200_RGBA = 3, LVKB200_YUV = 4, LVKB200_GRAY = 5,
    LVKB200_UNKNOWN = 6
} lvkb200_format;

typedef enum lvkb200_memspace { LVKB200_MEM_HOST = 0, LVKB200_MEM_DEVICE = 1 } lvkb200_memspace;

/* Flat POD mirror of lvk::StabilizationFilterSettings and its bases:
 *   FeatureDetectorSettings  Vision/FeatureDetector.hpp:28-37
 *   FrameTrackerSettings     Vision/FrameTracker.hpp:31-44
 *   PathSmootherSettings     Vision/PathSmoother.hpp:29-39
 *   StabilizationFilterSettings  Filters/StabilizationFilter.hpp:28-39
 * Field names and defaults are the reference's.  The three motion_resolution fields of the reference's
 * multiple-inheritance struct are linked by configure() (StabilizationFilter.cpp:57-58) and are one field here. */
typedef struct lvkb200_settings
{
    int32_t detection_resolution_width, detection_resolution_height; /* {256,256} */
    int32_t detection_regions_width, detection_regions_height;       /* {2,2} */
    int32_t force_detection;                                         /* false */
    float max_feature_density;                                       /* 0.20 */
    float min_feature_density;                                       /* 0.05 */
    float accumulation_rate;                                         /* 2.0 */

    int32_t motion_resolution_width, motion_resolution_height; /* {2,2} (StabilizationFilterSettings) */
    int32_t track_local_motions;                               /* true */
    float temporal_smoothing;                                  /* 1.0 */
    float local_smoothing;                                     /* 20.0 */
    uint64_t min_motion_samples;                               /* 75 */
    float acceptance_threshold;                                /* 8.0 */
    float uniformity_threshold;                                /* 0.20 */

    uint64_t predictive_samples;                          /* 10 */
    float corrective_limits_width, corrective_limits_height; /* {0.1,0.1} */
    float smoothing_steps;                                /* 20.0 */
    float response_rate;                                  /* 0.04 */

    double background_colour[4]; /* cv::Scalar {255,0,255,0}, in frame channel order */
    int32_t crop_to_stable_region; /* false */
    int32_t stabilize_output;      /* true */
    float min_scene_quality;       /* 0.8 */
    float min_tracking_quality;    /* 0.3 */
} lvkb200_settings;

/* Per-submit result (replaces the out-parameter VideoFrame's emptiness/timestamp/format plus
 * FrameTracker::tracking_stability(), Vision/FrameTracker.hpp:56). */
typedef struct lvkb200_result
{
    int32_t has_output;          /* 0 <=> the reference released `output` (StabilizationFilter.cpp:94,134) */
    uint64_t out_timestamp;      /* timestamp of the (delayed) frame written to `out` (WarpMesh.cpp:221) */
    int32_t out_format;          /* its lvkb200_format (WarpMesh.cpp:222) */
    float tracking_stability;    /* inlier ratio of this frame's motion estimate */
    float scene_quality;         /* m_SceneQuality */
    float trust_factor;          /* m_TrustFactor */
    int32_t feature_count;       /* features alive after propagate() */
    int32_t has_motion;          /* 0 <=> FrameTracker::track returned nullopt */
} lvkb200_result;

typedef struct lvkb200_keypoint
{
    float x, y;        /* cv::KeyPoint::pt */
    float response;    /* FAST corner score */
    int32_t class_id;  /* LVK borrows it as the feature age (FrameTracker.cpp:188) */
} lvkb200_keypoint;

typedef struct lvkb200_stream lvkb200_stream;

/* ---- library ------------------------------------------------------------------------------------------------ */

int lvkb200_abi_version(void);
int lvkb200_device_count(void);                 /* 0 when no CUDA device/driver is usable */
const char* lvkb200_last_error(void);           /* thread-local description of the last non-OK status */
const char* lvkb200_status_string(lvkb200_status s);

/* lvk::context::assert_handler — Directives.hpp:37-44, Directives.cpp:27-42.  NULL restores the default
 * (which only records the message; it never aborts inside the library). */
typedef void (*lvkb200_assert_handler)(const char* file, const char* function, const char* assertion);
void lvkb200_set_assert_handler(lvkb200_assert_handler handler);

/* StabilizationFilterSettings{} defaults / the OBS "Homography" preset (Modules/OBS-Plugin/Sources/
 * Stabilisation/VSFilter.cpp:269-280) that north_star names (FAST grid -> LK -> homography RANSAC). */
void lvkb200_settings_default(lvkb200_settings* s);
void lvkb200_settings_obs_homography(lvkb200_settings* s);
/* The OBS "Vector Field" subsystem preset (VSFilter.cpp:257-268): 16x16 motion mesh, local motions on. */
void lvkb200_settings_obs_field(lvkb200_settings* s);

/* ---- lvk::StabilizationFilter ------------------------------------------------------------------------------- */

/* StabilizationFilter::StabilizationFilter(settings) — Filters/StabilizationFilter.cpp:34-38. */
lvkb200_status lvkb200_stream_create(int device, const lvkb200_settings* settings, lvkb200_stream** out);
/* Destroying a stream never writes into caller memory: a device output whose remap is still held back (see
 * lvkb200_stream_submit) is DISCARDED - call lvkb200_stream_sync first if that output is still wanted. */
void lvkb200_stream_destroy(lvkb200_stream* s);

/* StabilizationFilter::configure — StabilizationFilter.cpp:42-65. */
lvkb200_status lvkb200_stream_configure(lvkb200_stream* s, const lvkb200_settings* settings);
lvkb200_status lvkb200_stream_get_settings(const lvkb200_stream* s, lvkb200_settings* out);
/* restart() :139-144, reset_context() :155-159, ready() :148-151, frame_delay() :192-195, stable_region() :199-. */
lvkb200_status lvkb200_stream_restart(lvkb200_stream* s);
lvkb200_status lvkb200_stream_reset_context(lvkb200_stream* s);
int lvkb200_stream_ready(const lvkb200_stream* s);
uint64_t lvkb200_stream_frame_delay(const lvkb200_stream* s);
lvkb200_status lvkb200_stream_stable_region(const lvkb200_stream* s, int frame_width, int frame_height,
                                            int* x, int* y, int* width, int* height);

/* VideoFilter::apply(VideoFrame&& input, VideoFrame& output, profile) -> StabilizationFilter::filter
 * (Filters/VideoFilter.cpp:46-58, Filters/StabilizationFilter.cpp:69-135).
 * The input frame is copied into the stream's device ring (the caller keeps its buffer; the reference moves it into
 * m_FrameQueue): a HOST frame has been consumed when the call returns; a DEVICE frame is copied by stream-ordered
 * work of the library's own CUDA streams, so the caller must not overwrite it before lvkb200_stream_sync (or an
 * event recorded with lvkb200_stream_event_record) has completed - reading it concurrently is fine.
 * `out` may alias `frame` (OBS calls
 * apply(std::move(frame), frame), VSFilter.cpp:358).  When res->has_output == 0 `out` is untouched.
 * With out_space == HOST the call returns after the output has landed in `out`; with DEVICE the output is
 * produced by stream-ordered work that may still be pending (the library can hold the remap back until the next
 * submit, to run it beside that frame's tracking kernels): keep `out` valid and untouched and call
 * lvkb200_stream_sync (the equivalent of Stopwatch::sync_gpu / cv::ocl::finish, Timing/Stopwatch.cpp:127-131)
 * before reading it.  lvkb200_stream_sync and _event_record cover every output requested so far. */
lvkb200_status lvkb200_stream_submit(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                     lvkb200_format format, uint64_t timestamp, lvkb200_memspace frame_space,
                                     void* out, size_t out_pitch, lvkb200_memspace out_space, lvkb200_result* res);
lvkb200_status lvkb200_stream_sync(lvkb200_stream* s);
/* Stopwatch::sync_gpu (Timing/Stopwatch.cpp:127-131: cv::ocl::finish()) for callers that hold no stream handle: drains
 * every CUDA stream of the calling thread's current device. */
lvkb200_status lvkb200_device_synchronize(void);

/* Pipelined operation — VideoFilter::stream(cap, callback, profile) (Filters/VideoFilter.cpp:62-209) overlaps input,
 * filtering and output with three host threads and bounded queues.  The GPU analogue uses two extra CUDA streams:
 *   lvkb200_stream_prefetch(next frame)   starts the host->device upload of the NEXT input (host memory, ideally
 *                                         pinned) while the current frame is processed; the following submit of the
 *                                         same pointer adopts the uploaded buffer.  The caller must keep the frame
 *                                         unmodified until that submit returns.
 *   lvkb200_stream_submit_async(...)      == lvkb200_stream_submit, except that a HOST output is downloaded on the
 *                                         copy-out stream after the call returns; *ticket (0 when there is no
 *                                         output) identifies it.
 *   lvkb200_stream_wait_output(ticket)    returns once that output has landed in the caller's buffer.
 * Results are identical to lvkb200_stream_submit; only the copies overlap.  For full overlap keep TWO outputs in
 * flight (three output buffers): output t's remap is launched inside submit t+1 and its download overlaps frame
 * t+2, so collect it after submit t+2 (waiting earlier is correct, it only gives up part of the overlap). */
lvkb200_status lvkb200_stream_prefetch(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height);
/* The same announcement with the frame's format and memory space (host: upload, device: copy into the stream's ring
 * on the copy-in stream).  Knowing the format, the library also builds the announced frame's detection image and
 * optical-flow pyramid right behind the CURRENT frame's tracking kernels — while the host digests their results and
 * the GPU would idle — so the next submit starts at the optical flow (the input thread of VideoFilter::stream running
 * ahead of the filter thread, Filters/VideoFilter.cpp:76-105).  Results are identical to submitting without it.  The
 * announced frame is processed with the chained-deblocking setting in force at the submit that follows the
 * announcement. */
lvkb200_status lvkb200_stream_prefetch_frame(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                             lvkb200_format format, lvkb200_memspace frame_space);
lvkb200_status lvkb200_stream_submit_async(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                           lvkb200_format format, uint64_t timestamp, lvkb200_memspace frame_space,
                                           void* out, size_t out_pitch, lvkb200_memspace out_space,
                                           lvkb200_result* res, uint64_t* ticket);
lvkb200_status lvkb200_stream_wait_output(lvkb200_stream* s, uint64_t ticket);
/* VideoFilter::stream (Filters/VideoFilter.cpp:62-209) for a sequence that is already in memory (a decoded clip, a
 * capture ring): `count` frames of one geometry and format are filtered back to back, frame i into outs[i], exactly as
 * `count` calls of lvkb200_stream_prefetch_frame(frame i+1) + lvkb200_stream_submit(frame i) would - one call instead of
 * 2 * count crossings of the FFI.  results[i] reports has_output / timestamp of step i (the first frame_delay steps have
 * none and leave outs[i] untouched).  timestamps may be NULL (0, 1, ...).  Device outputs follow the rule of
 * lvkb200_stream_submit: complete after lvkb200_stream_sync.  HOST frames with HOST outputs run pipelined inside the
 * call (upload of frame i+1 and download of output i-1 overlap frame i, two outputs in flight) and every output has
 * landed when it returns; outs[] may then cycle through >= 4 buffers if only the most recent outputs are needed. */
lvkb200_status lvkb200_stream_submit_batch(lvkb200_stream* s, const void* const* frames, size_t pitch, int width, int height,
                                           lvkb200_format format, const uint64_t* timestamps, lvkb200_memspace frame_space,
                                           void* const* outs, size_t out_pitch, lvkb200_memspace out_space, int count,
                                           lvkb200_result* results);

/* ---- lvk::DeblockingFilter --------------------------------------------------------------------- */

/* lvk::DeblockingFilterSettings — Filters/DeblockingFilter.hpp:26-32 (same names and defaults). */
typedef struct lvkb200_deblock_settings
{
    uint32_t detection_levels; /* 3   (> 0) */
    uint32_t block_size;       /* 16  (> 0) */
    uint32_t filter_size;      /* 5   (odd, >= 3) */
    float filter_scaling;      /* 4.0 (> 1): the smooth frame is filtered at 1/filter_scaling resolution */
} lvkb200_deblock_settings;
void lvkb200_deblock_settings_default(lvkb200_deblock_settings* s);

/* DeblockingFilter::filter — Filters/DeblockingFilter.cpp:48-118: adaptively blends the largest whole-macroblock
 * region of the frame with a median-smoothed copy of itself; pixels outside that region pass through.  The
 * reference works in place and moves the input to the output; here `out` may be `frame` (in place) or a separate
 * buffer.  Settings are checked like DeblockingFilter::configure (:36-45); this build covers integer
 * filter_scaling dividing block_size, block_size in {4, 8, 16, 32}, filter_size 3/5/7 (the reference's callers
 * only ever change detection_levels: ADBFilter.cpp:96-97) and returns LVKB200_ERR_INVALID otherwise.
 * Host buffers are staged through the stream's device scratch; the call returns when a host `out` is filled. */
lvkb200_status lvkb200_deblock(lvkb200_stream* s, const lvkb200_deblock_settings* settings, const void* frame,
                               size_t pitch, int width, int height, lvkb200_format format,
                               lvkb200_memspace frame_space, void* out, size_t out_pitch, lvkb200_memspace out_space);

/* CompositeFilter{DeblockingFilter, StabilizationFilter} (Filters/CompositeFilter.cpp:58-88, BASELINE config 5)
 * without leaving the device: every frame submitted from now on is deblocked inside the stream's frame ring before
 * it is tracked and queued.  NULL switches the stage off.  Identical to calling lvkb200_deblock on the frame and
 * submitting the result. */
lvkb200_status lvkb200_stream_set_deblocking(lvkb200_stream* s, const lvkb200_deblock_settings* settings);

/* ---- lvk::ScalingFilter: FSR upscale + sharpen -------------------------------------------------- */

/* lvk::ScalingFilterSettings — Filters/ScalingFilter.hpp:27-32 (same names and defaults). */
typedef struct lvkb200_scaling_settings
{
    int32_t output_width;  /* 1920 (> 0) */
    int32_t output_height; /* 1080 (> 0) */
    float sharpness;       /* 0.8, in [0, 1] */
    int32_t yuv_input;     /* 1: the EASU luma is computed for a YUV frame (Image.cpp:168-179) */
} lvkb200_scaling_settings;
void lvkb200_scaling_settings_default(lvkb200_scaling_settings* s);

/* lvk::upscale(src, dst, size, yuv) — Functions/Image.cpp:155-201, kernel easu_scale (FSR.cl:326-358): FSR-EASU
 * upsampling of a packed 8UC3 frame to dst_width x dst_height (both >= the source's, Image.cpp:157; equal sizes
 * are a plain copy, :162-166).  `dst` must not overlap `src`. */
lvkb200_status lvkb200_upscale(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                               lvkb200_memspace src_space, void* dst, size_t dst_pitch, int dst_width, int dst_height,
                               lvkb200_memspace dst_space, int yuv_input);

/* lvk::sharpen(src, dst, sharpness) — Functions/Image.cpp:205-233, kernel rcas (FSR.cl:460-535): FSR-RCAS on a packed
 * 8UC3 frame, sharpness in [0, 1] (LVK_ASSERT_01, :209).  Every tap reads the unsharpened input: `dst` may equal
 * `src` (the frame is then sharpened through the stream's scratch buffer), which is what the reference's in-place
 * call (ScalingFilter.cpp:57) means but, racing between work-groups, does not guarantee. */
lvkb200_status lvkb200_sharpen(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                               lvkb200_memspace src_space, void* dst, size_t dst_pitch, lvkb200_memspace dst_space,
                               float sharpness);

/* ScalingFilter::filter — Filters/ScalingFilter.cpp:52-59: upscale to settings->output_{width,height}, then sharpen;
 * the intermediate frame never leaves the device.  Settings are checked like ScalingFilter::configure (:41-48). */
lvkb200_status lvkb200_scaling_filter(lvkb200_stream* s, const lvkb200_scaling_settings* settings, const void* frame,
                                      size_t pitch, int width, int height, lvkb200_memspace frame_space, void* out,
                                      size_t out_pitch, lvkb200_memspace out_space);

/* ---- FrameIngest: OBS frame layouts <-> packed 8UC3 frames ---------------------------------------- */

/* The video_format values FrameIngest::Select accepts (Modules/OBS-Plugin/Interop/FrameIngest.cpp:38-76); the
 * numbering is this library's, the names are libobs'. */
typedef enum lvkb200_video_format
{
    LVKB200_VIDEO_I420 = 0, LVKB200_VIDEO_I422 = 1, LVKB200_VIDEO_I444 = 2,  /* planar Y, U, V */
    LVKB200_VIDEO_I40A = 3, LVKB200_VIDEO_I42A = 4, LVKB200_VIDEO_YUVA = 5,  /* the same + an alpha plane LVK ignores */
    LVKB200_VIDEO_NV12 = 6,                                                   /* planar Y + interleaved UV at half size */
    LVKB200_VIDEO_YVYU = 7, LVKB200_VIDEO_YUY2 = 8, LVKB200_VIDEO_UYVY = 9,   /* packed 4:2:2 */
    LVKB200_VIDEO_AYUV = 10,                                                  /* packed 4:4:4 with alpha first */
    LVKB200_VIDEO_Y800 = 11, LVKB200_VIDEO_BGR3 = 12,                         /* direct copies (GRAY / BGR) */
    LVKB200_VIDEO_FORMAT_COUNT = 13
} lvkb200_video_format;

/* The fields of obs_source_frame the ingest reads (FrameIngest.cpp:134-141, :330-470).  linesize[i] is the row
 * pitch of plane i in bytes; 0 means tightly packed (the only case the reference handles: it copies
 * width*height*channels contiguous bytes per plane). */
typedef struct lvkb200_obs_frame
{
    uint8_t* data[4];
    uint32_t linesize[4];
    uint32_t width, height;
    int32_t format;     /* lvkb200_video_format */
    uint64_t timestamp;
} lvkb200_obs_frame;

/* FrameIngest::ocl_format (FrameIngest.cpp:113-116): the lvk::VideoFrame format a video format is ingested as
 * (YUV for every YUV layout, GRAY for Y800, BGR for BGR3); LVKB200_UNKNOWN for an unsupported value
 * (FrameIngest::Select returning nullptr). */
lvkb200_format lvkb200_video_format_ocl(int video_format);

/* FrameIngest::upload_obs_frame -> to_ocl (FrameIngest.cpp:93-103, :479-522 I4XX, :566-584 NV12, :618-645 packed
 * 4:2:2, :689-698 AYUV, :738-747 direct): converts the planes of `src` (host or device memory) into a packed 8UC3
 * frame at `dst` (8UC1 for Y800).  Subsampled chroma is upsampled with cv::resize(INTER_LINEAR) arithmetic,
 * bit-exact with OpenCV's CPU path.  4:2:x layouts need an even width, 4:2:0 an even height. */
lvkb200_status lvkb200_frame_upload(lvkb200_stream* s, const lvkb200_obs_frame* src, lvkb200_memspace src_space,
                                    void* dst, size_t dst_pitch, lvkb200_memspace dst_space);

/* FrameIngest::download_ocl_frame -> to_obs (FrameIngest.cpp:106-110, :526-560, :588-604, :649-678, :702-716,
 * :751-757): the reverse; chroma is subsampled with cv::resize(INTER_AREA) arithmetic.  `format` must be the
 * layout's own ocl format (the reference would colour-convert first via VideoFrame::viewAsFormat; that conversion
 * is outside this path and reported as LVKB200_ERR_INVALID).  An alpha plane of I40A/I42A/YUVA is left untouched,
 * AYUV's alpha is set to 255, as in the reference. */
lvkb200_status lvkb200_frame_download(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                      lvkb200_format format, lvkb200_memspace src_space, lvkb200_obs_frame* dst,
                                      lvkb200_memspace dst_space);

/* VSFilter::filter for asynchronous OBS sources (VSFilter.cpp:352-364 + OBSFrame::from_obs_frame / to_obs_frame,
 * OBSFrame.cpp:93-123): ingest `in`, run StabilizationFilter::filter, and write the delayed stabilized frame into
 * `out` in the same video format (`out` may be `in`: OBS filters the frame in place).  res->has_output == 0 leaves
 * `out` untouched.  Only 1.5 B/px (4:2:0) cross PCIe in each direction instead of 3. */
lvkb200_status lvkb200_stream_submit_obs(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_memspace in_space,
                                         lvkb200_obs_frame* out, lvkb200_memspace out_space, lvkb200_result* res);

/* The pipelined form of lvkb200_stream_submit_obs for HOST frames (ideally pinned) in a planar / semi-planar / packed YUV
 * layout - the OBS analogue of lvkb200_stream_prefetch_frame + lvkb200_stream_submit_async:
 *   lvkb200_stream_prefetch_obs(next)        uploads the planes of the NEXT frame and converts them to the packed frame
 *                                            (FrameIngest to_ocl) on the copy-in stream while the current frame is
 *                                            tracked; its detection image and pyramid are built behind the current
 *                                            frame's tracking chain.  The planes must stay unmodified until the submit
 *                                            of that frame returns.
 *   lvkb200_stream_submit_obs_async(in, out) filters `in` (announced or not); the stabilized frame is converted back to
 *                                            the planes of `out` (to_obs) and downloaded on the copy-out stream after
 *                                            the call returns; *ticket (0: no output) identifies it for
 *                                            lvkb200_stream_wait_output.  `out` must stay valid until then.
 * Same pixels as lvkb200_stream_submit_obs; 1.5 B/px (4:2:0) cross PCIe in each direction and no copy or conversion
 * sits on the frame's critical path.  Keep two outputs in flight as with lvkb200_stream_submit_async. */
lvkb200_status lvkb200_stream_prefetch_obs(lvkb200_stream* s, const lvkb200_obs_frame* in);
lvkb200_status lvkb200_stream_submit_obs_async(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_obs_frame* out,
                                               lvkb200_result* res, uint64_t* ticket);
/* The two calls above over a whole sequence of host frames (in[i] -> out[i], two outputs in flight, all outputs landed
 * on return): lvkb200_stream_submit_batch for OBS plane layouts. */
lvkb200_status lvkb200_stream_submit_obs_batch(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_obs_frame* out, int count,
                                               lvkb200_result* results);

/* CUDA-event timing on the stream's own CUDA stream (torch.cuda.Event cannot see it): record slot `index`
 * (0..LVKB200_EVENT_SLOTS-1) now; elapsed returns the device time between two recorded slots after waiting for
 * the later one.  The per-stage analogue of Stopwatch (Timing/Stopwatch.cpp:42-64) for the bench harness. */
#define LVKB200_EVENT_SLOTS 16
lvkb200_status lvkb200_stream_event_record(lvkb200_stream* s, int index);
lvkb200_status lvkb200_stream_event_elapsed_ms(lvkb200_stream* s, int start_index, int stop_index, float* ms);

/* Debug / parity taps on the last submitted frame (what OBS "test mode" draws, StabilizationFilter.cpp:163-188,
 * and what the parity tests compare against the oracle).  `which`: */
typedef enum lvkb200_debug_item
{
    LVKB200_DBG_DETECTION_IMAGE = 0, /* uint8 det_w*det_h: m_CurrentFrame after resize (FrameTracker.cpp:117) */
    LVKB200_DBG_DETECTED = 1,        /* lvkb200_keypoint[]: FeatureDetector::detect output (FeatureDetector.cpp:170) */
    LVKB200_DBG_LK_MATCHED = 2,      /* float2[]: m_MatchedPoints before filtering (FrameTracker.cpp:140-146) */
    LVKB200_DBG_LK_STATUS = 3,       /* uint8[]: m_MatchStatus */
    LVKB200_DBG_TRACKED = 4,         /* float2[]: m_TrackedPoints after fast_filter (:149) */
    LVKB200_DBG_MATCHED = 5,         /* float2[]: m_MatchedPoints after fast_filter */
    LVKB200_DBG_INLIERS = 6,         /* uint8[]: m_InlierStatus */
    LVKB200_DBG_HOMOGRAPHY = 7,      /* double[9]: estimated frame-to-frame homography (global-motion path) */
    LVKB200_DBG_MOTION = 8,          /* float[rows*cols*2]: motion mesh before the trust factor */
    LVKB200_DBG_CORRECTION = 9,      /* float[rows*cols*2]: PathSmoother::next result (+ scene crop if enabled) */
    LVKB200_DBG_WARP_TRANSFORM = 10, /* double[9]: dst->src transform handed to the remap (WarpMesh.cpp:214) */
    LVKB200_DBG_PROPAGATED = 11,     /* lvkb200_keypoint[]: m_TrackedFeatures after propagate (:183-193) */
    LVKB200_DBG_FAST_COUNTS = 12,    /* int32[regions]: raw FAST keypoints per region this frame, -1 = region skipped */
    LVKB200_DBG_MESH_ITERATIONS = 13 /* int32[1]: conjugate-gradient iterations of the last estimate_local_motions solve */
} lvkb200_debug_item;
/* Taps are only recorded while capture is enabled (it adds a device->host copy of the detection image and a
 * synchronisation per frame, so it is off by default — the equivalent of OBS "test mode", VSFilter.cpp:356-383). */
lvkb200_status lvkb200_stream_set_debug_capture(lvkb200_stream* s, int enable);
/* Copies up to `capacity` bytes; *size receives the full size in bytes (0 if the item was not produced). */
lvkb200_status lvkb200_stream_debug_fetch(lvkb200_stream* s, lvkb200_debug_item which, void* buffer,
                                          size_t capacity, size_t* size);

/* Per-stage device time of the last submit, in microseconds (CUDA events): the per-stage analogue of
 * VideoFilter::timings() (Filters/VideoFilter.hpp:58).  Stage order: ingest, pyramid, fast, lk, estimate, remap. */
#define LVKB200_STAGE_COUNT 6
lvkb200_status lvkb200_stream_stage_times_us(lvkb200_stream* s, float times[LVKB200_STAGE_COUNT]);
/* Running per-stage totals (microseconds of device time, CUDA events on the stream's own CUDA stream) and sample
 * counts since the last reset — the Stopwatch history (Timing/Stopwatch.cpp:142-166) per stage.  Harvested lazily,
 * so keeping them costs no synchronisation inside submit. */
lvkb200_status lvkb200_stream_stage_totals_us(lvkb200_stream* s, double totals[LVKB200_STAGE_COUNT],
                                              uint64_t counts[LVKB200_STAGE_COUNT], int reset);
/* VideoFilter::apply(..., profile) (Filters/VideoFilter.cpp:46-51): with profiling on, every stage is bracketed by
 * CUDA events and the tracking chain runs as individual launches; off (default) the chain replays as one CUDA graph
 * and only the remap kernel is timed. */
lvkb200_status lvkb200_stream_set_profiling(lvkb200_stream* s, int enable);
/* Number of CUDA kernels this library has launched in this process (all streams). */
uint64_t lvkb200_kernel_launch_count(void);
/* Arithmetic build of the FSR-EASU kernels (lvk::remap / lvk::upscale, Functions/OpenCL/Sources/FSR.cl:98-452), process
 * wide.  0 (default): the "contract" build — multiply-adds fused by the compiler and native_recip = one hardware
 * reciprocal, the liberties OpenCL C gives the reference's own device compiler (<= 1 LSB from the reference's kernels
 * compiled with contraction).  1: the exact build — bit-identical to oracle/easu_ref.c (one explicit
 * contraction rule, IEEE division), ~2x slower; also the initial value when LVKB200_REMAP_EXACT=1 is set. */
void lvkb200_set_remap_exact(int exact);
int lvkb200_remap_exact(void);

/* ---- stage-level entry points (used by the parity tests: oracle inputs -> one GPU stage) --------------------- */

/* lvk::remap(src, dst, homography, background, inverted=true) + easu_remap_homography
 * (Functions/Image.cpp:85-151, Functions/OpenCL/Sources/FSR.cl:407-452).  t_inv: row-major dst->src 3x3. */
lvkb200_status lvkb200_remap_homography(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                        lvkb200_memspace src_space, void* dst, size_t dst_pitch,
                                        lvkb200_memspace dst_space, const double t_inv[9],
                                        const uint8_t background[3], int yuv_input);

/* WarpMesh::apply for meshes larger than 2x2 + lvk::remap(src, dst, offset_map, bg) + easu_remap
 * (Math/WarpMesh.cpp:187-192, Image.cpp:28-81, FSR.cl:362-403).  offsets: rows*cols float2, normalized. */
lvkb200_status lvkb200_remap_mesh(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                  lvkb200_memspace src_space, void* dst, size_t dst_pitch, lvkb200_memspace dst_space,
                                  const float* offsets, int mesh_cols, int mesh_rows, const uint8_t background[3],
                                  int yuv_input);

/* WarpMesh::apply (both branches) — Math/WarpMesh.cpp:183-223. */
lvkb200_status lvkb200_warp_mesh_apply(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                       lvkb200_memspace src_space, void* dst, size_t dst_pitch,
                                       lvkb200_memspace dst_space, const float* offsets, int mesh_cols, int mesh_rows,
                                       const uint8_t background[3], int yuv_input, double t_inv_out[9]);

/* VideoFrame::viewAsFormat(GRAY) + cv::resize(INTER_AREA) (Data/VideoFrame.cpp:187-317,
 * Vision/FrameTracker.cpp:117), fused.  det_out: host buffer det_w*det_h bytes. */
lvkb200_status lvkb200_detection_image(lvkb200_stream* s, const void* frame, size_t pitch, int width, int height,
                                       lvkb200_format format, lvkb200_memspace space, uint8_t* det_out, int det_w,
                                       int det_h);

/* cv::FastFeatureDetector(threshold, nonmax=true, TYPE_9_16)::detect(image(roi)) as called at
 * Vision/FeatureDetector.cpp:130-134.  image: host uint8 width*height; keypoints are ROI-relative, in OpenCV's
 * (y,x) emission order. */
lvkb200_status lvkb200_fast_detect(lvkb200_stream* s, const uint8_t* image, int width, int height, int roi_x,
                                   int roi_y, int roi_w, int roi_h, int threshold, lvkb200_keypoint* keypoints,
                                   int capacity, int* count);

/* cv::SparsePyrLKOpticalFlow((11,11), 3, {COUNT+EPS,5,0.01})::calc as configured at
 * Vision/FrameTracker.cpp:41-48 and called at :140-146.  Host buffers.
 * call_index = number of calc() calls already made on the same tracker object: upstream OpenCV squares the object's
 * TermCriteria::epsilon in place on every call, and the reference reuses one m_OpticalTracker for all frames, so the
 * (n+1)-th call stops on |delta|^2 <= 0.01^(2^(n+1)).  lvkb200_stream_submit advances this per stream by itself. */
lvkb200_status lvkb200_lk_track(lvkb200_stream* s, const uint8_t* prev, const uint8_t* next, int width, int height,
                                const float* points, int count, int call_index, float* matched, uint8_t* status);

/* FrameTracker::estimate_global_motion's cv::findHomography(..., UsacParams) (Vision/FrameTracker.cpp:337-359).
 * Host buffers; h_out row-major 3x3 (h33 == 1); mask[i] = 1 <=> reprojection error < threshold. */
lvkb200_status lvkb200_find_homography(lvkb200_stream* s, const float* src_points, const float* dst_points,
                                       int count, float threshold, double h_out[9], uint8_t* mask);

/* FrameTracker::estimate_global_motion's cv::estimateAffinePartial2D(..., cv::RANSAC, threshold, 50) branch, taken
 * when the feature distribution is <= 0.6 (Vision/FrameTracker.cpp:37,171,362-373).  h_out = the 2x3 similarity
 * [a -b tx; b a ty] promoted to 3x3 (Homography::FromAffineMatrix, Math/Homography.cpp:44-57); mask = inliers of the
 * best minimal model, the transform = least squares over them. */
lvkb200_status lvkb200_estimate_affine_partial(lvkb200_stream* s, const float* src_points, const float* dst_points,
                                               int count, float threshold, double h_out[9], uint8_t* mask);

/* FrameTracker::estimate_local_motions (Vision/FrameTracker.cpp:200-321): least-squares motion mesh.
 * mesh_state: in/out m_OptimizedMesh (2*cols*rows floats); offsets_out rows*cols float2; mask per point. */
lvkb200_status lvkb200_estimate_local_motions(lvkb200_stream* s, const float* tracked, const float* matched,
                                              int count, float* mesh_state, float* offsets_out, uint8_t* mask);

#ifdef __cplusplus
}
#endif
#endif /* LVKB200_H */
