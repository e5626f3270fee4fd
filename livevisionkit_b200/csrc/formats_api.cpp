// C-ABI of the frame ingest / egress stage (include/lvkb200.h, "FrameIngest") — the host side of formats.cu.
// Mirrors Modules/OBS-Plugin/Interop/FrameIngest.cpp: Select (:38-76), upload_obs_frame / download_ocl_frame
// (:93-110) and the per-layout to_ocl / to_obs pairs (:479-757).
#include <cstring>

#include "common.hpp"
#include "stream_impl.hpp"

namespace lvkb200
{
namespace
{

inline size_t align16(size_t v) { return (v + 15) / 16 * 16; }

enum class Kind { Planar, SemiPlanar, Packed422, Packed444, Direct };

// What FrameIngest::Select + the ingest constructors decide from the video format.
struct Layout
{
    Kind kind;
    lvkb200_format ocl;  // FrameIngest::ocl_format
    int sub_x, sub_y;    // chroma subsampling (m_ChromaScaling: FrameIngest.cpp:461-467)
    int y_off, u_off, v_off;  // byte offsets inside a packed macro-pixel (Packed422 / Packed444)
    int channels;        // Direct: channels copied (DirectIngest::to_ocl, :738-747)
};

bool describe(int format, Layout& l)
{
    switch (format)
    {
    case LVKB200_VIDEO_I420: case LVKB200_VIDEO_I40A: l = {Kind::Planar, LVKB200_YUV, 2, 2, 0, 0, 0, 0}; return true;
    case LVKB200_VIDEO_I422: case LVKB200_VIDEO_I42A: l = {Kind::Planar, LVKB200_YUV, 2, 1, 0, 0, 0, 0}; return true;
    case LVKB200_VIDEO_I444: case LVKB200_VIDEO_YUVA: l = {Kind::Planar, LVKB200_YUV, 1, 1, 0, 0, 0, 0}; return true;
    case LVKB200_VIDEO_NV12: l = {Kind::SemiPlanar, LVKB200_YUV, 2, 2, 0, 0, 1, 0}; return true;
    // packed 4:2:2 — m_YFirst = (format != UYVY), m_UFirst = (format != YVYU)  (FrameIngest.cpp:608-614)
    case LVKB200_VIDEO_YUY2: l = {Kind::Packed422, LVKB200_YUV, 2, 1, 0, 1, 3, 0}; return true;  // Y0 U Y1 V
    case LVKB200_VIDEO_YVYU: l = {Kind::Packed422, LVKB200_YUV, 2, 1, 0, 3, 1, 0}; return true;  // Y0 V Y1 U
    case LVKB200_VIDEO_UYVY: l = {Kind::Packed422, LVKB200_YUV, 2, 1, 1, 0, 2, 0}; return true;  // U Y0 V Y1
    case LVKB200_VIDEO_AYUV: l = {Kind::Packed444, LVKB200_YUV, 1, 1, 1, 2, 3, 0}; return true;  // A Y U V
    case LVKB200_VIDEO_Y800: l = {Kind::Direct, LVKB200_GRAY, 1, 1, 0, 0, 0, 1}; return true;
    case LVKB200_VIDEO_BGR3: l = {Kind::Direct, LVKB200_BGR, 1, 1, 0, 0, 0, 3}; return true;
    default: return false;
    }
}

struct PlaneGeometry
{
    int planes;           // planes that carry data LVK touches
    size_t row_bytes[3];  // payload bytes per row
    int rows[3];
};

PlaneGeometry geometry(const Layout& l, int w, int h)
{
    const int cw = w / l.sub_x, ch = h / l.sub_y;
    PlaneGeometry g{};
    switch (l.kind)
    {
    case Kind::Planar:
        g.planes = 3;
        g.row_bytes[0] = (size_t)w; g.rows[0] = h;
        g.row_bytes[1] = g.row_bytes[2] = (size_t)cw; g.rows[1] = g.rows[2] = ch;
        break;
    case Kind::SemiPlanar:
        g.planes = 2;
        g.row_bytes[0] = (size_t)w; g.rows[0] = h;
        g.row_bytes[1] = (size_t)cw * 2; g.rows[1] = ch;
        break;
    case Kind::Packed422: g.planes = 1; g.row_bytes[0] = (size_t)w * 2; g.rows[0] = h; break;
    case Kind::Packed444: g.planes = 1; g.row_bytes[0] = (size_t)w * 4; g.rows[0] = h; break;
    case Kind::Direct: g.planes = 1; g.row_bytes[0] = (size_t)w * l.channels; g.rows[0] = h; break;
    }
    return g;
}

// Component references into (device) planes p[] with pitches pitch[].
void components(const Layout& l, uint8_t* const p[3], const size_t pitch[3], PlaneRef& Y, PlaneRef& U, PlaneRef& V,
                PlaneRef& A)
{
    A = PlaneRef{nullptr, 0, 0};
    switch (l.kind)
    {
    case Kind::Planar:
        Y = {p[0], pitch[0], 1}; U = {p[1], pitch[1], 1}; V = {p[2], pitch[2], 1};
        break;
    case Kind::SemiPlanar:
        Y = {p[0], pitch[0], 1}; U = {p[1] + l.u_off, pitch[1], 2}; V = {p[1] + l.v_off, pitch[1], 2};
        break;
    case Kind::Packed422:
        Y = {p[0] + l.y_off, pitch[0], 2}; U = {p[0] + l.u_off, pitch[0], 4}; V = {p[0] + l.v_off, pitch[0], 4};
        break;
    case Kind::Packed444:
        Y = {p[0] + l.y_off, pitch[0], 4}; U = {p[0] + l.u_off, pitch[0], 4}; V = {p[0] + l.v_off, pitch[0], 4};
        A = {p[0], pitch[0], 4};
        break;
    case Kind::Direct: Y = U = V = {p[0], pitch[0], 1}; break;
    }
}

// FrameIngest::test_obs_frame (FrameIngest.cpp:130-141) + the size limits of upload_planes (:352-358) + the even-size
// requirement of the subsampled layouts.
lvkb200_status check_frame(const lvkb200_obs_frame* f, const Layout& l, const PlaneGeometry& g)
{
    LVKB_REQUIRE(f->data[0] != nullptr && f->width > 0 && f->height > 0);
    LVKB_REQUIRE(f->width <= 8192 && f->height <= 8192);  // MAX_TEXTURE_SIZE
    LVKB_REQUIRE(f->width % l.sub_x == 0 && f->height % l.sub_y == 0);
    for (int i = 0; i < g.planes; i++)
    {
        LVKB_REQUIRE(f->data[i] != nullptr);
        LVKB_REQUIRE(f->linesize[i] == 0 || f->linesize[i] >= g.row_bytes[i]);
    }
    return LVKB200_OK;
}

}  // namespace
}  // namespace lvkb200

using namespace lvkb200;

extern "C" {

lvkb200_format lvkb200_video_format_ocl(int video_format)
{
    Layout l;
    return describe(video_format, l) ? l.ocl : LVKB200_UNKNOWN;
}

lvkb200_status lvkb200_frame_upload(lvkb200_stream* s, const lvkb200_obs_frame* src, lvkb200_memspace src_space,
                                    void* dst, size_t dst_pitch, lvkb200_memspace dst_space)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr);
    Layout l;
    LVKB_REQUIRE(describe(src->format, l));  // FrameIngest::Select returned nullptr
    const int w = (int)src->width, h = (int)src->height;
    const PlaneGeometry g = geometry(l, w, h);
    LVKB_TRY(check_frame(src, l, g));
    const int out_ch = l.kind == Kind::Direct ? l.channels : 3;
    LVKB_REQUIRE(dst_pitch >= (size_t)w * out_ch);
    LVKB_CUDA(cudaSetDevice(s->device));

    // device view of the planes
    uint8_t* p[3] = {nullptr, nullptr, nullptr};
    size_t pitch[3] = {0, 0, 0};
    if (src_space == LVKB200_MEM_DEVICE)
    {
        for (int i = 0; i < g.planes; i++)
        {
            p[i] = src->data[i];
            pitch[i] = src->linesize[i] ? src->linesize[i] : g.row_bytes[i];
        }
    }
    else
    {
        size_t off[3], total = 0;
        for (int i = 0; i < g.planes; i++)
        {
            pitch[i] = align16(g.row_bytes[i]);
            off[i] = total;
            total += pitch[i] * g.rows[i];
        }
        LVKB_CUDA(s->planes_in.ensure(total));
        for (int i = 0; i < g.planes; i++)
        {
            p[i] = s->planes_in.as<uint8_t>() + off[i];
            const size_t sp = src->linesize[i] ? src->linesize[i] : g.row_bytes[i];
            LVKB_CUDA(cudaMemcpy2DAsync(p[i], pitch[i], src->data[i], sp, g.row_bytes[i], g.rows[i],
                                        cudaMemcpyHostToDevice, s->cs));
        }
    }

    uint8_t* d = nullptr;
    size_t dp = 0;
    LVKB_TRY(s->stage_frame_out(dst, dst_pitch, w, h, out_ch, dst_space, &d, &dp));
    if (l.kind == Kind::Direct)
    {
        LVKB_CUDA(cudaMemcpy2DAsync(d, dp, p[0], pitch[0], g.row_bytes[0], h, cudaMemcpyDeviceToDevice, s->cs));
    }
    else
    {
        PlaneRef Y, U, V, A;
        components(l, p, pitch, Y, U, V, A);
        LVKB_CUDA(s->format_plan.prepare(w, h, w / l.sub_x, h / l.sub_y, s->cs));
        LVKB_CUDA(launch_planes_to_packed(s->cs, s->format_plan, Y, U, V, d, dp));
    }
    LVKB_TRY(s->finish_frame_out(dst, dst_pitch, w, h, out_ch, dst_space));
    if (src_space == LVKB200_MEM_HOST && dst_space == LVKB200_MEM_DEVICE) LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_frame_download(lvkb200_stream* s, const void* src, size_t src_pitch, int width, int height,
                                      lvkb200_format format, lvkb200_memspace src_space, lvkb200_obs_frame* dst,
                                      lvkb200_memspace dst_space)
{
    LVKB_REQUIRE(s != nullptr && src != nullptr && dst != nullptr);
    Layout l;
    LVKB_REQUIRE(describe(dst->format, l));
    LVKB_REQUIRE((int)dst->width == width && (int)dst->height == height);
    // download_ocl_frame converts to the ingest's own format first (viewAsFormat); that colour conversion is not
    // part of this path
    LVKB_REQUIRE(format == l.ocl);
    const PlaneGeometry g = geometry(l, width, height);
    LVKB_TRY(check_frame(dst, l, g));
    const int in_ch = l.kind == Kind::Direct ? l.channels : 3;
    LVKB_CUDA(cudaSetDevice(s->device));

    const uint8_t* d = nullptr;
    size_t dp = 0;
    LVKB_TRY(s->stage_frame_in(src, src_pitch, width, height, in_ch, src_space, &d, &dp));

    uint8_t* p[3] = {nullptr, nullptr, nullptr};
    size_t pitch[3] = {0, 0, 0}, off[3] = {0, 0, 0};
    if (dst_space == LVKB200_MEM_DEVICE)
    {
        for (int i = 0; i < g.planes; i++)
        {
            p[i] = dst->data[i];
            pitch[i] = dst->linesize[i] ? dst->linesize[i] : g.row_bytes[i];
        }
    }
    else
    {
        size_t total = 0;
        for (int i = 0; i < g.planes; i++)
        {
            pitch[i] = align16(g.row_bytes[i]);
            off[i] = total;
            total += pitch[i] * g.rows[i];
        }
        LVKB_CUDA(s->planes_out.ensure(total));
        for (int i = 0; i < g.planes; i++) p[i] = s->planes_out.as<uint8_t>() + off[i];
    }

    if (l.kind == Kind::Direct)
    {
        LVKB_CUDA(cudaMemcpy2DAsync(p[0], pitch[0], d, dp, g.row_bytes[0], height, cudaMemcpyDeviceToDevice, s->cs));
    }
    else
    {
        PlaneRef Y, U, V, A;
        components(l, p, pitch, Y, U, V, A);
        LVKB_CUDA(launch_packed_to_planes(s->cs, d, dp, width, height, l.sub_x, l.sub_y, Y, U, V,
                                          l.kind == Kind::Packed444 ? &A : nullptr, l.kind == Kind::SemiPlanar));
    }
    if (dst_space == LVKB200_MEM_HOST)
    {
        for (int i = 0; i < g.planes; i++)
        {
            const size_t hp = dst->linesize[i] ? dst->linesize[i] : g.row_bytes[i];
            LVKB_CUDA(cudaMemcpy2DAsync(dst->data[i], hp, p[i], pitch[i], g.row_bytes[i], g.rows[i],
                                        cudaMemcpyDeviceToHost, s->cs));
        }
        LVKB_CUDA(cudaStreamSynchronize(s->cs));
    }
    else if (src_space == LVKB200_MEM_HOST)
        LVKB_CUDA(cudaStreamSynchronize(s->cs));
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_submit_obs(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_memspace in_space,
                                         lvkb200_obs_frame* out, lvkb200_memspace out_space, lvkb200_result* res)
{
    LVKB_REQUIRE(s != nullptr && in != nullptr && out != nullptr && res != nullptr);
    Layout l;
    LVKB_REQUIRE(describe(in->format, l));
    LVKB_REQUIRE(l.ocl != LVKB200_GRAY);  // lvk::remap takes CV_8UC3 only (Functions/Image.cpp:32,96)
    LVKB_REQUIRE(out->format == in->format && out->width == in->width && out->height == in->height);
    const int w = (int)in->width, h = (int)in->height;
    LVKB_CUDA(cudaSetDevice(s->device));
    const size_t pitch = align16((size_t)w * 3);
    LVKB_CUDA(s->obs_frame_in.ensure(pitch * h));
    LVKB_CUDA(s->obs_frame_out.ensure(pitch * h));
    // OBSFrame::from_obs_frame -> StabilizationFilter::filter -> OBSFrame::to_obs_frame, all on the device
    LVKB_TRY(lvkb200_frame_upload(s, in, in_space, s->obs_frame_in.ptr, pitch, LVKB200_MEM_DEVICE));
    LVKB_TRY(lvkb200_stream_submit(s, s->obs_frame_in.ptr, pitch, w, h, l.ocl, in->timestamp, LVKB200_MEM_DEVICE,
                                   s->obs_frame_out.ptr, pitch, LVKB200_MEM_DEVICE, res));
    if (!res->has_output) return LVKB200_OK;
    LVKB_TRY(s->join_remap(s->cs));  // the egress kernel reads the remap's output (which runs on its own stream)
    LVKB_TRY(lvkb200_frame_download(s, s->obs_frame_out.ptr, pitch, w, h, (lvkb200_format)res->out_format,
                                    LVKB200_MEM_DEVICE, out, out_space));
    out->timestamp = res->out_timestamp;
    return LVKB200_OK;
}

// ---- pipelined OBS-layout path -----------------------------------------------------------------------------------------
// VideoFilter::stream for asynchronous OBS sources: the planes of frame t+1 are uploaded and converted (to_ocl) on the
// copy-in stream while frame t is tracked, and the stabilized frame t-1 is converted back (to_obs) and downloaded on the
// copy-out stream: 1.5 B/px (4:2:0) cross PCIe in each direction instead of 3, and none of it sits on the frame's
// critical path.  Results are identical to lvkb200_stream_submit_obs.

lvkb200_status lvkb200_stream_prefetch_obs(lvkb200_stream* s, const lvkb200_obs_frame* in)
{
    LVKB_REQUIRE(s != nullptr && in != nullptr);
    Layout l;
    LVKB_REQUIRE(describe(in->format, l));
    LVKB_REQUIRE(l.kind != Kind::Direct && l.ocl != LVKB200_GRAY);  // planar / semi-planar / packed YUV layouts
    const int w = (int)in->width, h = (int)in->height;
    const PlaneGeometry g = geometry(l, w, h);
    LVKB_TRY(check_frame(in, l, g));
    LVKB_CUDA(cudaSetDevice(s->device));
    const size_t row = (size_t)w * 3, fpitch = align16(row);
    LVKB_TRY(s->ensure_pipeline());
    LVKB_TRY(s->ensure_frame_pool(fpitch * h));
    // a slot that still holds an announced, not yet submitted frame is never overwritten while the other one is free
    // (submit_obs_async of an unannounced frame comes BETWEEN the announcement of frame t+1 and its submit)
    int k = s->prefetch_next;
    if (s->prefetched_ptr[k] != nullptr && s->prefetched_ptr[k ^ 1] == nullptr) k ^= 1;
    s->prefetch_next = k ^ 1;
    lvkb200_stream::QueuedFrame& ps = s->prefetch_slot[k];
    ps.pitch = fpitch;
    ps.w = w;
    ps.h = h;
    if (ps.buf.capacity < fpitch * h)
    {
        LVKB_TRY(s->sync_all());  // the buffer being replaced may still be read by a queued remap
        LVKB_CUDA(ps.buf.ensure(fpitch * h));
    }
    LVKB_TRY(s->wait_frame_buffers_free(s->cs_in));
    uint8_t* p[3] = {nullptr, nullptr, nullptr};
    size_t pitch[3] = {0, 0, 0}, off[3] = {0, 0, 0}, total = 0;
    for (int i = 0; i < g.planes; i++)
    {
        pitch[i] = align16(g.row_bytes[i]);
        off[i] = total;
        total += pitch[i] * g.rows[i];
    }
    // staging of this slot's PREVIOUS frame was consumed by a kernel queued earlier on the same (in-order) stream
    if (s->planes_in_async[k].capacity < total)
    {
        LVKB_CUDA(cudaStreamSynchronize(s->cs_in));
        LVKB_CUDA(s->planes_in_async[k].ensure(total));
    }
    for (int i = 0; i < g.planes; i++)
    {
        p[i] = s->planes_in_async[k].template as<uint8_t>() + off[i];
        const size_t sp = in->linesize[i] ? in->linesize[i] : g.row_bytes[i];
        LVKB_CUDA(cudaMemcpy2DAsync(p[i], pitch[i], in->data[i], sp, g.row_bytes[i], g.rows[i], cudaMemcpyHostToDevice, s->cs_in));
    }
    PlaneRef Y, U, V, A;
    components(l, p, pitch, Y, U, V, A);
    LVKB_CUDA(s->format_plan_in.prepare(w, h, w / l.sub_x, h / l.sub_y, s->cs_in));
    LVKB_CUDA(launch_planes_to_packed(s->cs_in, s->format_plan_in, Y, U, V, ps.buf.template as<uint8_t>(), ps.pitch));
    LVKB_CUDA(cudaEventRecord(s->prefetch_done[k], s->cs_in));
    s->prefetched_ptr[k] = in->data[0];
    s->lookahead[k] = lvkb200_stream::Lookahead{};
    s->lookahead[k].announced = true;
    s->lookahead[k].format = l.ocl;
    return LVKB200_OK;
}

lvkb200_status lvkb200_stream_submit_obs_async(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_obs_frame* out,
                                               lvkb200_result* res, uint64_t* ticket)
{
    LVKB_REQUIRE(s != nullptr && in != nullptr && out != nullptr && res != nullptr && ticket != nullptr);
    Layout l;
    LVKB_REQUIRE(describe(in->format, l));
    LVKB_REQUIRE(l.kind != Kind::Direct && l.ocl != LVKB200_GRAY);
    LVKB_REQUIRE(out->format == in->format && out->width == in->width && out->height == in->height);
    LVKB_REQUIRE(s->settings.stabilize_output);  // the pass-through configuration uses lvkb200_stream_submit_obs
    const int w = (int)in->width, h = (int)in->height;
    LVKB_TRY(check_frame(out, l, geometry(l, w, h)));
    LVKB_CUDA(cudaSetDevice(s->device));
    bool announced = false;
    for (int k = 0; k < 2; k++) announced |= (s->prefetched_ptr[k] == in->data[0]);
    if (!announced) LVKB_TRY(lvkb200_stream_prefetch_obs(s, in));
    s->deferred_output = true;
    s->next_egress = true;
    s->next_egress_frame = *out;
    s->last_ticket = 0;
    // the key of the announced frame is its first plane; the packed frame already sits in the prefetch slot
    const lvkb200_status st = s->submit(in->data[0], align16((size_t)w * 3), w, h, l.ocl, in->timestamp, LVKB200_MEM_HOST,
                                        out->data[0], align16((size_t)w * 3), LVKB200_MEM_HOST, res);
    s->deferred_output = false;
    s->next_egress = false;
    *ticket = (st == LVKB200_OK && res->has_output) ? s->last_ticket : 0;
    if (st == LVKB200_OK && res->has_output) out->timestamp = res->out_timestamp;
    return st;
}

lvkb200_status lvkb200_stream_submit_obs_batch(lvkb200_stream* s, const lvkb200_obs_frame* in, lvkb200_obs_frame* out, int count,
                                               lvkb200_result* results)
{
    LVKB_REQUIRE(s != nullptr && in != nullptr && out != nullptr && results != nullptr && count >= 0);
    uint64_t in_flight[3] = {0, 0, 0};
    int n_flight = 0;
    lvkb200_status st = LVKB200_OK;
    for (int i = 0; i < count && st == LVKB200_OK; i++)
    {
        if (i + 1 < count) LVKB_TRY(lvkb200_stream_prefetch_obs(s, &in[i + 1]));
        uint64_t ticket = 0;
        st = lvkb200_stream_submit_obs_async(s, &in[i], &out[i], &results[i], &ticket);
        if (st == LVKB200_OK && ticket)
        {
            in_flight[n_flight++] = ticket;
            if (n_flight > 2)
            {
                st = s->wait_output(in_flight[0]);
                in_flight[0] = in_flight[1]; in_flight[1] = in_flight[2];
                n_flight = 2;
            }
        }
    }
    for (int k = 0; k < n_flight; k++)
    {
        const lvkb200_status w = s->wait_output(in_flight[k]);
        if (st == LVKB200_OK) st = w;
    }
    return st;
}

}  // extern "C"

lvkb200_status lvkb200_stream::egress_planes(cudaStream_t stream, const uint8_t* packed, size_t pitch, int width, int height,
                                             lvkb200_format format, const lvkb200_obs_frame& dst, int slot)
{
    using namespace lvkb200;
    Layout l;
    LVKB_REQUIRE(describe(dst.format, l) && format == l.ocl);
    const PlaneGeometry g = geometry(l, width, height);
    uint8_t* p[3] = {nullptr, nullptr, nullptr};
    size_t pp[3] = {0, 0, 0}, off[3] = {0, 0, 0}, total = 0;
    for (int i = 0; i < g.planes; i++)
    {
        pp[i] = align16(g.row_bytes[i]);
        off[i] = total;
        total += pp[i] * g.rows[i];
    }
    if (planes_out_async[slot].capacity < total)
    {
        LVKB_CUDA(cudaStreamSynchronize(stream));
        LVKB_CUDA(planes_out_async[slot].ensure(total));
    }
    for (int i = 0; i < g.planes; i++) p[i] = planes_out_async[slot].as<uint8_t>() + off[i];
    PlaneRef Y, U, V, A;
    components(l, p, pp, Y, U, V, A);
    LVKB_CUDA(launch_packed_to_planes(stream, packed, pitch, width, height, l.sub_x, l.sub_y, Y, U, V,
                                      l.kind == Kind::Packed444 ? &A : nullptr, l.kind == Kind::SemiPlanar));
    for (int i = 0; i < g.planes; i++)
    {
        const size_t hp = dst.linesize[i] ? dst.linesize[i] : g.row_bytes[i];
        LVKB_CUDA(cudaMemcpy2DAsync(dst.data[i], hp, p[i], pp[i], g.row_bytes[i], g.rows[i], cudaMemcpyDeviceToHost, stream));
    }
    return LVKB200_OK;
}
