// MOCK of the part of libobs' <obs-module.h> that the OBS plugin's filter sources (Modules/OBS-Plugin/Sources/
// Stabilisation/VSFilter.cpp, Sources/Enhancement/ADBFilter.cpp and the plugin headers they include) name: opaque
// handles, obs_source_frame, the properties / settings calls, the few graphics calls its header-only templates make.
// Declarations follow libobs' public C API; bodies are inert.  It exists so that those translation units can be
// COMPILED UNCHANGED against lvk-compat in an image without OBS Studio (tests/test_compat_cpu.py).  Test infrastructure.
#pragma once

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdlib>

#define MAX_AV_PLANES 8

// ---- opaque handles ---------------------------------------------------------------------------------------------------
struct obs_source;     typedef struct obs_source obs_source_t;
struct obs_data;       typedef struct obs_data obs_data_t;
struct obs_properties; typedef struct obs_properties obs_properties_t;
struct obs_property;   typedef struct obs_property obs_property_t;
struct gs_texture;     typedef struct gs_texture gs_texture_t;
struct gs_effect;      typedef struct gs_effect gs_effect_t;
struct gs_effect_param; typedef struct gs_effect_param gs_eparam_t;
struct gs_stage_surface; typedef struct gs_stage_surface gs_stagesurf_t;
struct vec2 { float x, y; };
struct vec4 { float x, y, z, w; };

// ---- media-io ---------------------------------------------------------------------------------------------------------
enum video_format
{
    VIDEO_FORMAT_NONE, VIDEO_FORMAT_I420, VIDEO_FORMAT_NV12, VIDEO_FORMAT_YVYU, VIDEO_FORMAT_YUY2, VIDEO_FORMAT_UYVY,
    VIDEO_FORMAT_RGBA, VIDEO_FORMAT_BGRA, VIDEO_FORMAT_BGRX, VIDEO_FORMAT_Y800, VIDEO_FORMAT_I444, VIDEO_FORMAT_BGR3,
    VIDEO_FORMAT_I422, VIDEO_FORMAT_I40A, VIDEO_FORMAT_I42A, VIDEO_FORMAT_YUVA, VIDEO_FORMAT_AYUV
};
enum video_colorspace { VIDEO_CS_DEFAULT, VIDEO_CS_601, VIDEO_CS_709, VIDEO_CS_SRGB };
enum video_range_type { VIDEO_RANGE_DEFAULT, VIDEO_RANGE_PARTIAL, VIDEO_RANGE_FULL };
struct obs_source_frame
{
    uint8_t* data[MAX_AV_PLANES];
    uint32_t linesize[MAX_AV_PLANES];
    uint32_t width, height;
    uint64_t timestamp;
    enum video_format format;
    float color_matrix[16];
    bool full_range;
    float color_range_min[3], color_range_max[3];
    bool flip;
    uint8_t flags;
    volatile long refs;
    bool prev_frame;
};
struct obs_video_info
{
    const char* graphics_module;
    uint32_t fps_num, fps_den;
    uint32_t base_width, base_height, output_width, output_height;
    enum video_format output_format;
    uint32_t adapter;
    bool gpu_conversion;
    enum video_colorspace colorspace;
    enum video_range_type range;
    int scale_type;
};
inline bool obs_get_video_info(struct obs_video_info* info) { if (info) { info->fps_num = 60; info->fps_den = 1; } return true; }

// ---- util -------------------------------------------------------------------------------------------------------------
enum { LOG_ERROR = 100, LOG_WARNING = 200, LOG_INFO = 300, LOG_DEBUG = 400 };
inline void blog(int, const char*, ...) {}
inline void bfree(void* ptr) { std::free(ptr); }
inline const char* obs_module_text(const char* lookup) { return lookup; }
inline char* obs_module_file(const char*) { return nullptr; }

// ---- sources ----------------------------------------------------------------------------------------------------------
enum obs_allow_direct_render { OBS_NO_DIRECT_RENDERING, OBS_ALLOW_DIRECT_RENDERING };
inline const char* obs_source_get_name(const obs_source_t*) { return ""; }
inline const char* obs_source_get_id(const obs_source_t*) { return ""; }
inline obs_source_t* obs_filter_get_parent(const obs_source_t*) { return nullptr; }
inline obs_source_t* obs_filter_get_target(const obs_source_t*) { return nullptr; }
inline uint32_t obs_source_get_base_width(obs_source_t*) { return 0; }
inline uint32_t obs_source_get_base_height(obs_source_t*) { return 0; }
inline void obs_source_update_properties(obs_source_t*) {}
inline void obs_enter_graphics() {}
inline void obs_leave_graphics() {}

// ---- graphics ---------------------------------------------------------------------------------------------------------
enum gs_color_format { GS_UNKNOWN, GS_A8, GS_R8, GS_RGBA, GS_BGRX, GS_BGRA };
inline bool obs_source_process_filter_begin(obs_source_t*, enum gs_color_format, enum obs_allow_direct_render) { return false; }
inline void obs_source_process_filter_tech_end(obs_source_t*, gs_effect_t*, uint32_t, uint32_t, const char*) {}
inline uint32_t gs_texture_get_width(const gs_texture_t*) { return 0; }
inline uint32_t gs_texture_get_height(const gs_texture_t*) { return 0; }
inline bool gs_get_linear_srgb() { return false; }
inline bool gs_framebuffer_srgb_enabled() { return false; }
inline void gs_enable_framebuffer_srgb(bool) {}
inline gs_eparam_t* gs_effect_get_param_by_name(const gs_effect_t*, const char*) { return nullptr; }
inline void gs_effect_set_texture(gs_eparam_t*, gs_texture_t*) {}
inline void gs_effect_set_texture_srgb(gs_eparam_t*, gs_texture_t*) {}
inline bool gs_effect_loop(gs_effect_t*, const char*) { return false; }
inline void gs_draw_sprite(gs_texture_t*, uint32_t, uint32_t, uint32_t) {}
inline gs_effect_t* gs_effect_create_from_file(const char*, char**) { return nullptr; }

// ---- settings ---------------------------------------------------------------------------------------------------------
inline bool obs_data_get_bool(obs_data_t*, const char*) { return false; }
inline long long obs_data_get_int(obs_data_t*, const char*) { return 0; }
inline double obs_data_get_double(obs_data_t*, const char*) { return 0.0; }
inline const char* obs_data_get_string(obs_data_t*, const char*) { return ""; }
inline void obs_data_set_int(obs_data_t*, const char*, long long) {}
inline void obs_data_set_default_bool(obs_data_t*, const char*, bool) {}
inline void obs_data_set_default_int(obs_data_t*, const char*, long long) {}
inline void obs_data_set_default_double(obs_data_t*, const char*, double) {}
inline void obs_data_set_default_string(obs_data_t*, const char*, const char*) {}

// ---- properties -------------------------------------------------------------------------------------------------------
enum obs_combo_type { OBS_COMBO_TYPE_INVALID, OBS_COMBO_TYPE_EDITABLE, OBS_COMBO_TYPE_LIST };
enum obs_combo_format { OBS_COMBO_FORMAT_INVALID, OBS_COMBO_FORMAT_INT, OBS_COMBO_FORMAT_FLOAT, OBS_COMBO_FORMAT_STRING };
enum obs_group_type { OBS_COMBO_INVALID, OBS_GROUP_NORMAL, OBS_GROUP_CHECKABLE };
typedef bool (*obs_property_modified_t)(obs_properties_t* props, obs_property_t* property, obs_data_t* settings);
inline obs_properties_t* obs_properties_create() { return nullptr; }
inline obs_property_t* obs_properties_get(obs_properties_t*, const char*) { return nullptr; }
inline obs_property_t* obs_properties_add_bool(obs_properties_t*, const char*, const char*) { return nullptr; }
inline obs_property_t* obs_properties_add_int(obs_properties_t*, const char*, const char*, int, int, int) { return nullptr; }
inline obs_property_t* obs_properties_add_int_slider(obs_properties_t*, const char*, const char*, int, int, int) { return nullptr; }
inline obs_property_t* obs_properties_add_float_slider(obs_properties_t*, const char*, const char*, double, double, double) { return nullptr; }
inline obs_property_t* obs_properties_add_list(obs_properties_t*, const char*, const char*, enum obs_combo_type, enum obs_combo_format) { return nullptr; }
inline obs_property_t* obs_properties_add_color(obs_properties_t*, const char*, const char*) { return nullptr; }
inline obs_property_t* obs_properties_add_group(obs_properties_t*, const char*, const char*, enum obs_group_type, obs_properties_t*) { return nullptr; }
inline size_t obs_property_list_add_string(obs_property_t*, const char*, const char*) { return 0; }
inline void obs_property_int_set_suffix(obs_property_t*, const char*) {}
inline void obs_property_set_enabled(obs_property_t*, bool) {}
inline void obs_property_set_modified_callback(obs_property_t*, obs_property_modified_t) {}

// ---- source registration (obs-source.h) ---------------------------------------------------------------------------------
enum obs_source_type { OBS_SOURCE_TYPE_INPUT, OBS_SOURCE_TYPE_FILTER, OBS_SOURCE_TYPE_TRANSITION, OBS_SOURCE_TYPE_SCENE };
#define OBS_SOURCE_VIDEO (1 << 0)
#define OBS_SOURCE_AUDIO (1 << 1)
#define OBS_SOURCE_ASYNC (1 << 2)
#define OBS_SOURCE_ASYNC_VIDEO (OBS_SOURCE_ASYNC | OBS_SOURCE_VIDEO)
#define OBS_SOURCE_CUSTOM_DRAW (1 << 3)
struct obs_source_info
{
    const char* id;
    enum obs_source_type type;
    uint32_t output_flags;
    const char* (*get_name)(void* type_data);
    void* (*create)(obs_data_t* settings, obs_source_t* source);
    void (*destroy)(void* data);
    uint32_t (*get_width)(void* data);
    uint32_t (*get_height)(void* data);
    void (*get_defaults)(obs_data_t* settings);
    obs_properties_t* (*get_properties)(void* data);
    void (*update)(void* data, obs_data_t* settings);
    void (*video_tick)(void* data, float seconds);
    void (*video_render)(void* data, gs_effect_t* effect);
    struct obs_source_frame* (*filter_video)(void* data, struct obs_source_frame* frame);
};
inline void obs_register_source(const struct obs_source_info*) {}
