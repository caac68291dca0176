// TEST INFRASTRUCTURE — hand-instantiated SWGL program scaffolding.
//
// Restates webrender/res/brush.glsl + prim_shared.glsl (vertex side) for the
// SWGL feature set (SWGL_CLIP_MASK, SWGL_ANTIALIAS defined: clip masks and edge
// AA are requested from the rasteriser through swgl_clipMask/swgl_antiAlias
// instead of varyings; base.glsl:37-43).  Derived programs provide
// VECS_PER_SPECIFIC_BRUSH and brush_vs().
#pragma once

#define WR_BRUSH_FLAG_PERSPECTIVE_INTERPOLATION 1
#define WR_BRUSH_FLAG_SEGMENT_RELATIVE 2
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_X 4
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_Y 8
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_X_ROUND 16
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_Y_ROUND 32
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_X_CENTERED 64
#define WR_BRUSH_FLAG_SEGMENT_REPEAT_Y_CENTERED 128
#define WR_BRUSH_FLAG_SEGMENT_NINEPATCH_MIDDLE 256
#define WR_BRUSH_FLAG_TEXEL_RECT 512
#define WR_BRUSH_FLAG_FORCE_AA 1024
#define WR_BRUSH_FLAG_NORMALIZED_UVS 2048

struct PrimVertBase : VertexShaderImpl, WrCommon {
  vec2 aPosition;
  ivec4_scalar aData;
  int a_aPosition, a_aData;

  PrimVertBase() {
    a_aPosition = attrib_locations.add("aPosition");
    a_aData = attrib_locations.add("aData");
    sampler_mask |= WR_S_TransformPalette | WR_S_RenderTasks | WR_S_GpuCache |
                    WR_S_PrimitiveHeadersF | WR_S_PrimitiveHeadersI | WR_S_ClipMask;
  }

  // prim_shared.glsl:44-60
  struct Instance {
    int prim_header_address, clip_address, segment_index, flags,
        resource_address, brush_kind;
  };
  Instance decode_instance_attributes() {
    Instance instance;
    instance.prim_header_address = aData.x;
    instance.clip_address = aData.y;
    instance.segment_index = aData.z & 0xffff;
    instance.flags = aData.z >> 16;
    instance.resource_address = aData.w & 0xffffff;
    instance.brush_kind = aData.w >> 24;
    return instance;
  }

  // prim_shared.glsl:62-96
  struct PrimitiveHeader {
    RectWithEndpoint local_rect, local_clip_rect;
    float z;
    int specific_prim_address, transform_id, picture_task_address;
    ivec4_scalar user_data;
  };
  PrimitiveHeader fetch_prim_header(int index) {
    PrimitiveHeader ph;
    ivec2_scalar uv_f = get_fetch_uv(index, 2U);
    vec4_scalar local_rect = texelFetch(sPrimitiveHeadersF, uv_f + ivec2_scalar(0, 0), 0);
    vec4_scalar local_clip_rect = texelFetch(sPrimitiveHeadersF, uv_f + ivec2_scalar(1, 0), 0);
    ph.local_rect = RectWithEndpoint{local_rect.sel(X, Y), local_rect.sel(Z, W)};
    ph.local_clip_rect = RectWithEndpoint{local_clip_rect.sel(X, Y), local_clip_rect.sel(Z, W)};
    ivec2_scalar uv_i = get_fetch_uv(index, 2U);
    ivec4_scalar data0 = texelFetch(sPrimitiveHeadersI, uv_i + ivec2_scalar(0, 0), 0);
    ivec4_scalar data1 = texelFetch(sPrimitiveHeadersI, uv_i + ivec2_scalar(1, 0), 0);
    ph.z = float(data0.x);
    ph.specific_prim_address = data0.y;
    ph.transform_id = data0.z;
    ph.picture_task_address = data0.w;
    ph.user_data = data1;
    return ph;
  }

  // prim_shared.glsl:98-128
  struct VertexInfo {
    vec2 local_pos;
    vec4 world_pos;
  };
  VertexInfo write_vertex(vec2 local_pos, RectWithEndpoint local_clip_rect, float z,
                          Transform& transform, PictureTask& task) {
    vec2 clamped_local_pos = clamp(local_pos, vec2(local_clip_rect.p0), vec2(local_clip_rect.p1));
    vec4 world_pos = transform.m * vec4(clamped_local_pos, Float(0.0f), Float(1.0f));
    vec2 device_pos = world_pos.sel(X, Y) * Float(task.device_pixel_scale);
    vec2_scalar final_offset = -task.content_origin + task.task_rect.p0;
    gl_Position = uTransform * vec4(device_pos + final_offset * world_pos.w,
                                    z * world_pos.w, world_pos.w);
    return VertexInfo{clamped_local_pos, world_pos};
  }

  // prim_shared.glsl:130-179 (SWGL_ANTIALIAS branch)
  RectWithEndpoint clip_and_init_antialiasing(RectWithEndpoint segment_rect,
                                              RectWithEndpoint clip_rect, int edge_flags) {
    bool cx = clip_rect.p0.x > segment_rect.p0.x, cy = clip_rect.p0.y > segment_rect.p0.y;
    bool cz = clip_rect.p1.x < segment_rect.p1.x, cw = clip_rect.p1.y < segment_rect.p1.y;
    swgl_antiAlias(edge_flags | (cx ? 1 : 0) | (cy ? 2 : 0) | (cz ? 4 : 0) | (cw ? 8 : 0));
    segment_rect.p0 = clamp(segment_rect.p0, clip_rect.p0, clip_rect.p1);
    segment_rect.p1 = clamp(segment_rect.p1, clip_rect.p0, clip_rect.p1);
    return segment_rect;
  }

  // prim_shared.glsl:181-200 (SWGL_CLIP_MASK branch)
  void write_clip(ClipArea& area, PictureTask& task) {
    swgl_clipMask(sClipMask,
                  (task.task_rect.p0 - task.content_origin) -
                      (area.task_rect.p0 - area.screen_origin),
                  area.task_rect.p0, (area.task_rect.p1 - area.task_rect.p0));
  }

  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start,
                           int instance, int count) {
    PrimVertBase* self = (PrimVertBase*)impl;
    load_attrib(self->aPosition, attribs[self->attrib_locations.locs[self->a_aPosition]], start,
                instance, count);
    load_flat_attrib(self->aData, attribs[self->attrib_locations.locs[self->a_aData]], start,
                     instance, count);
  }
};

// brush.glsl:95-222.  D = derived program's vertex struct (CRTP) providing
//   static const int VECS_PER_SPECIFIC_BRUSH;
//   void brush_vs(VertexInfo&, int prim_address, RectWithEndpoint local_rect,
//                 RectWithEndpoint segment_rect, ivec4_scalar prim_user_data,
//                 int specific_resource_address, mat4_scalar transform,
//                 PictureTask&, int brush_flags, vec4_scalar segment_data);
template <typename D>
struct BrushVertBase : PrimVertBase {
  void brush_shader_main_vs(Instance& instance, PrimitiveHeader& ph, Transform& transform,
                            PictureTask& pic_task, ClipArea& clip_area) {
    int edge_flags = (instance.flags >> 12) & 0xf;
    int brush_flags = instance.flags & 0xfff;
    vec4_scalar segment_data;
    RectWithEndpoint segment_rect;
    if (instance.segment_index == 0xffff) {
      segment_rect = ph.local_rect;
      segment_data = vec4_scalar(0.0f);
    } else {
      int segment_address =
          ph.specific_prim_address + D::VECS_PER_SPECIFIC_BRUSH + instance.segment_index * 2;
      vec4_scalar s0 = fetch_gpu_cache(segment_address, 0);
      vec4_scalar s1 = fetch_gpu_cache(segment_address, 1);
      segment_rect = RectWithEndpoint{s0.sel(X, Y), s0.sel(Z, W)};
      segment_rect.p0 += ph.local_rect.p0;
      segment_rect.p1 += ph.local_rect.p0;
      segment_data = s1;
    }
    RectWithEndpoint adjusted_segment_rect = segment_rect;
    bool antialiased =
        !transform.is_axis_aligned || ((brush_flags & WR_BRUSH_FLAG_FORCE_AA) != 0);
    if (antialiased) {
      adjusted_segment_rect =
          clip_and_init_antialiasing(segment_rect, ph.local_clip_rect, edge_flags);
      ph.local_clip_rect.p0 = vec2_scalar(-1.0e16f);
      ph.local_clip_rect.p1 = vec2_scalar(1.0e16f);
    }
    vec2 local_pos = mix(adjusted_segment_rect.p0, adjusted_segment_rect.p1, aPosition);
    VertexInfo vi = write_vertex(local_pos, ph.local_clip_rect, ph.z, transform, pic_task);
    write_clip(clip_area, pic_task);
    static_cast<D*>(this)->brush_vs(vi, ph.specific_prim_address, ph.local_rect, segment_rect,
                                    ph.user_data, instance.resource_address, transform.m,
                                    pic_task, brush_flags, segment_data);
  }

  void main() {
    Instance instance = decode_instance_attributes();
    PrimitiveHeader ph = fetch_prim_header(instance.prim_header_address);
    Transform transform = fetch_transform(ph.transform_id);
    PictureTask task = fetch_picture_task(ph.picture_task_address);
    ClipArea clip_area = fetch_clip_area(instance.clip_address);
    brush_shader_main_vs(instance, ph, transform, task, clip_area);
  }
};
