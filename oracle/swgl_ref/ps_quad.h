// TEST INFRASTRUCTURE — hand-instantiated SWGL program scaffolding.
//
// Restates the vertex-side infrastructure shared by the ps_quad_* programs,
// webrender/res/ps_quad.glsl (SWGL branches: SWGL_ANTIALIAS is defined, so
// vLocalPos / vTransformBounds AA varyings are compiled out; base.glsl:37-43),
// in the glsl.h vocabulary glsl-to-cxx would emit.  Run classes follow the
// translator's rule: anything that depends on the per-vertex attribute
// `aPosition` is a 4-lane vector (one lane per quad corner), everything else
// is scalar.

#pragma once

#define WR_EDGE_AA_LEFT 1
#define WR_EDGE_AA_TOP 2
#define WR_EDGE_AA_RIGHT 4
#define WR_EDGE_AA_BOTTOM 8

#define WR_PART_CENTER 0
#define WR_PART_LEFT 1
#define WR_PART_TOP 2
#define WR_PART_RIGHT 3
#define WR_PART_BOTTOM 4
#define WR_PART_ALL 5

#define WR_QF_IS_OPAQUE 1
#define WR_QF_APPLY_DEVICE_CLIP 2
#define WR_QF_IGNORE_DEVICE_SCALE 4
#define WR_QF_USE_AA_SEGMENTS 8
#define WR_QF_IS_MASK 16

#define WR_INVALID_SEGMENT_INDEX 0xff
#define WR_AA_PIXEL_RADIUS 2.0f

struct PsQuadVertBase : VertexShaderImpl, WrCommon {
  // attributes (ps_quad.glsl:75, shared.glsl:71)
  vec2 aPosition;
  ivec4_scalar aData;
  int a_aPosition, a_aData;

  // flat varyings (ps_quad.glsl:41-47)
  vec4_scalar v_color;
  ivec4_scalar v_flags;

  PsQuadVertBase() {
    a_aPosition = attrib_locations.add("aPosition");
    a_aData = attrib_locations.add("aData");
    sampler_mask |= WR_S_TransformPalette | WR_S_RenderTasks | WR_S_GpuBufferF |
                    WR_S_GpuBufferI;
  }

  struct QuadSegment {
    RectWithEndpoint rect;
    RectWithEndpoint uv_rect;
  };

  struct PrimitiveInfo {
    vec2 local_pos;
    RectWithEndpoint local_prim_rect;
    RectWithEndpoint local_clip_rect;
    QuadSegment segment;
    int edge_flags;
    int quad_flags;
    ivec2_scalar pattern_input;
  };

  struct QuadPrimitive {
    RectWithEndpoint bounds;
    RectWithEndpoint clip;
    RectWithEndpoint uv_rect;
    vec4_scalar pattern_scale_offset;
    vec4_scalar color;
  };

  // ps_quad.glsl:96-105
  QuadSegment fetch_segment(int base, int index) {
    QuadSegment seg;
    int addr = base + 5 + index * 2;
    vec4_scalar t0 = fetch_gpu_buffer_f(addr, 0);
    vec4_scalar t1 = fetch_gpu_buffer_f(addr, 1);
    seg.rect = RectWithEndpoint{t0.sel(X, Y), t0.sel(Z, W)};
    seg.uv_rect = RectWithEndpoint{t1.sel(X, Y), t1.sel(Z, W)};
    return seg;
  }

  // ps_quad.glsl:107-119
  QuadPrimitive fetch_primitive(int index) {
    QuadPrimitive prim;
    vec4_scalar t0 = fetch_gpu_buffer_f(index, 0);
    vec4_scalar t1 = fetch_gpu_buffer_f(index, 1);
    vec4_scalar t2 = fetch_gpu_buffer_f(index, 2);
    vec4_scalar t3 = fetch_gpu_buffer_f(index, 3);
    vec4_scalar t4 = fetch_gpu_buffer_f(index, 4);
    prim.bounds = RectWithEndpoint{t0.sel(X, Y), t0.sel(Z, W)};
    prim.clip = RectWithEndpoint{t1.sel(X, Y), t1.sel(Z, W)};
    prim.uv_rect = RectWithEndpoint{t2.sel(X, Y), t2.sel(Z, W)};
    prim.pattern_scale_offset = t3;
    prim.color = t4;
    return prim;
  }

  // ps_quad.glsl:185-220
  vec2 write_vertex(vec2 local_pos, float z, Transform& transform,
                    vec2_scalar content_origin, RectWithEndpoint task_rect,
                    float device_pixel_scale, int quad_flags) {
    vec2 vi_local_pos;
    vec4 world_pos = transform.m * vec4(local_pos, Float(0.0f), Float(1.0f));
    vec2 device_pos = world_pos.sel(X, Y) * Float(device_pixel_scale);
    if ((quad_flags & WR_QF_APPLY_DEVICE_CLIP) != 0) {
      RectWithEndpoint device_clip_rect = RectWithEndpoint{
          content_origin, content_origin + task_rect.p1 - task_rect.p0};
      device_pos = clamp(device_pos, vec2(device_clip_rect.p0),
                         vec2(device_clip_rect.p1));
      vi_local_pos =
          (transform.inv_m * vec4(device_pos / Float(device_pixel_scale),
                                  Float(0.0f), Float(1.0f)))
              .sel(X, Y);
    } else {
      vi_local_pos = local_pos;
    }
    vec2_scalar final_offset = -content_origin + task_rect.p0;
    gl_Position =
        uTransform * vec4(device_pos + final_offset * world_pos.w,
                          z * world_pos.w, world_pos.w);
    return vi_local_pos;
  }

  static float edge_aa_offset(int edge, int flags) {
    return ((flags & edge) != 0) ? WR_AA_PIXEL_RADIUS : 0.0f;
  }
  static vec2_scalar scale_offset_map_point(vec4_scalar so, vec2_scalar p) {
    return p * so.sel(X, Y) + so.sel(Z, W);
  }
  static vec2 scale_offset_map_point(vec4_scalar so, vec2 p) {
    return p * vec2(so.sel(X, Y)) + vec2(so.sel(Z, W));
  }
  static RectWithEndpoint scale_offset_map_rect(vec4_scalar so,
                                                RectWithEndpoint r) {
    return RectWithEndpoint{scale_offset_map_point(so, r.p0),
                            scale_offset_map_point(so, r.p1)};
  }

  // ps_quad.glsl:239-358
  PrimitiveInfo quad_primive_info() {
    // decode_instance (ps_quad.glsl:166-183)
    int prim_address_i = aData.x;
    int prim_address_f = aData.y;
    int quad_flags = (aData.z >> 24) & 0xff;
    int edge_flags = (aData.z >> 16) & 0xff;
    int part_index = (aData.z >> 8) & 0xff;
    int segment_index = (aData.z >> 0) & 0xff;
    int picture_task_address = aData.w;

    // fetch_header (ps_quad.glsl:133-145)
    ivec4_scalar header = fetch_from_gpu_buffer_1i(prim_address_i);
    int transform_id = header.x;
    int z_id = header.y;
    ivec2_scalar pattern_input(header.z, header.w);

    Transform transform = fetch_transform(transform_id);
    PictureTask task = fetch_picture_task(picture_task_address);
    QuadPrimitive prim = fetch_primitive(prim_address_f);
    float z = float(z_id);

    QuadSegment seg;
    if (segment_index == WR_INVALID_SEGMENT_INDEX) {
      seg.rect = prim.bounds;
      seg.uv_rect = prim.uv_rect;
    } else {
      seg = fetch_segment(prim_address_f, segment_index);
    }

    RectWithEndpoint local_coverage_rect = seg.rect;
    local_coverage_rect.p0 = max(local_coverage_rect.p0, prim.clip.p0);
    local_coverage_rect.p1 = min(local_coverage_rect.p1, prim.clip.p1);
    local_coverage_rect.p1 =
        max(local_coverage_rect.p0, local_coverage_rect.p1);

    switch (part_index) {
      case WR_PART_LEFT:
        local_coverage_rect.p1.x = local_coverage_rect.p0.x + WR_AA_PIXEL_RADIUS;
        swgl_antiAlias(WR_EDGE_AA_LEFT);
        break;
      case WR_PART_TOP:
        local_coverage_rect.p0.x = local_coverage_rect.p0.x + WR_AA_PIXEL_RADIUS;
        local_coverage_rect.p1.x = local_coverage_rect.p1.x - WR_AA_PIXEL_RADIUS;
        local_coverage_rect.p1.y = local_coverage_rect.p0.y + WR_AA_PIXEL_RADIUS;
        swgl_antiAlias(WR_EDGE_AA_TOP);
        break;
      case WR_PART_RIGHT:
        local_coverage_rect.p0.x = local_coverage_rect.p1.x - WR_AA_PIXEL_RADIUS;
        swgl_antiAlias(WR_EDGE_AA_RIGHT);
        break;
      case WR_PART_BOTTOM:
        local_coverage_rect.p0.x = local_coverage_rect.p0.x + WR_AA_PIXEL_RADIUS;
        local_coverage_rect.p1.x = local_coverage_rect.p1.x - WR_AA_PIXEL_RADIUS;
        local_coverage_rect.p0.y = local_coverage_rect.p1.y - WR_AA_PIXEL_RADIUS;
        swgl_antiAlias(WR_EDGE_AA_BOTTOM);
        break;
      case WR_PART_CENTER:
        local_coverage_rect.p0.x += edge_aa_offset(WR_EDGE_AA_LEFT, edge_flags);
        local_coverage_rect.p1.x -= edge_aa_offset(WR_EDGE_AA_RIGHT, edge_flags);
        local_coverage_rect.p0.y += edge_aa_offset(WR_EDGE_AA_TOP, edge_flags);
        local_coverage_rect.p1.y -=
            edge_aa_offset(WR_EDGE_AA_BOTTOM, edge_flags);
        break;
      case WR_PART_ALL:
      default:
        swgl_antiAlias(edge_flags);
        break;
    }

    vec2 local_pos =
        mix(local_coverage_rect.p0, local_coverage_rect.p1, aPosition);

    float device_pixel_scale = task.device_pixel_scale;
    if ((quad_flags & WR_QF_IGNORE_DEVICE_SCALE) != 0) {
      device_pixel_scale = 1.0f;
    }

    vec2 vi_local_pos =
        write_vertex(local_pos, z, transform, task.content_origin,
                     task.task_rect, device_pixel_scale, quad_flags);

    v_color = prim.color;

    vec4_scalar pattern_tx = prim.pattern_scale_offset;
    seg.rect = scale_offset_map_rect(pattern_tx, seg.rect);

    return PrimitiveInfo{scale_offset_map_point(pattern_tx, vi_local_pos),
                         scale_offset_map_rect(pattern_tx, prim.bounds),
                         scale_offset_map_rect(pattern_tx, prim.clip),
                         seg,
                         edge_flags,
                         quad_flags,
                         pattern_input};
  }

  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs,
                           uint32_t start, int instance, int count) {
    PsQuadVertBase* self = (PsQuadVertBase*)impl;
    load_attrib(self->aPosition,
                attribs[self->attrib_locations.locs[self->a_aPosition]], start,
                instance, count);
    load_flat_attrib(self->aData,
                     attribs[self->attrib_locations.locs[self->a_aData]],
                     start, instance, count);
  }
};
