// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "cs_clip_rectangle" and
// "cs_clip_rectangle FAST_PATH": webrender/res/cs_clip_rectangle.glsl +
// clip_shared.glsl + ellipse.glsl + transform.glsl (get_node_pos), restated in
// the glsl.h vocabulary.  Run classes follow glsl-to-cxx: values depending on
// aPosition (vertex stage) or on the varying vLocalPos (fragment stage) are
// 4-lane vectors, everything else scalar.
#pragma once

struct ClipVertBase : VertexShaderImpl, WrCommon {
  vec2 aPosition;
  vec4_scalar aClipDeviceArea, aClipOrigins;
  float aDevicePixelScale;
  ivec2_scalar aTransformIds;
  int a_aPosition, a_aClipDeviceArea, a_aClipOrigins, a_aDevicePixelScale, a_aTransformIds;

  vec4_scalar vTransformBounds;  // transform.glsl:5 (flat)

  ClipVertBase() {
    a_aPosition = attrib_locations.add("aPosition");
    a_aClipDeviceArea = attrib_locations.add("aClipDeviceArea");
    a_aClipOrigins = attrib_locations.add("aClipOrigins");
    a_aDevicePixelScale = attrib_locations.add("aDevicePixelScale");
    a_aTransformIds = attrib_locations.add("aTransformIds");
    sampler_mask |= WR_S_TransformPalette | WR_S_RenderTasks | WR_S_GpuCache;
  }

  // transform.glsl:48-86
  static vec4 get_node_pos(vec2 pos, Transform& transform) {
    vec4_scalar ah = transform.m * vec4_scalar(0.0f, 0.0f, 0.0f, 1.0f);
    vec3_scalar a = ah.sel(X, Y, Z) / ah.w;
    vec3_scalar n = transpose(make_mat3(transform.inv_m)) * vec3_scalar(0.0f, 0.0f, 1.0f);
    // untransform(pos, n, a, inv_m)
    vec3 p = vec3(pos, Float(-10000.0f));
    vec3_scalar d = vec3_scalar(0.0f, 0.0f, 1.0f);
    Float t = 0.0f;
    // ray_plane
    float denom = dot(n, d);
    if (abs(denom) > 1e-6f) {
      vec3 dd = vec3(a) - p;
      t = dot(dd, vec3(n)) / denom;
    }
    Float z = p.z + d.z * t;
    vec4 r = transform.inv_m * vec4(pos, z, Float(1.0f));
    return r;
  }

  struct ClipVertexInfo {
    vec4 local_pos;
    RectWithEndpoint clipped_local_rect;
  };

  // clip_shared.glsl:43-78
  ClipVertexInfo write_clip_tile_vertex(RectWithEndpoint local_clip_rect, Transform& prim_transform,
                                        Transform& clip_transform, RectWithEndpoint sub_rect,
                                        vec2_scalar task_origin, vec2_scalar screen_origin,
                                        float device_pixel_scale) {
    vec2 device_pos = screen_origin + mix(sub_rect.p0, sub_rect.p1, aPosition);
    vec2 world_pos = device_pos / Float(device_pixel_scale);
    vec4 pos = prim_transform.m * vec4(world_pos, Float(0.0f), Float(1.0f));
    pos.x /= pos.w;
    pos.y /= pos.w;
    pos.z /= pos.w;
    vec4 p = get_node_pos(pos.sel(X, Y), clip_transform);
    vec4 local_pos = p * pos.w;
    vec4 vertex_pos = vec4(task_origin + mix(sub_rect.p0, sub_rect.p1, aPosition), Float(0.0f),
                           Float(1.0f));
    gl_Position = uTransform * vertex_pos;
    vTransformBounds = make_vec4(local_clip_rect.p0, local_clip_rect.p1);
    return ClipVertexInfo{local_pos, local_clip_rect};
  }

  void load_common_attribs(VertexAttrib* attribs, uint32_t start, int instance, int count) {
    load_attrib(aPosition, attribs[attrib_locations.locs[a_aPosition]], start, instance, count);
    load_flat_attrib(aClipDeviceArea, attribs[attrib_locations.locs[a_aClipDeviceArea]], start, instance, count);
    load_flat_attrib(aClipOrigins, attribs[attrib_locations.locs[a_aClipOrigins]], start, instance, count);
    load_flat_attrib(aDevicePixelScale, attribs[attrib_locations.locs[a_aDevicePixelScale]], start, instance, count);
    load_flat_attrib(aTransformIds, attribs[attrib_locations.locs[a_aTransformIds]], start, instance, count);
  }
};

static inline vec2_scalar wr_inverse_radii_squared(vec2_scalar radii) {  // ellipse.glsl:7-9
  return 1.0f / max(radii * radii, 1.0e-6f);
}

template <bool FAST>
struct cs_clip_rectangle_vert_t : ClipVertBase {
  typedef cs_clip_rectangle_vert_t Self;
  vec2_scalar aClipLocalPos;
  vec4_scalar aClipLocalRect;
  float aClipMode;
  vec4_scalar aClipRect[4], aClipRadii[4];  // TL, TR, BL, BR
  int a_loc[11];

  // outputs
  vec4 vLocalPos;
  vec3_scalar vClipParams;
  vec4_scalar vClipCenter_Radius_TL, vClipCenter_Radius_TR, vClipCenter_Radius_BL, vClipCenter_Radius_BR;
  vec3_scalar vClipPlane_TL, vClipPlane_TR, vClipPlane_BL, vClipPlane_BR;
  vec2_scalar vClipMode;

  struct InterpOutputs {
    vec4_scalar vLocalPos;
  };

  cs_clip_rectangle_vert_t() {
    static const char* names[11] = {"aClipLocalPos", "aClipLocalRect", "aClipMode", "aClipRect_TL",
                                    "aClipRadii_TL", "aClipRect_TR", "aClipRadii_TR", "aClipRect_BL",
                                    "aClipRadii_BL", "aClipRect_BR", "aClipRadii_BR"};
    for (int i = 0; i < 11; i++) a_loc[i] = attrib_locations.add(names[i]);
    init_vertex_abi();
  }

  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    self->load_common_attribs(attribs, start, instance, count);
    auto& L = self->attrib_locations.locs;
    load_flat_attrib(self->aClipLocalPos, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aClipLocalRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aClipMode, attribs[L[self->a_loc[2]]], start, instance, count);
    for (int i = 0; i < 4; i++) {
      load_flat_attrib(self->aClipRect[i], attribs[L[self->a_loc[3 + 2 * i]]], start, instance, count);
      load_flat_attrib(self->aClipRadii[i], attribs[L[self->a_loc[4 + 2 * i]]], start, instance, count);
    }
  }

  // cs_clip_rectangle.glsl:81-153
  void main() {
    vec2_scalar cmi_local_pos = aClipLocalPos;
    Transform clip_transform = fetch_transform(aTransformIds.x);
    Transform prim_transform = fetch_transform(aTransformIds.y);
    RectWithEndpoint local_rect = RectWithEndpoint{aClipLocalRect.sel(X, Y), aClipLocalRect.sel(Z, W)};
    vec2_scalar diff = cmi_local_pos - local_rect.p0;
    local_rect.p0 = cmi_local_pos;
    local_rect.p1 += diff;
    ClipVertexInfo vi = write_clip_tile_vertex(
        local_rect, prim_transform, clip_transform,
        RectWithEndpoint{aClipDeviceArea.sel(X, Y), aClipDeviceArea.sel(Z, W)}, aClipOrigins.sel(X, Y),
        aClipOrigins.sel(Z, W), aDevicePixelScale);
    vClipMode.x = aClipMode;
    vLocalPos = vi.local_pos;
    if (FAST) {
      vec2_scalar half_size = 0.5f * (local_rect.p1 - local_rect.p0);
      float radius = aClipRadii[0].x;
      vec2 sub = (half_size + cmi_local_pos) * vi.local_pos.w;
      vLocalPos.x -= sub.x;
      vLocalPos.y -= sub.y;
      vec2_scalar hs = half_size - vec2_scalar(radius);
      vClipParams = vec3_scalar(hs.x, hs.y, radius);
    } else {
      RectWithEndpoint clip_rect = local_rect;
      vec2_scalar r_tl = aClipRadii[0].sel(X, Y);
      vec2_scalar r_tr = aClipRadii[1].sel(X, Y);
      vec2_scalar r_bl = aClipRadii[2].sel(X, Y);
      vec2_scalar r_br = aClipRadii[3].sel(X, Y);
      vClipCenter_Radius_TL = make_vec4(clip_rect.p0 + r_tl, wr_inverse_radii_squared(r_tl));
      vClipCenter_Radius_TR = make_vec4(clip_rect.p1.x - r_tr.x, clip_rect.p0.y + r_tr.y,
                                        wr_inverse_radii_squared(r_tr));
      vClipCenter_Radius_BR = make_vec4(clip_rect.p1 - r_br, wr_inverse_radii_squared(r_br));
      vClipCenter_Radius_BL = make_vec4(clip_rect.p0.x + r_bl.x, clip_rect.p1.y - r_bl.y,
                                        wr_inverse_radii_squared(r_bl));
      vec2_scalar n_tl = -r_tl.sel(Y, X);
      vec2_scalar n_tr = vec2_scalar(r_tr.y, -r_tr.x);
      vec2_scalar n_br = r_br.sel(Y, X);
      vec2_scalar n_bl = vec2_scalar(-r_bl.y, r_bl.x);
      vClipPlane_TL = vec3_scalar(n_tl.x, n_tl.y, dot(n_tl, vec2_scalar(clip_rect.p0.x, clip_rect.p0.y + r_tl.y)));
      vClipPlane_TR = vec3_scalar(n_tr.x, n_tr.y, dot(n_tr, vec2_scalar(clip_rect.p1.x - r_tr.x, clip_rect.p0.y)));
      vClipPlane_BR = vec3_scalar(n_br.x, n_br.y, dot(n_br, vec2_scalar(clip_rect.p1.x, clip_rect.p1.y - r_br.y)));
      vClipPlane_BL = vec3_scalar(n_bl.x, n_bl.y, dot(n_bl, vec2_scalar(clip_rect.p0.x + r_bl.x, clip_rect.p1.y)));
    }
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vLocalPos = get_nth(vLocalPos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_clip_rectangle)
};

template <bool FAST>
struct cs_clip_rectangle_frag_t : FragmentShaderImpl, cs_clip_rectangle_vert_t<FAST> {
  typedef cs_clip_rectangle_frag_t Self;
  typedef typename cs_clip_rectangle_vert_t<FAST>::InterpOutputs InterpInputs;
  typedef typename cs_clip_rectangle_vert_t<FAST>::InterpOutputs InterpOutputs;

  vec4 vLocalPos;
  InterpInputs interp_step;

  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vLocalPos = init_interp(init->vLocalPos, step->vLocalPos);
    self->interp_step.vLocalPos = step->vLocalPos * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vLocalPos += interp_step.vLocalPos * chunks;
  }

  // shared.glsl:145-148 (SWGL), 184-189
  static float compute_aa_range(vec2 position) { return recip(fwidth(position).x); }
  static Float distance_aa(float aa_range, Float signed_distance) {
    Float dist = signed_distance * aa_range;
    return clamp(0.5f - dist, Float(0.0f), Float(1.0f));
  }
  // cs_clip_rectangle.glsl:160-168
  static Float sd_box(vec2 pos, vec2_scalar box_size) {
    vec2 d = abs(pos) - box_size;
    return length(max(d, Float(0.0f))) + min(max(d.x, d.y), Float(0.0f));
  }
  static Float sd_rounded_box(vec2 pos, vec2_scalar box_size, float radius) {
    return sd_box(pos, box_size) - radius;
  }
  // ellipse.glsl:14-19
  static Float distance_to_ellipse_approx(vec2 p, vec2 inv_radii_sq, float scale) {
    vec2 p_r = p * inv_radii_sq;
    Float g = dot(p, p_r) - scale;
    vec2 dG = (1.0f + scale) * p_r;
    return g * inversesqrt(dot(dG, dG));
  }
  static Float signed_distance_rect(vec2 pos, vec2_scalar p0, vec2_scalar p1) {  // rect.glsl
    vec2 d = max(p0 - pos, pos - p1);
    return max(d.x, d.y);
  }
  // ellipse.glsl:48-92
  Float distance_to_rounded_rect(vec2 pos) {
    vec4 corner = vec4(vec4_scalar(1.0e-6f, 1.0e-6f, 1.0f, 1.0f));
    vec4 crtl = vec4(this->vClipCenter_Radius_TL), crtr = vec4(this->vClipCenter_Radius_TR);
    vec4 crbr = vec4(this->vClipCenter_Radius_BR), crbl = vec4(this->vClipCenter_Radius_BL);
    vec3_scalar ptl = this->vClipPlane_TL, ptr_ = this->vClipPlane_TR, pbr = this->vClipPlane_BR,
                pbl = this->vClipPlane_BL;
    vec2 t;
    t = crtl.sel(X, Y) - pos; crtl.x = t.x; crtl.y = t.y;
    t = (crtr.sel(X, Y) - pos) * vec2_scalar(-1.0f, 1.0f); crtr.x = t.x; crtr.y = t.y;
    t = pos - crbr.sel(X, Y); crbr.x = t.x; crbr.y = t.y;
    t = (crbl.sel(X, Y) - pos) * vec2_scalar(1.0f, -1.0f); crbl.x = t.x; crbl.y = t.y;
    corner = if_then_else(dot(pos, vec2(ptl.sel(X, Y))) > ptl.z, crtl, corner);
    corner = if_then_else(dot(pos, vec2(ptr_.sel(X, Y))) > ptr_.z, crtr, corner);
    corner = if_then_else(dot(pos, vec2(pbr.sel(X, Y))) > pbr.z, crbr, corner);
    corner = if_then_else(dot(pos, vec2(pbl.sel(X, Y))) > pbl.z, crbl, corner);
    return max(distance_to_ellipse_approx(corner.sel(X, Y), corner.sel(Z, W), 1.0f),
               signed_distance_rect(pos, this->vTransformBounds.sel(X, Y), this->vTransformBounds.sel(Z, W)));
  }

  // cs_clip_rectangle.glsl:170-199
  void main() {
    vec2 local_pos = vLocalPos.sel(X, Y) / vLocalPos.w;
    float aa_range = compute_aa_range(local_pos);
    Float dist;
    if (FAST) {
      dist = sd_rounded_box(local_pos, this->vClipParams.sel(X, Y), this->vClipParams.z);
    } else {
      dist = distance_to_rounded_rect(local_pos);
    }
    Float alpha = distance_aa(aa_range, dist);
    Float final_alpha = mix(alpha, 1.0f - alpha, Float(this->vClipMode.x));
    Float final_final_alpha = if_then_else(vLocalPos.w > 0.0f, final_alpha, Float(0.0f));
    this->gl_FragColor = vec4(final_final_alpha, Float(0.0f), Float(0.0f), Float(1.0f));
  }

  // cs_clip_rectangle.glsl:223-495
  void swgl_drawSpanR8() {
    if (interp_step.vLocalPos.w != 0.0f) return;
    float w = swgl_forceScalar(vLocalPos.w);
    if (w <= 0.0f) {
      swgl_commitSolidR8(0.0f);
      return;
    }
    w = 1.0f / w;
    vec2 local_pos = vLocalPos.sel(X, Y) * w;
    vec2_scalar local_pos0 = swgl_forceScalar(local_pos);
    vec2_scalar local_step = interp_step.vLocalPos.sel(X, Y) * w;
    float step_scale = max(dot(local_step, local_step), 1.0e-6f);
    float aa_range = compute_aa_range(local_pos);
    float aa_margin = inversesqrt(aa_range * aa_range * step_scale);
    vec4_scalar clip_rect;
    if (FAST) {
      vec3_scalar cp = this->vClipParams;
      clip_rect = make_vec4(-cp.sel(X, Y) - cp.z, cp.sel(X, Y) + cp.z);
    } else {
      clip_rect = this->vTransformBounds;
    }
    bvec2_scalar neg = lessThan(local_step, vec2_scalar(0.0f));
    vec4_scalar clip_dist = mix(clip_rect, clip_rect.sel(Z, W, X, Y), neg.sel(X, Y, X, Y)) -
                            local_pos0.sel(X, Y, X, Y);
    bvec2_scalar ne = bvec2_scalar(local_step.x != 0.0f, local_step.y != 0.0f);
    clip_dist = mix(1.0e6f * step(vec4_scalar(0.0f), clip_dist),
                    clip_dist * recip(local_step).sel(X, Y, X, Y), ne.sel(X, Y, X, Y));
    float opaque_start = max(clip_dist.x, clip_dist.y);
    float opaque_end = min(clip_dist.z, clip_dist.w);
    float aa_start = opaque_start;
    float aa_end = opaque_end;
    vec3_scalar start_plane = vec3_scalar(1.0e6f);
    vec3_scalar end_plane = vec3_scalar(1.0e6f);
    vec4_scalar start_corner = vec4_scalar(1.0e6f, 1.0e6f, 1.0f, 1.0f);
    vec4_scalar end_corner = vec4_scalar(1.0e6f, 1.0e6f, 1.0f, 1.0f);
#define WR_CLIP_CORNER(plane, info)                                                 \
  do {                                                                              \
    float dist = dot(local_pos0, plane.sel(X, Y)) - plane.z;                        \
    float scale = -dot(local_step, plane.sel(X, Y));                                \
    if (scale >= 0.0f) {                                                            \
      if (dist > opaque_start * scale) {                                            \
        start_corner = info;                                                        \
        start_plane = plane;                                                        \
        float inv_scale = recip(max(scale, 1.0e-6f));                               \
        opaque_start = dist * inv_scale;                                            \
        float apex = (0.7071f - 0.5f) * 2.0f * abs(plane.x * plane.y);              \
        aa_start = opaque_start - apex * inv_scale;                                 \
      }                                                                             \
    } else if (dist > opaque_end * scale) {                                         \
      end_corner = info;                                                            \
      end_plane = plane;                                                            \
      float inv_scale = recip(min(scale, -1.0e-6f));                                \
      opaque_end = dist * inv_scale;                                                \
      float apex = (0.7071f - 0.5f) * 2.0f * abs(plane.x * plane.y);                \
      aa_end = opaque_end - apex * inv_scale;                                       \
    }                                                                               \
  } while (false)
    if (FAST) {
      vec3_scalar cp = this->vClipParams;
      float offset = (cp.x + cp.y + cp.z) * cp.z;
      vec3_scalar plane_tl = vec3_scalar(-cp.z, -cp.z, offset);
      vec3_scalar plane_tr = vec3_scalar(cp.z, -cp.z, offset);
      vec3_scalar plane_br = vec3_scalar(cp.z, cp.z, offset);
      vec3_scalar plane_bl = vec3_scalar(-cp.z, cp.z, offset);
      vec4_scalar none = start_corner;
      WR_CLIP_CORNER(plane_tl, none);
      WR_CLIP_CORNER(plane_tr, none);
      WR_CLIP_CORNER(plane_br, none);
      WR_CLIP_CORNER(plane_bl, none);
    } else {
      WR_CLIP_CORNER(this->vClipPlane_TL, this->vClipCenter_Radius_TL);
      WR_CLIP_CORNER(this->vClipPlane_TR, this->vClipCenter_Radius_TR);
      WR_CLIP_CORNER(this->vClipPlane_BR, this->vClipCenter_Radius_BR);
      WR_CLIP_CORNER(this->vClipPlane_BL, this->vClipCenter_Radius_BL);
    }
#undef WR_CLIP_CORNER
    aa_margin = max(aa_margin - max(aa_start - aa_end, 0.0f), 0.0f);
    aa_start -= aa_margin;
    aa_end += aa_margin;
    vec4_scalar stepsf = clamp(
        float(this->swgl_SpanLength) -
            float(swgl_StepSize) * vec4_scalar(floor(aa_start), ceil(opaque_start), floor(opaque_end), ceil(aa_end)),
        0.0f, float(this->swgl_SpanLength));
    int aa_start_len = int(stepsf.x), opaque_start_len = int(stepsf.y), opaque_end_len = int(stepsf.z),
        aa_end_len = int(stepsf.w);
    float mode = this->vClipMode.x;
    auto AA_RECT = [&](vec2 lp) -> Float {
      if (FAST) return sd_rounded_box(lp, this->vClipParams.sel(X, Y), this->vClipParams.z);
      return signed_distance_rect(lp, this->vTransformBounds.sel(X, Y), this->vTransformBounds.sel(Z, W));
    };
    auto AA_CORNER = [&](vec2 lp, vec4_scalar corner) -> Float {
      return distance_to_ellipse_approx(lp - corner.sel(X, Y), vec2(corner.sel(Z, W)), 1.0f);
    };
    if (this->swgl_SpanLength > aa_start_len) {
      int num_aa = this->swgl_SpanLength - aa_start_len;
      swgl_commitPartialSolidR8(num_aa, mode);
      local_pos += float(num_aa / swgl_StepSize) * local_step;
    }
    if (!FAST && start_plane.x < 1.0e5f) {
      while (this->swgl_SpanLength > opaque_start_len) {
        Float alpha = distance_aa(aa_range,
                                  if_then_else(dot(local_pos, vec2(start_plane.sel(X, Y))) > start_plane.z,
                                               AA_CORNER(local_pos, start_corner), AA_RECT(local_pos)));
        swgl_commitColorR8(mix(alpha, 1.0f - alpha, Float(mode)));
        local_pos += local_step;
      }
    }
    while (this->swgl_SpanLength > opaque_start_len) {
      Float alpha = distance_aa(aa_range, AA_RECT(local_pos));
      swgl_commitColorR8(mix(alpha, 1.0f - alpha, Float(mode)));
      local_pos += local_step;
    }
    if (this->swgl_SpanLength > opaque_end_len) {
      int num_opaque = this->swgl_SpanLength - opaque_end_len;
      swgl_commitPartialSolidR8(num_opaque, 1.0f - mode);
      local_pos += float(num_opaque / swgl_StepSize) * local_step;
    }
    if (!FAST && end_plane.x < 1.0e5f) {
      while (this->swgl_SpanLength > aa_end_len) {
        Float alpha = distance_aa(aa_range,
                                  if_then_else(dot(local_pos, vec2(end_plane.sel(X, Y))) > end_plane.z,
                                               AA_CORNER(local_pos, end_corner), AA_RECT(local_pos)));
        swgl_commitColorR8(mix(alpha, 1.0f - alpha, Float(mode)));
        local_pos += local_step;
      }
    }
    while (this->swgl_SpanLength > aa_end_len) {
      Float alpha = distance_aa(aa_range, AA_RECT(local_pos));
      swgl_commitColorR8(mix(alpha, 1.0f - alpha, Float(mode)));
      local_pos += local_step;
    }
    if (this->swgl_SpanLength > 0) {
      swgl_commitPartialSolidR8(this->swgl_SpanLength, mode);
    }
  }
  static int draw_span_R8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, R8);
  }
  WR_FRAGMENT_ABI()
  cs_clip_rectangle_frag_t() {
    this->init_fragment_abi();
    this->draw_span_R8_func = &draw_span_R8;
  }
};

typedef cs_clip_rectangle_frag_t<false> cs_clip_rectangle_frag;
typedef cs_clip_rectangle_frag_t<true> cs_clip_rectangle_FAST_PATH_frag;
WR_PROGRAM(cs_clip_rectangle, "cs_clip_rectangle")
WR_PROGRAM(cs_clip_rectangle_FAST_PATH, "cs_clip_rectangle FAST_PATH")
