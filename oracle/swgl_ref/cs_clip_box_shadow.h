// TEST INFRASTRUCTURE — hand-instantiated SWGL program "cs_clip_box_shadow
// TEXTURE_2D": webrender/res/cs_clip_box_shadow.glsl + clip_shared.glsl +
// transform.glsl (rectangle_aa_rough_fragment) restated in the glsl.h
// vocabulary.  Needs ClipVertBase from cs_clip_rectangle.h.
#pragma once

struct cs_clip_box_shadow_vert : ClipVertBase {
  typedef cs_clip_box_shadow_vert Self;
  ivec2_scalar aClipDataResourceAddress;
  vec2_scalar aClipSrcRectSize;
  int aClipMode;
  ivec2_scalar aStretchMode;
  vec4_scalar aClipDestRect;
  int a_loc[5];

  // outputs
  vec4 vLocalPos;
  vec2 vUv;
  vec4_scalar vUvBounds, vEdge, vUvBounds_NoClamp;
  vec2_scalar vClipMode;

  struct InterpOutputs {
    vec4_scalar vLocalPos;
    vec2_scalar vUv;
  };

  cs_clip_box_shadow_vert() {
    static const char* names[5] = {"aClipDataResourceAddress", "aClipSrcRectSize", "aClipMode", "aStretchMode",
                                   "aClipDestRect"};
    for (int i = 0; i < 5; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_Color0;
    init_vertex_abi();
  }

  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    self->load_common_attribs(attribs, start, instance, count);
    auto& L = self->attrib_locations.locs;
    load_flat_attrib(self->aClipDataResourceAddress, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aClipSrcRectSize, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aClipMode, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aStretchMode, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aClipDestRect, attribs[L[self->a_loc[4]]], start, instance, count);
  }

  // cs_clip_box_shadow.glsl:59-124
  void main() {
    Transform clip_transform = fetch_transform(aTransformIds.x);
    Transform prim_transform = fetch_transform(aTransformIds.y);
    // fetch_image_source_direct (gpu_cache.glsl:111-115): texel address given as (x, y)
    vec4_scalar res0 = texelFetch(sGpuCache, aClipDataResourceAddress, 0);
    RectWithEndpoint dest_rect = RectWithEndpoint{aClipDestRect.sel(X, Y), aClipDestRect.sel(Z, W)};
    ClipVertexInfo vi = write_clip_tile_vertex(
        dest_rect, prim_transform, clip_transform,
        RectWithEndpoint{aClipDeviceArea.sel(X, Y), aClipDeviceArea.sel(Z, W)}, aClipOrigins.sel(X, Y),
        aClipOrigins.sel(Z, W), aDevicePixelScale);
    vClipMode.x = float(aClipMode);
    vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
    vec2 local_pos = vi.local_pos.sel(X, Y) / vi.local_pos.w;
    vLocalPos = vi.local_pos;
    vec2_scalar dest_rect_size = dest_rect.p1 - dest_rect.p0;
    switch (aStretchMode.x) {
      case 0:
        vEdge.x = 0.5f;
        vEdge.z = (dest_rect_size.x / aClipSrcRectSize.x) - 0.5f;
        vUv.x = (local_pos.x - dest_rect.p0.x) / aClipSrcRectSize.x;
        break;
      case 1:
      default:
        vEdge.x = 1.0f;
        vEdge.z = 1.0f;
        vUv.x = (local_pos.x - dest_rect.p0.x) / dest_rect_size.x;
        break;
    }
    switch (aStretchMode.y) {
      case 0:
        vEdge.y = 0.5f;
        vEdge.w = (dest_rect_size.y / aClipSrcRectSize.y) - 0.5f;
        vUv.y = (local_pos.y - dest_rect.p0.y) / aClipSrcRectSize.y;
        break;
      case 1:
      default:
        vEdge.y = 1.0f;
        vEdge.w = 1.0f;
        vUv.y = (local_pos.y - dest_rect.p0.y) / dest_rect_size.y;
        break;
    }
    vUv *= vi.local_pos.w;
    vec2_scalar uv0 = res0.sel(X, Y);
    vec2_scalar uv1 = res0.sel(Z, W);
    vUvBounds = make_vec4(uv0 + vec2_scalar(0.5f), uv1 - vec2_scalar(0.5f)) / texture_size.sel(X, Y, X, Y);
    vUvBounds_NoClamp = make_vec4(uv0, uv1) / texture_size.sel(X, Y, X, Y);
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vLocalPos = get_nth(vLocalPos, n);
      dest->vUv = get_nth(vUv, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_clip_box_shadow)
};

struct cs_clip_box_shadow_frag : FragmentShaderImpl, cs_clip_box_shadow_vert {
  typedef cs_clip_box_shadow_frag Self;
  typedef cs_clip_box_shadow_vert::InterpOutputs InterpInputs;
  typedef cs_clip_box_shadow_vert::InterpOutputs InterpOutputs;

  vec4 vLocalPos;
  vec2 vUv;
  InterpInputs interp_step;

  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vLocalPos = init_interp(init->vLocalPos, step->vLocalPos);
    self->interp_step.vLocalPos = step->vLocalPos * 4.0f;
    self->vUv = init_interp(init->vUv, step->vUv);
    self->interp_step.vUv = step->vUv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vLocalPos += interp_step.vLocalPos * chunks;
    vUv += interp_step.vUv * chunks;
  }

  // rect.glsl point_inside_rect + transform.glsl:132-138
  Float rectangle_aa_rough_fragment(vec2 local_pos) {
    vec2_scalar p0 = this->vTransformBounds.sel(X, Y), p1 = this->vTransformBounds.sel(Z, W);
    vec2 s = step(vec2(p0), local_pos) - step(vec2(p1), local_pos);
    return s.x * s.y;
  }

  vec2 map_uv(vec2 uv_linear) {
    vec2 uv = clamp(uv_linear, vec2(vec2_scalar(0.0f)), vec2(this->vEdge.sel(X, Y)));
    uv += max(vec2(vec2_scalar(0.0f)), uv_linear - vec2(this->vEdge.sel(Z, W)));
    uv = mix(vec2(this->vUvBounds_NoClamp.sel(X, Y)), vec2(this->vUvBounds_NoClamp.sel(Z, W)), uv);
    return uv;
  }

  // cs_clip_box_shadow.glsl:122-138
  void main() {
    vec2 uv_linear = vUv / vLocalPos.w;
    vec2 uv = map_uv(uv_linear);
    uv = clamp(uv, vec2(this->vUvBounds.sel(X, Y)), vec2(this->vUvBounds.sel(Z, W)));
    Float in_shadow_rect = rectangle_aa_rough_fragment(vLocalPos.sel(X, Y) / vLocalPos.w);
    Float texel = texture(this->sColor0, uv).x;
    Float alpha = mix(texel, 1.0f - texel, Float(this->vClipMode.x));
    Float result = if_then_else(vLocalPos.w > 0.0f, mix(Float(this->vClipMode.x), alpha, in_shadow_rect),
                                Float(0.0f));
    this->gl_FragColor = vec4(result);
  }

  // cs_clip_box_shadow.glsl:150-323
  void swgl_drawSpanR8() {
    if (interp_step.vLocalPos.w != 0.0f) return;
    float w = swgl_forceScalar(vLocalPos.w);
    if (w <= 0.0f) {
      swgl_commitSolidR8(0.0f);
      return;
    }
    w = 1.0f / w;
    vec2 uv_linear = vUv * w;
    vec2_scalar uv_linear0 = swgl_forceScalar(uv_linear);
    vec2_scalar uv_linear_step = interp_step.vUv * w;
    vec2 local_pos = vLocalPos.sel(X, Y) * w;
    vec2_scalar local_pos0 = swgl_forceScalar(local_pos);
    vec2_scalar local_step = interp_step.vLocalPos.sel(X, Y) * w;

    vec4_scalar tb = this->vTransformBounds;
    bvec2_scalar neg = lessThan(local_step, vec2_scalar(0.0f));
    vec4_scalar clip_dist = mix(tb, tb.sel(Z, W, X, Y), neg.sel(X, Y, X, Y)) - local_pos0.sel(X, Y, X, Y);
    bvec2_scalar ne = bvec2_scalar(local_step.x != 0.0f, local_step.y != 0.0f);
    clip_dist = mix(1.0e6f * step(vec4_scalar(0.0f), clip_dist), clip_dist * recip(local_step).sel(X, Y, X, Y),
                    ne.sel(X, Y, X, Y));
    float shadow_start = max(clip_dist.x, clip_dist.y);
    float shadow_end = min(clip_dist.z, clip_dist.w);
    vec2_scalar ssf = clamp(float(this->swgl_SpanLength) -
                                float(swgl_StepSize) * vec2_scalar(floor(shadow_start), ceil(shadow_end)),
                            0.0f, float(this->swgl_SpanLength));
    int shadow_start_len = int(ssf.x), shadow_end_len = int(ssf.y);

    vec4_scalar edge = this->vEdge;
    bvec2_scalar uneg = lessThan(uv_linear_step, vec2_scalar(0.0f));
    vec4_scalar opaque_dist = mix(edge, edge.sel(Z, W, X, Y), uneg.sel(X, Y, X, Y)) - uv_linear0.sel(X, Y, X, Y);
    bvec2_scalar une = bvec2_scalar(uv_linear_step.x != 0.0f, uv_linear_step.y != 0.0f);
    opaque_dist = mix(1.0e6f * step(vec4_scalar(0.0f), opaque_dist),
                      opaque_dist * recip(uv_linear_step).sel(X, Y, X, Y), une.sel(X, Y, X, Y));
    vec4_scalar osf = clamp(float(this->swgl_SpanLength) -
                                float(swgl_StepSize) * vec4_scalar(floor(opaque_dist.x), floor(opaque_dist.y),
                                                                   floor(opaque_dist.z), floor(opaque_dist.w)),
                            float(shadow_end_len), float(this->swgl_SpanLength));
    int os_x = int(osf.x), os_y = int(osf.y), os_z = int(osf.z), os_w = int(osf.w);

    float mode = this->vClipMode.x;
    if (this->swgl_SpanLength > shadow_start_len) {
      int num_before = this->swgl_SpanLength - shadow_start_len;
      swgl_commitPartialSolidR8(num_before, mode);
      float steps_before = float(num_before / swgl_StepSize);
      uv_linear += steps_before * uv_linear_step;
      local_pos += steps_before * local_step;
    }
    while (this->swgl_SpanLength > 0) {
      {
        vec2 uv = map_uv(uv_linear);
        uv = clamp(uv, vec2(this->vUvBounds.sel(X, Y)), vec2(this->vUvBounds.sel(Z, W)));
        Float in_shadow_rect = rectangle_aa_rough_fragment(local_pos);
        Float texel = texture(this->sColor0, uv).x;
        Float alpha = mix(texel, 1.0f - texel, Float(mode));
        Float result = mix(Float(mode), alpha, in_shadow_rect);
        swgl_commitColorR8(result);
        uv_linear += uv_linear_step;
        local_pos += local_step;
      }
      if (this->swgl_SpanLength <= shadow_end_len) break;
      int num_inside = this->swgl_SpanLength - swgl_StepSize - shadow_end_len;
      vec4_scalar uv_bounds = this->vUvBounds;
      vec4_scalar nc = this->vUvBounds_NoClamp;
      if (this->swgl_SpanLength >= os_y) {
        num_inside = min(num_inside, this->swgl_SpanLength - os_y);
      } else if (this->swgl_SpanLength >= os_w) {
        num_inside = min(num_inside, this->swgl_SpanLength - os_w);
        float c = clamp(mix(nc.y, nc.w, edge.y), this->vUvBounds.y, this->vUvBounds.w);
        uv_bounds.y = c;
        uv_bounds.w = c;
      }
      if (this->swgl_SpanLength >= os_x) {
        num_inside = min(num_inside, this->swgl_SpanLength - os_x);
      } else if (this->swgl_SpanLength >= os_z) {
        num_inside = min(num_inside, this->swgl_SpanLength - os_z);
        float c = clamp(mix(nc.x, nc.z, edge.x), this->vUvBounds.x, this->vUvBounds.z);
        uv_bounds.x = c;
        uv_bounds.z = c;
      }
      if (num_inside > 0) {
        vec2 uv = map_uv(uv_linear);
        if (uv_bounds.sel(X, Y) == uv_bounds.sel(Z, W)) {
          uv = clamp(uv, vec2(uv_bounds.sel(X, Y)), vec2(uv_bounds.sel(Z, W)));
          float texel = swgl_forceScalar(texture(this->sColor0, uv).x);
          float alpha = mix(texel, 1.0f - texel, mode);
          swgl_commitPartialSolidR8(num_inside, alpha);
        } else if (mode != 0.0f) {
          swgl_commitPartialTextureLinearInvertR8(num_inside, this->sColor0, uv, uv_bounds);
        } else {
          swgl_commitPartialTextureLinearR8(num_inside, this->sColor0, uv, uv_bounds);
        }
        float steps_inside = float(num_inside / swgl_StepSize);
        uv_linear += steps_inside * uv_linear_step;
        local_pos += steps_inside * local_step;
      }
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
  cs_clip_box_shadow_frag() {
    this->init_fragment_abi();
    this->draw_span_R8_func = &draw_span_R8;
  }
};
typedef cs_clip_box_shadow_frag cs_clip_box_shadow_TEXTURE_2D_frag;
WR_PROGRAM(cs_clip_box_shadow_TEXTURE_2D, "cs_clip_box_shadow TEXTURE_2D")
