// TEST INFRASTRUCTURE — hand-instantiated SWGL programs
//   "ps_text_run ALPHA_PASS,GLYPH_TRANSFORM,TEXTURE_2D" and
//   "ps_text_run ALPHA_PASS,DUAL_SOURCE_BLENDING,GLYPH_TRANSFORM,TEXTURE_2D"
// (webrender/res/ps_text_run.glsl WITH WR_FEATURE_GLYPH_TRANSFORM: glyphs rasterised in the
// transformed space; the quad is trimmed to the glyph rect with gl_ClipDistance under SWGL_CLIP_DIST).
// DUAL: 1 for the DUAL_SOURCE_BLENDING variant (under SWGL_BLEND only the
// swizzle line and the span commit differ).
#pragma once

template <int DUAL>
struct ps_text_run_gt_vert_t : PrimVertBase {
  typedef ps_text_run_gt_vert_t Self;
  vec4_scalar v_color;
  vec3_scalar v_mask_swizzle;
  vec4_scalar v_uv_bounds;
  vec2 v_uv;
  struct InterpOutputs {
    vec4_scalar swgl_ClipDistance;  // clip distances travel in the first SIMD chunk (program.h:15-19)
    vec2_scalar v_uv;
  };
  ps_text_run_gt_vert_t() {
    sampler_mask |= WR_S_Color0;
    init_vertex_abi();
    enable_clip_distance();
  }
  // ps_text_run.glsl:26-35
  static RectWithEndpoint transform_rect(RectWithEndpoint rect, mat2_scalar transform) {
    vec2_scalar size = rect.p1 - rect.p0;
    vec2_scalar center = transform * (rect.p0 + size * 0.5f);
    vec2_scalar radius = mat2_scalar(abs(transform[0]), abs(transform[1])) * (size * 0.5f);
    return RectWithEndpoint{center - radius, center + radius};
  }
  static bool rect_inside_rect(RectWithEndpoint little, RectWithEndpoint big) {
    return big.p0.x <= little.p0.x && big.p0.y <= little.p0.y && little.p1.x <= big.p1.x && little.p1.y <= big.p1.y;
  }

  // ps_text_run.glsl:98-264
  void main() {
    Instance instance = decode_instance_attributes();
    PrimitiveHeader ph = fetch_prim_header(instance.prim_header_address);
    Transform transform = fetch_transform(ph.transform_id);
    ClipArea clip_area = fetch_clip_area(instance.clip_address);
    PictureTask task = fetch_picture_task(ph.picture_task_address);
    int glyph_index = instance.segment_index;
    int subpx_dir = (instance.flags >> 8) & 0xff;
    int color_mode = instance.flags & 0xff;
    vec4_scalar text_color = fetch_from_gpu_cache_1(ph.specific_prim_address);
    vec2_scalar text_offset = ph.local_rect.p1;
    // fetch_glyph
    int glyph_address = ph.specific_prim_address + 1 + int(uint32_t(glyph_index) / 2U);
    vec4_scalar data = fetch_from_gpu_cache_1(glyph_address);
    vec2_scalar glyph_offset = (uint32_t(glyph_index) % 2U == 1U) ? data.sel(Z, W) : data.sel(X, Y);
    glyph_offset += ph.local_rect.p0;
    // fetch_glyph_resource
    vec4_scalar res_uv_rect = fetch_gpu_cache(instance.resource_address, 0);
    vec4_scalar res1 = fetch_gpu_cache(instance.resource_address, 1);
    vec2_scalar res_offset = res1.sel(X, Y);
    float res_scale = res1.z;
    vec2_scalar snap_bias;
    switch (subpx_dir) {
      case 1: snap_bias = vec2_scalar(0.125f, 0.5f); break;
      case 2: snap_bias = vec2_scalar(0.5f, 0.125f); break;
      case 3: snap_bias = vec2_scalar(0.125f); break;
      default: snap_bias = vec2_scalar(0.5f); break;
    }
    // WR_FEATURE_GLYPH_TRANSFORM, ps_text_run.glsl:129-167
    mat2_scalar glyph_transform = mat2_scalar(transform.m) * task.device_pixel_scale;
    vec2_scalar glyph_translation = transform.m[3].sel(X, Y) * task.device_pixel_scale;
    mat2_scalar glyph_transform_inv = inverse(glyph_transform);
    vec2_scalar raster_glyph_offset = floor(glyph_transform * glyph_offset + snap_bias);
    vec2_scalar raster_text_offset = floor(glyph_transform * text_offset + glyph_translation + 0.5f) - glyph_translation;
    vec2_scalar glyph_origin = res_offset + raster_glyph_offset + raster_text_offset;
    RectWithEndpoint glyph_rect =
        RectWithEndpoint{glyph_origin, glyph_origin + res_uv_rect.sel(Z, W) - res_uv_rect.sel(X, Y)};
    RectWithEndpoint local_rect = transform_rect(glyph_rect, glyph_transform_inv);
    vec2 local_pos = mix(local_rect.p0, local_rect.p1, aPosition);
    if (rect_inside_rect(local_rect, ph.local_clip_rect)) {
      local_pos = glyph_transform_inv * mix(glyph_rect.p0, glyph_rect.p1, aPosition);
    }
    VertexInfo vi = write_vertex(local_pos, ph.local_clip_rect, ph.z, transform, task);
    vec2 f = (glyph_transform * vi.local_pos - vec2(glyph_rect.p0)) / vec2(glyph_rect.p1 - glyph_rect.p0);
    gl_ClipDistance[0] = f.x;
    gl_ClipDistance[1] = f.y;
    gl_ClipDistance[2] = 1.0f - f.x;
    gl_ClipDistance[3] = 1.0f - f.y;
    write_clip(clip_area, task);
    switch (color_mode) {
      case 0:  // COLOR_MODE_ALPHA
        v_mask_swizzle = vec3_scalar(0.0f, 1.0f, 1.0f);
        v_color = text_color;
        break;
      case 2:  // COLOR_MODE_BITMAP_SHADOW
        swgl_blendDropShadow(text_color);
        v_mask_swizzle = vec3_scalar(1.0f, 0.0f, 0.0f);
        v_color = vec4_scalar(1.0f);
        break;
      case 3:  // COLOR_MODE_COLOR_BITMAP
        v_mask_swizzle = vec3_scalar(1.0f, 0.0f, 0.0f);
        v_color = vec4_scalar(text_color.w);
        break;
      case 1:  // COLOR_MODE_SUBPX_DUAL_SOURCE
        swgl_blendSubpixelText(text_color);
        v_mask_swizzle = vec3_scalar(1.0f, 0.0f, 0.0f);
        v_color = vec4_scalar(1.0f);
        break;
      default:
        v_mask_swizzle = vec3_scalar(0.0f, 0.0f, 0.0f);
        v_color = vec4_scalar(1.0f);
    }
    vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
    vec2_scalar st0 = res_uv_rect.sel(X, Y) / texture_size;
    vec2_scalar st1 = res_uv_rect.sel(Z, W) / texture_size;
    v_uv = mix(st0, st1, f);
    v_uv_bounds = (res_uv_rect + vec4_scalar(0.5f, 0.5f, -0.5f, -0.5f)) / texture_size.sel(X, Y, X, Y);
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->swgl_ClipDistance = vec4_scalar(get_nth(gl_ClipDistance[0], n), get_nth(gl_ClipDistance[1], n),
                                            get_nth(gl_ClipDistance[2], n), get_nth(gl_ClipDistance[3], n));
      dest->v_uv = get_nth(v_uv, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(ps_text_run_gt)
};

template <int DUAL>
struct ps_text_run_gt_frag_t : FragmentShaderImpl, ps_text_run_gt_vert_t<DUAL> {
  typedef ps_text_run_gt_frag_t Self;
  typedef typename ps_text_run_gt_vert_t<DUAL>::InterpOutputs InterpInputs;
  typedef typename ps_text_run_gt_vert_t<DUAL>::InterpOutputs InterpOutputs;
  vec2 v_uv;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->v_uv = init_interp(init->v_uv, step->v_uv);
    self->interp_step.v_uv = step->v_uv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    v_uv += interp_step.v_uv * chunks;
  }
  // ps_text_run.glsl:278-317
  void main() {
    vec2 tc = clamp(v_uv, vec2(this->v_uv_bounds.sel(X, Y)), vec2(this->v_uv_bounds.sel(Z, W)));
    vec4 mask = texture(this->sColor0, tc);
    if (this->v_mask_swizzle.z != 0.0f) mask = mask.sel(X, X, X, X);
    if (!DUAL) {
      vec3 rgb = mask.sel(X, Y, Z) * Float(this->v_mask_swizzle.x) + mask.sel(W, W, W) * Float(this->v_mask_swizzle.y);
      mask.x = rgb.x; mask.y = rgb.y; mask.z = rgb.z;
    }
    vec4 color = vec4(this->v_color) * mask;
    color *= Float(1.0f);  // do_clip()
    this->gl_FragColor = color;
  }
  // ps_text_run.glsl:321-337
  void swgl_drawSpanRGBA8() {
    if (this->v_mask_swizzle.x != 0.0f && this->v_mask_swizzle.x != 1.0f) return;
    if (DUAL) {
      swgl_commitTextureLinearRGBA8(this->sColor0, v_uv, this->v_uv_bounds);
    } else if (swgl_isTextureR8(this->sColor0)) {
      swgl_commitTextureLinearColorR8ToRGBA8(this->sColor0, v_uv, this->v_uv_bounds, this->v_color);
    } else {
      swgl_commitTextureLinearColorRGBA8(this->sColor0, v_uv, this->v_uv_bounds, this->v_color);
    }
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  ps_text_run_gt_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
typedef ps_text_run_gt_frag_t<0> ps_text_run_ALPHA_PASS_GLYPH_TRANSFORM_TEXTURE_2D_frag;
typedef ps_text_run_gt_frag_t<1> ps_text_run_ALPHA_PASS_DUAL_SOURCE_BLENDING_GLYPH_TRANSFORM_TEXTURE_2D_frag;
WR_PROGRAM(ps_text_run_ALPHA_PASS_GLYPH_TRANSFORM_TEXTURE_2D, "ps_text_run ALPHA_PASS,GLYPH_TRANSFORM,TEXTURE_2D")
WR_PROGRAM(ps_text_run_ALPHA_PASS_DUAL_SOURCE_BLENDING_GLYPH_TRANSFORM_TEXTURE_2D,
           "ps_text_run ALPHA_PASS,DUAL_SOURCE_BLENDING,GLYPH_TRANSFORM,TEXTURE_2D")
