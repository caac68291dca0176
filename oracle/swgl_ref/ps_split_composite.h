// TEST INFRASTRUCTURE — hand-instantiated SWGL program "ps_split_composite"
// (webrender/res/ps_split_composite.glsl; instances = SplitCompositeInstance,
// gpu_types.rs:531-552): a plane-split polygon of a preserve-3d picture — four
// local points from the GPU cache, transformed (usually with perspective) and
// textured from the picture's surface.  SWGL branches: clip mask through
// swgl_clipMask, span shader swgl_commitTextureRGBA8.
#pragma once

struct ps_split_composite_vert : PrimVertBase {
  typedef ps_split_composite_vert Self;
  vec2 vUv;
  vec2_scalar vPerspective;
  vec4_scalar vUvSampleBounds;
  struct InterpOutputs {
    vec2_scalar vUv;
  };

  ps_split_composite_vert() {
    this->sampler_mask |= WR_S_Color0;
    this->init_vertex_abi();
  }

  static vec2 bilerp(vec2_scalar a, vec2_scalar b, vec2_scalar c, vec2_scalar d, Float s, Float t) {
    vec2 x = mix(vec2(a), vec2(b), t);
    vec2 y = mix(vec2(c), vec2(d), t);
    return mix(x, y, s);
  }

  // ps_split_composite.glsl:64-118
  void main() {
    int prim_header_index = aData.x, polygons_address = aData.y, render_task_index = aData.w;
    float ci_z = float(aData.z);
    vec4_scalar data0 = this->fetch_gpu_cache(polygons_address, 0);
    vec4_scalar data1 = this->fetch_gpu_cache(polygons_address, 1);
    vec2_scalar local[4] = {data0.sel(X, Y), data0.sel(Z, W), data1.sel(X, Y), data1.sel(Z, W)};
    PrimitiveHeader ph = fetch_prim_header(prim_header_index);
    PictureTask dest_task = fetch_picture_task(render_task_index);
    Transform transform = fetch_transform(ph.transform_id);
    vec4_scalar res0 = this->fetch_gpu_cache(ph.user_data.x, 0);
    RectWithEndpoint res_uv_rect = RectWithEndpoint{res0.sel(X, Y), res0.sel(Z, W)};
    ClipArea clip_area = fetch_clip_area(ph.user_data.w);

    vec2_scalar dest_origin = dest_task.task_rect.p0 - dest_task.content_origin;
    vec2 local_pos = bilerp(local[0], local[1], local[3], local[2], aPosition.y, aPosition.x);
    vec4 world_pos = transform.m * vec4(local_pos, Float(0.0f), Float(1.0f));
    vec4 final_pos = vec4(vec2(dest_origin) * world_pos.w + world_pos.sel(X, Y) * Float(dest_task.device_pixel_scale),
                          world_pos.w * ci_z, world_pos.w);
    write_clip(clip_area, dest_task);
    gl_Position = uTransform * final_pos;

    vec2_scalar texture_size = make_vec2(textureSize(this->sColor0, 0));
    vec2_scalar uv0 = res_uv_rect.p0;
    vec2_scalar uv1 = res_uv_rect.p1;
    vec2_scalar min_uv = min(uv0, uv1);
    vec2_scalar max_uv = max(uv0, uv1);
    vUvSampleBounds = make_vec4(min_uv + make_vec2(0.5f), max_uv - make_vec2(0.5f)) / texture_size.sel(X, Y, X, Y);
    vec2 f = (local_pos - vec2(ph.local_rect.p0)) / vec2(ph.local_rect.p1 - ph.local_rect.p0);
    {
      // get_image_quad_uv (prim_shared.glsl:202-210)
      vec4_scalar st_tl = this->fetch_gpu_cache(ph.user_data.x + 2, 0);
      vec4_scalar st_tr = this->fetch_gpu_cache(ph.user_data.x + 2, 1);
      vec4_scalar st_bl = this->fetch_gpu_cache(ph.user_data.x + 2, 2);
      vec4_scalar st_br = this->fetch_gpu_cache(ph.user_data.x + 2, 3);
      vec4 x = mix(st_tl, st_tr, f.x);
      vec4 y = mix(st_bl, st_br, f.x);
      vec4 z = mix(x, y, f.y);
      f = z.sel(X, Y) / z.w;
    }
    vec2 uv = mix(uv0, uv1, f);
    float perspective_interpolate = float(ph.user_data.y);
    vUv = uv / vec2(texture_size) * mix(gl_Position.w, Float(1.0f), Float(perspective_interpolate));
    vPerspective.x = perspective_interpolate;
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUv = get_nth(vUv, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(ps_split_composite)
};

struct ps_split_composite_frag : FragmentShaderImpl, ps_split_composite_vert {
  typedef ps_split_composite_frag Self;
  typedef ps_split_composite_vert::InterpOutputs InterpInputs;
  typedef ps_split_composite_vert::InterpOutputs InterpOutputs;
  vec2 vUv;
  InterpInputs interp_step;
  struct InterpPerspective {
    vec2 vUv;
  };
  InterpPerspective interp_perspective;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vUv = init_interp(init->vUv, step->vUv);
    self->interp_step.vUv = step->vUv * 4.0f;
  }
  static void read_perspective_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    Float w = 1.0f / self->gl_FragCoord.w;
    self->interp_perspective.vUv = init_interp(init->vUv, step->vUv);
    self->vUv = self->interp_perspective.vUv * w;
    self->interp_step.vUv = step->vUv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vUv += interp_step.vUv * chunks;
  }
  ALWAYS_INLINE void step_perspective_inputs(int steps = 4) {
    this->step_perspective(steps);
    float chunks = steps * 0.25f;
    Float w = 1.0f / this->gl_FragCoord.w;
    interp_perspective.vUv += interp_step.vUv * chunks;
    vUv = w * interp_perspective.vUv;
  }

  // ps_split_composite.glsl:121-127 (do_clip() == 1 under SWGL_CLIP_MASK)
  void main() {
    Float alpha = 1.0f;
    Float perspective_divisor = mix(this->gl_FragCoord.w, Float(1.0f), Float(this->vPerspective.x));
    vec2 uv = clamp(vUv * perspective_divisor, vec2(this->vUvSampleBounds.sel(X, Y)), vec2(this->vUvSampleBounds.sel(Z, W)));
    this->gl_FragColor = alpha * texture(this->sColor0, uv);
  }

  // ps_split_composite.glsl:129-136
  void swgl_drawSpanRGBA8() {
    float perspective_divisor = mix(swgl_forceScalar(this->gl_FragCoord.w), 1.0f, this->vPerspective.x);
    vec2 uv = vUv * perspective_divisor;
    swgl_commitTextureRGBA8(this->sColor0, uv, this->vUvSampleBounds);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI_W()
  ps_split_composite_frag() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

WR_PROGRAM(ps_split_composite, "ps_split_composite")
