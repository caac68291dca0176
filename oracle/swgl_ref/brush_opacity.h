// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_opacity",
// "brush_opacity ALPHA_PASS" and "brush_opacity ALPHA_PASS,ANTIALIASING"
// (webrender/res/brush_opacity.glsl; antialias_brush() == 1 under SWGL).
#pragma once

template <int VARIANT>
struct brush_opacity_vert_t : BrushVertBase<brush_opacity_vert_t<VARIANT>> {
  typedef brush_opacity_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 3;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2 v_uv;
  vec4_scalar v_uv_sample_bounds;
  vec2_scalar v_opacity_perspective_vec;
  struct InterpOutputs {
    vec2_scalar v_uv;
  };

  brush_opacity_vert_t() {
    this->sampler_mask |= WR_S_Color0;
    this->init_vertex_abi();
  }

  // brush_opacity.glsl:23-52
  void brush_vs(VertexInfo& vi, int, RectWithEndpoint local_rect, RectWithEndpoint, ivec4_scalar prim_user_data,
                int, mat4_scalar, PictureTask&, int brush_flags, vec4_scalar) {
    vec4_scalar r0 = this->fetch_gpu_cache(prim_user_data.x, 0);
    vec2_scalar uv0 = r0.sel(X, Y);
    vec2_scalar uv1 = r0.sel(Z, W);
    vec2_scalar texture_size = make_vec2(textureSize(this->sColor0, 0));
    vec2 f = (vi.local_pos - vec2(local_rect.p0)) / vec2(local_rect.p1 - local_rect.p0);
    {
      // get_image_quad_uv (prim_shared.glsl:204-210)
      vec4_scalar st_tl = this->fetch_gpu_cache(prim_user_data.x + 2, 0);
      vec4_scalar st_tr = this->fetch_gpu_cache(prim_user_data.x + 2, 1);
      vec4_scalar st_bl = this->fetch_gpu_cache(prim_user_data.x + 2, 2);
      vec4_scalar st_br = this->fetch_gpu_cache(prim_user_data.x + 2, 3);
      vec4 x = mix(st_tl, st_tr, f.x);
      vec4 y = mix(st_bl, st_br, f.x);
      vec4 z = mix(x, y, f.y);
      f = z.sel(X, Y) / z.w;
    }
    vec2 uv = mix(uv0, uv1, f);
    float perspective_interpolate = (brush_flags & WR_BRUSH_FLAG_PERSPECTIVE_INTERPOLATION) != 0 ? 1.0f : 0.0f;
    v_uv = uv / vec2(texture_size) * mix(vi.world_pos.w, Float(1.0f), Float(perspective_interpolate));
    v_opacity_perspective_vec.y = perspective_interpolate;
    v_uv_sample_bounds = make_vec4(uv0 + make_vec2(0.5f), uv1 - make_vec2(0.5f)) / texture_size.sel(X, Y, X, Y);
    v_opacity_perspective_vec.x = clamp(float(prim_user_data.y) / 65536.0f, 0.0f, 1.0f);
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_uv = get_nth(v_uv, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_opacity)
};

template <int VARIANT>
struct brush_opacity_frag_t : FragmentShaderImpl, brush_opacity_vert_t<VARIANT> {
  typedef brush_opacity_frag_t Self;
  typedef typename brush_opacity_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_opacity_vert_t<VARIANT>::InterpOutputs InterpOutputs;
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
  // draw_perspective: the varyings arrive divided by w and are interpolated linearly in screen space; each chunk
  // multiplies them back by w = 1 / gl_FragCoord.w (what glsl-to-cxx generates next to the plain pair)
  struct InterpPerspective {
    vec2 v_uv;
  };
  InterpPerspective interp_perspective;
  static void read_perspective_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    Float w = 1.0f / self->gl_FragCoord.w;
    self->interp_perspective.v_uv = init_interp(init->v_uv, step->v_uv);
    self->v_uv = self->interp_perspective.v_uv * w;
    self->interp_step.v_uv = step->v_uv * 4.0f;
  }
  ALWAYS_INLINE void step_perspective_inputs(int steps = 4) {
    this->step_perspective(steps);
    float chunks = steps * 0.25f;
    Float w = 1.0f / this->gl_FragCoord.w;
    interp_perspective.v_uv += interp_step.v_uv * chunks;
    v_uv = w * interp_perspective.v_uv;
  }

  // brush_opacity.glsl:56-74 + brush.glsl main
  void main() {
    Float perspective_divisor = mix(this->gl_FragCoord.w, Float(1.0f), Float(this->v_opacity_perspective_vec.y));
    vec2 uv = v_uv * perspective_divisor;
    uv = clamp(uv, vec2(this->v_uv_sample_bounds.sel(X, Y)), vec2(this->v_uv_sample_bounds.sel(Z, W)));
    vec4 color = texture(this->sColor0, uv);
    float alpha = this->v_opacity_perspective_vec.x;
    if (VARIANT == 1) alpha *= 1.0f;  // antialias_brush()
    vec4 frag = alpha * color;
    if (VARIANT == 1) frag *= Float(1.0f);  // do_clip()
    this->gl_FragColor = frag;
  }

  // brush_opacity.glsl:76-82
  void swgl_drawSpanRGBA8() {
    float perspective_divisor = mix(swgl_forceScalar(this->gl_FragCoord.w), 1.0f, this->v_opacity_perspective_vec.y);
    vec2 uv = v_uv * perspective_divisor;
    swgl_commitTextureLinearColorRGBA8(this->sColor0, uv, this->v_uv_sample_bounds, this->v_opacity_perspective_vec.x);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI_W()
  brush_opacity_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

typedef brush_opacity_frag_t<0> brush_opacity_frag;
typedef brush_opacity_frag_t<1> brush_opacity_ALPHA_PASS_frag;
typedef brush_opacity_frag_t<1> brush_opacity_ALPHA_PASS_ANTIALIASING_frag;
WR_PROGRAM(brush_opacity, "brush_opacity")
WR_PROGRAM(brush_opacity_ALPHA_PASS, "brush_opacity ALPHA_PASS")
WR_PROGRAM(brush_opacity_ALPHA_PASS_ANTIALIASING, "brush_opacity ALPHA_PASS,ANTIALIASING")
