// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "cs_blur ALPHA_TARGET" and
// "cs_blur COLOR_TARGET" (webrender/res/cs_blur.glsl): one direction of a
// separable Gaussian blur.  Span path: swgl_commitGaussianBlurR8 / RGBA8.
#pragma once

template <bool COLOR>
struct cs_blur_vert_t : VertexShaderImpl, WrCommon {
  typedef cs_blur_vert_t Self;
  vec2 aPosition;
  int aBlurRenderTaskAddress, aBlurSourceTaskAddress, aBlurDirection;
  vec3_scalar aBlurParams;
  int a_loc[5];

  vec2 vUv;
  vec4_scalar vUvRect;
  vec2_scalar vOffsetScale;
  ivec2_scalar vSupport;
  vec2_scalar vGaussCoefficients;
  struct InterpOutputs {
    vec2_scalar vUv;
  };

  cs_blur_vert_t() {
    static const char* names[5] = {"aPosition", "aBlurRenderTaskAddress", "aBlurSourceTaskAddress", "aBlurDirection",
                                   "aBlurParams"};
    for (int i = 0; i < 5; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_Color0 | WR_S_RenderTasks;
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aBlurRenderTaskAddress, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aBlurSourceTaskAddress, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aBlurDirection, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aBlurParams, attribs[L[self->a_loc[4]]], start, instance, count);
  }

  // cs_blur.glsl:47-68
  void calculate_gauss_coefficients(float sigma) {
    vGaussCoefficients = vec2_scalar(1.0f / (sqrt(2.0f * 3.14159265f) * sigma), exp(-0.5f / (sigma * sigma)));
    vec3_scalar gauss_coefficient = vec3_scalar(vGaussCoefficients.x, vGaussCoefficients.y,
                                                vGaussCoefficients.y * vGaussCoefficients.y);
    float gauss_coefficient_total = gauss_coefficient.x;
    for (int i = 1; i <= vSupport.x; i += 2) {
      gauss_coefficient.x *= gauss_coefficient.y;
      gauss_coefficient.y *= gauss_coefficient.z;
      float gauss_coefficient_subtotal = gauss_coefficient.x;
      gauss_coefficient.x *= gauss_coefficient.y;
      gauss_coefficient.y *= gauss_coefficient.z;
      gauss_coefficient_subtotal += gauss_coefficient.x;
      gauss_coefficient_total += 2.0f * gauss_coefficient_subtotal;
    }
    vGaussCoefficients.x /= gauss_coefficient_total;
  }

  // cs_blur.glsl:70-116
  void main() {
    RenderTaskData blur_data = fetch_render_task_data(aBlurRenderTaskAddress);
    RenderTaskData src_data = fetch_render_task_data(aBlurSourceTaskAddress);
    RectWithEndpoint target_rect = blur_data.task_rect;
    RectWithEndpoint src_rect = src_data.task_rect;
    float blur_radius = aBlurParams.x;
    vec2_scalar blur_region = aBlurParams.sel(Y, Z);
    vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
    vSupport.x = int(ceil(1.5f * blur_radius)) * 2;
    if (vSupport.x > 0) {
      calculate_gauss_coefficients(blur_radius);
    } else {
      vGaussCoefficients = vec2_scalar(1.0f, 1.0f);
    }
    switch (aBlurDirection) {
      case 0: vOffsetScale = vec2_scalar(1.0f / texture_size.x, 0.0f); break;
      case 1: vOffsetScale = vec2_scalar(0.0f, 1.0f / texture_size.y); break;
      default: vOffsetScale = vec2_scalar(0.0f);
    }
    vUvRect = make_vec4(src_rect.p0 + vec2_scalar(0.5f), src_rect.p0 + blur_region - vec2_scalar(0.5f));
    vUvRect /= texture_size.sel(X, Y, X, Y);
    vec2 pos = mix(target_rect.p0, target_rect.p1, aPosition);
    vec2_scalar uv0 = src_rect.p0 / texture_size;
    vec2_scalar uv1 = src_rect.p1 / texture_size;
    vUv = mix(uv0, uv1, aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUv = get_nth(vUv, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_blur)
};

template <bool COLOR>
struct cs_blur_frag_t : FragmentShaderImpl, cs_blur_vert_t<COLOR> {
  typedef cs_blur_frag_t Self;
  typedef typename cs_blur_vert_t<COLOR>::InterpOutputs InterpInputs;
  typedef typename cs_blur_vert_t<COLOR>::InterpOutputs InterpOutputs;
  vec2 vUv;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vUv = init_interp(init->vUv, step->vUv);
    self->interp_step.vUv = step->vUv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vUv += interp_step.vUv * chunks;
  }

  // cs_blur.glsl:132-182
  void main() {
    vec4 original_color = texture(this->sColor0, vUv);
    if (!COLOR) original_color = vec4(original_color.x);
    vec3_scalar gauss_coefficient = vec3_scalar(this->vGaussCoefficients.x, this->vGaussCoefficients.y,
                                                this->vGaussCoefficients.y * this->vGaussCoefficients.y);
    vec4 avg_color = original_color * gauss_coefficient.x;
    int support = min(this->vSupport.x, 300);
    for (int i = 1; i <= support; i += 2) {
      gauss_coefficient.x *= gauss_coefficient.y;
      gauss_coefficient.y *= gauss_coefficient.z;
      float gauss_coefficient_subtotal = gauss_coefficient.x;
      gauss_coefficient.x *= gauss_coefficient.y;
      gauss_coefficient.y *= gauss_coefficient.z;
      gauss_coefficient_subtotal += gauss_coefficient.x;
      float gauss_ratio = gauss_coefficient.x / gauss_coefficient_subtotal;
      vec2_scalar offset = this->vOffsetScale * (float(i) + gauss_ratio);
      vec2 st0 = max(vUv - vec2(offset), vec2(this->vUvRect.sel(X, Y)));
      vec2 st1 = min(vUv + vec2(offset), vec2(this->vUvRect.sel(Z, W)));
      vec4 s0 = texture(this->sColor0, st0), s1 = texture(this->sColor0, st1);
      if (!COLOR) { s0 = vec4(s0.x); s1 = vec4(s1.x); }
      avg_color += (s0 + s1) * gauss_coefficient_subtotal;
    }
    this->gl_FragColor = avg_color;
  }

  void swgl_drawSpanRGBA8() {
    if (COLOR)
      swgl_commitGaussianBlurRGBA8(this->sColor0, vUv, this->vUvRect, this->vOffsetScale.x != 0.0f, this->vSupport.x,
                                   this->vGaussCoefficients);
  }
  void swgl_drawSpanR8() {
    if (!COLOR)
      swgl_commitGaussianBlurR8(this->sColor0, vUv, this->vUvRect, this->vOffsetScale.x != 0.0f, this->vSupport.x,
                                this->vGaussCoefficients);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  static int draw_span_R8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, R8);
  }
  WR_FRAGMENT_ABI()
  cs_blur_frag_t() {
    this->init_fragment_abi();
    if (COLOR) this->draw_span_RGBA8_func = &draw_span_RGBA8;
    else this->draw_span_R8_func = &draw_span_R8;
  }
};

typedef cs_blur_frag_t<false> cs_blur_ALPHA_TARGET_frag;
typedef cs_blur_frag_t<true> cs_blur_COLOR_TARGET_frag;
WR_PROGRAM(cs_blur_ALPHA_TARGET, "cs_blur ALPHA_TARGET")
WR_PROGRAM(cs_blur_COLOR_TARGET, "cs_blur COLOR_TARGET")
