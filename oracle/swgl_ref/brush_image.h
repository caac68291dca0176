// TEST INFRASTRUCTURE — hand-instantiated SWGL programs
//   "brush_image TEXTURE_2D", "brush_image ALPHA_PASS,TEXTURE_2D",
//   "brush_image ADVANCED_BLEND,ALPHA_PASS,TEXTURE_2D"
// (webrender/res/brush_image.glsl without WR_FEATURE_REPETITION; ANTIALIASING
// is a no-op under SWGL_ANTIALIAS; ADVANCED_BLEND only changes GLSL layout
// qualifiers).  VARIANT: 0 = opaque, 1 = ALPHA_PASS.
#pragma once

template <int VARIANT>
struct brush_image_vert_t : BrushVertBase<brush_image_vert_t<VARIANT>> {
  typedef brush_image_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 3;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2 v_uv;
  vec4_scalar v_color;
  vec2_scalar v_mask_swizzle, v_tile_repeat_bounds;
  vec4_scalar v_uv_bounds, v_uv_sample_bounds;
  vec2_scalar v_perspective;
  struct InterpOutputs {
    vec2_scalar v_uv;
  };

  brush_image_vert_t() {
    this->sampler_mask |= WR_S_Color0;
    this->init_vertex_abi();
  }

  // brush_image.glsl:57-283 (non-REPETITION)
  void brush_vs(VertexInfo& vi, int prim_address, RectWithEndpoint prim_rect, RectWithEndpoint segment_rect,
                ivec4_scalar prim_user_data, int specific_resource_address, mat4_scalar, PictureTask&,
                int brush_flags, vec4_scalar segment_data) {
    vec4_scalar d0 = this->fetch_gpu_cache(prim_address, 0);
    vec4_scalar d2 = this->fetch_gpu_cache(prim_address, 2);
    vec4_scalar image_color = d0;
    vec2_scalar stretch_size = d2.sel(X, Y);
    vec2_scalar texture_size = make_vec2(textureSize(this->sColor0, 0));
    vec4_scalar r0 = this->fetch_gpu_cache(specific_resource_address, 0);
    RectWithEndpoint res_uv_rect = RectWithEndpoint{r0.sel(X, Y), r0.sel(Z, W)};
    vec2_scalar uv0 = res_uv_rect.p0;
    vec2_scalar uv1 = res_uv_rect.p1;
    RectWithEndpoint local_rect = prim_rect;
    if (stretch_size.x < 0.0f) {
      stretch_size = local_rect.p1 - local_rect.p0;
    }
    if ((brush_flags & WR_BRUSH_FLAG_SEGMENT_RELATIVE) != 0) {
      local_rect = segment_rect;
      stretch_size = local_rect.p1 - local_rect.p0;
      if ((brush_flags & WR_BRUSH_FLAG_TEXEL_RECT) != 0) {
        vec2_scalar uv_size = res_uv_rect.p1 - res_uv_rect.p0;
        uv0 = res_uv_rect.p0 + segment_data.sel(X, Y) * uv_size;
        uv1 = res_uv_rect.p0 + segment_data.sel(Z, W) * uv_size;
      }
    }
    float perspective_interpolate = (brush_flags & WR_BRUSH_FLAG_PERSPECTIVE_INTERPOLATION) != 0 ? 1.0f : 0.0f;
    v_perspective.x = perspective_interpolate;
    if ((brush_flags & WR_BRUSH_FLAG_NORMALIZED_UVS) != 0) {
      uv0 *= texture_size;
      uv1 *= texture_size;
    }
    vec2_scalar min_uv = min(uv0, uv1);
    vec2_scalar max_uv = max(uv0, uv1);
    v_uv_sample_bounds = make_vec4(min_uv + make_vec2(0.5f), max_uv - make_vec2(0.5f)) / texture_size.sel(X, Y, X, Y);
    vec2 f = (vi.local_pos - vec2(local_rect.p0)) / vec2(local_rect.p1 - local_rect.p0);
    int color_mode = prim_user_data.x & 0xffff;
    int blend_mode = prim_user_data.x >> 16;
    int raster_space = prim_user_data.y;
    if (raster_space == 1) {
      // get_image_quad_uv (prim_shared.glsl:202-208)
      vec4_scalar st_tl = this->fetch_gpu_cache(specific_resource_address + 2, 0);
      vec4_scalar st_tr = this->fetch_gpu_cache(specific_resource_address + 2, 1);
      vec4_scalar st_bl = this->fetch_gpu_cache(specific_resource_address + 2, 2);
      vec4_scalar st_br = this->fetch_gpu_cache(specific_resource_address + 2, 3);
      vec4 x = mix(st_tl, st_tr, f.x);
      vec4 y = mix(st_bl, st_br, f.x);
      vec4 z = mix(x, y, f.y);
      f = z.sel(X, Y) / z.w;
    }
    vec2_scalar repeat = (local_rect.p1 - local_rect.p0) / stretch_size;
    v_uv = mix(uv0, uv1, f) - min_uv;
    v_uv *= vec2(repeat);
    vec2_scalar normalized_offset = vec2_scalar(0.0f);
    v_uv /= vec2(texture_size);
    if (perspective_interpolate == 0.0f) {
      v_uv *= vi.world_pos.w;
    }
    v_uv_bounds = make_vec4(min_uv, max_uv) / texture_size.sel(X, Y, X, Y);
    if (VARIANT == 1) {
      v_tile_repeat_bounds = repeat + normalized_offset;
      float opacity = float(prim_user_data.z) / 65535.0f;
      switch (blend_mode) {
        case 0:
          image_color.w *= opacity;
          break;
        default:
          image_color *= opacity;
          break;
      }
      switch (color_mode) {
        case 0:  // COLOR_MODE_ALPHA
        case 2:  // COLOR_MODE_BITMAP_SHADOW
          swgl_blendDropShadow(image_color);
          v_mask_swizzle = vec2_scalar(1.0f, 0.0f);
          v_color = vec4_scalar(1.0f);
          break;
        case 4:  // COLOR_MODE_IMAGE
          v_mask_swizzle = vec2_scalar(1.0f, 0.0f);
          v_color = image_color;
          break;
        case 3:  // COLOR_MODE_COLOR_BITMAP
          v_mask_swizzle = vec2_scalar(1.0f, 0.0f);
          v_color = vec4_scalar(image_color.w);
          break;
        case 1:  // COLOR_MODE_SUBPX_DUAL_SOURCE
          v_mask_swizzle = vec2_scalar(image_color.w, 0.0f);
          v_color = image_color;
          break;
        case 5:  // COLOR_MODE_MULTIPLY_DUAL_SOURCE
          v_mask_swizzle = vec2_scalar(-image_color.w, image_color.w);
          v_color = image_color;
          break;
        default:
          v_mask_swizzle = vec2_scalar(0.0f);
          v_color = vec4_scalar(1.0f);
      }
    }
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_uv = get_nth(v_uv, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_image)
};

template <int VARIANT>
struct brush_image_frag_t : FragmentShaderImpl, brush_image_vert_t<VARIANT> {
  typedef brush_image_frag_t Self;
  typedef typename brush_image_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_image_vert_t<VARIANT>::InterpOutputs InterpOutputs;
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

  // brush_image.glsl:319-352 + brush.glsl main
  void main() {
    Float perspective_divisor = mix(this->gl_FragCoord.w, Float(1.0f), Float(this->v_perspective.x));
    vec2 repeated_uv = v_uv * perspective_divisor + vec2(this->v_uv_bounds.sel(X, Y));
    vec2 uv = clamp(repeated_uv, vec2(this->v_uv_sample_bounds.sel(X, Y)), vec2(this->v_uv_sample_bounds.sel(Z, W)));
    vec4 texel = texture(this->sColor0, uv);
    vec4 color;
    if (VARIANT == 1) {
      Float alpha = 1.0f;
      vec3 rgb = texel.sel(X, Y, Z) * Float(this->v_mask_swizzle.x) + texel.sel(W, W, W) * Float(this->v_mask_swizzle.y);
      texel.x = rgb.x; texel.y = rgb.y; texel.z = rgb.z;
      vec4 alpha_mask = texel * alpha;
      color = vec4(this->v_color) * alpha_mask;
      color *= Float(1.0f);  // do_clip()
    } else {
      color = texel;
    }
    this->gl_FragColor = color;
  }

  // brush_image.glsl:386-429
  void swgl_drawSpanRGBA8() {
    if (!swgl_isTextureRGBA8(this->sColor0)) return;
    if (VARIANT == 1) {
      if (this->v_mask_swizzle != vec2_scalar(1.0f, 0.0f)) return;
    }
    float perspective_divisor = mix(swgl_forceScalar(this->gl_FragCoord.w), 1.0f, this->v_perspective.x);
    vec2 uv = v_uv * perspective_divisor + vec2(this->v_uv_bounds.sel(X, Y));
    if (VARIANT == 1) {
      if (this->v_color != vec4_scalar(1.0f)) {
        swgl_commitTextureColorRGBA8(this->sColor0, uv, this->v_uv_sample_bounds, this->v_color);
        return;
      }
    }
    swgl_commitTextureRGBA8(this->sColor0, uv, this->v_uv_sample_bounds);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI_W()
  brush_image_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

typedef brush_image_frag_t<0> brush_image_TEXTURE_2D_frag;
typedef brush_image_frag_t<1> brush_image_ALPHA_PASS_TEXTURE_2D_frag;
typedef brush_image_frag_t<1> brush_image_ADVANCED_BLEND_ALPHA_PASS_TEXTURE_2D_frag;
WR_PROGRAM(brush_image_TEXTURE_2D, "brush_image TEXTURE_2D")
WR_PROGRAM(brush_image_ALPHA_PASS_TEXTURE_2D, "brush_image ALPHA_PASS,TEXTURE_2D")
WR_PROGRAM(brush_image_ADVANCED_BLEND_ALPHA_PASS_TEXTURE_2D, "brush_image ADVANCED_BLEND,ALPHA_PASS,TEXTURE_2D")
