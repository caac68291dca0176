// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_yuv_image TEXTURE_2D,YUV",
// "brush_yuv_image ALPHA_PASS,TEXTURE_2D,YUV" and "... ALPHA_PASS,ANTIALIASING,TEXTURE_2D,YUV"
// (webrender/res/brush_yuv_image.glsl + yuv.glsl; antialias_brush() == 1 and do_clip() == 1 under SWGL).
#pragma once

template <int VARIANT>  // 0 opaque, 1 ALPHA_PASS
struct brush_yuv_image_vert_t : BrushVertBase<brush_yuv_image_vert_t<VARIANT>> {
  typedef brush_yuv_image_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 1;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2 vUv_Y, vUv_U, vUv_V;
  vec4_scalar vUvBounds_Y, vUvBounds_U, vUvBounds_V;
  vec3_scalar vYcbcrBias;
  mat3_scalar vRgbFromDebiasedYcbcr;
  ivec2_scalar vFormat;
  int32_t vRescaleFactor;
  struct InterpOutputs {
    vec2_scalar vUv_Y, vUv_U, vUv_V;
  };

  brush_yuv_image_vert_t() {
    this->sampler_mask |= WR_S_Color0 | WR_S_Color1 | WR_S_Color2;
    this->init_vertex_abi();
  }

  // yuv.glsl:163-178
  static void write_uv_rect(vec2_scalar uv0, vec2_scalar uv1, vec2 f, vec2_scalar texture_size, vec2& uv,
                            vec4_scalar& uv_bounds) {
    uv = mix(vec2(uv0), vec2(uv1), f);
    uv_bounds = make_vec4(uv0 + vec2_scalar(0.5f), uv1 - vec2_scalar(0.5f));
    uv /= vec2(texture_size);
    uv_bounds /= texture_size.sel(X, Y, X, Y);
  }

  // brush_yuv_image.glsl:41-93
  void brush_vs(VertexInfo& vi, int prim_address, RectWithEndpoint local_rect, RectWithEndpoint,
                ivec4_scalar prim_user_data, int, mat4_scalar, PictureTask&, int, vec4_scalar) {
    vec2 f = (vi.local_pos - vec2(local_rect.p0)) / vec2(local_rect.p1 - local_rect.p0);
    // fetch_yuv_primitive (brush_yuv_image.glsl:31-39)
    vec4_scalar data = this->fetch_gpu_cache(prim_address, 0);
    int channel_bit_depth = int(data.x);
    int color_space = int(data.y);
    int yuv_format = int(data.z);
    vRescaleFactor = 0;
    if (channel_bit_depth > 8 && yuv_format != 1) vRescaleFactor = 16 - channel_bit_depth;
    composite_yuv_vert::color_matrix(color_space, yuv_format, channel_bit_depth, vYcbcrBias, vRgbFromDebiasedYcbcr);
    vFormat.x = yuv_format;
    if (vFormat.x == 3 || vFormat.x == 99) {
      vec4_scalar ry = this->fetch_gpu_cache(prim_user_data.x, 0);
      vec4_scalar ru = this->fetch_gpu_cache(prim_user_data.y, 0);
      vec4_scalar rv = this->fetch_gpu_cache(prim_user_data.z, 0);
      write_uv_rect(ry.sel(X, Y), ry.sel(Z, W), f, make_vec2(textureSize(this->sColor0, 0)), vUv_Y, vUvBounds_Y);
      write_uv_rect(ru.sel(X, Y), ru.sel(Z, W), f, make_vec2(textureSize(this->sColor1, 0)), vUv_U, vUvBounds_U);
      write_uv_rect(rv.sel(X, Y), rv.sel(Z, W), f, make_vec2(textureSize(this->sColor2, 0)), vUv_V, vUvBounds_V);
    } else if (vFormat.x == 0 || vFormat.x == 1) {
      vec4_scalar ry = this->fetch_gpu_cache(prim_user_data.x, 0);
      vec4_scalar ru = this->fetch_gpu_cache(prim_user_data.y, 0);
      write_uv_rect(ry.sel(X, Y), ry.sel(Z, W), f, make_vec2(textureSize(this->sColor0, 0)), vUv_Y, vUvBounds_Y);
      write_uv_rect(ru.sel(X, Y), ru.sel(Z, W), f, make_vec2(textureSize(this->sColor1, 0)), vUv_U, vUvBounds_U);
    } else if (vFormat.x == 4) {
      vec4_scalar ry = this->fetch_gpu_cache(prim_user_data.x, 0);
      write_uv_rect(ry.sel(X, Y), ry.sel(Z, W), f, make_vec2(textureSize(this->sColor0, 0)), vUv_Y, vUvBounds_Y);
    }
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUv_Y = get_nth(vUv_Y, n);
      dest->vUv_U = get_nth(vUv_U, n);
      dest->vUv_V = get_nth(vUv_V, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_yuv_image)
};

template <int VARIANT>
struct brush_yuv_image_frag_t : FragmentShaderImpl, brush_yuv_image_vert_t<VARIANT> {
  typedef brush_yuv_image_frag_t Self;
  typedef typename brush_yuv_image_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_yuv_image_vert_t<VARIANT>::InterpOutputs InterpOutputs;
  vec2 vUv_Y, vUv_U, vUv_V;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vUv_Y = init_interp(init->vUv_Y, step->vUv_Y);
    self->interp_step.vUv_Y = step->vUv_Y * 4.0f;
    self->vUv_U = init_interp(init->vUv_U, step->vUv_U);
    self->interp_step.vUv_U = step->vUv_U * 4.0f;
    self->vUv_V = init_interp(init->vUv_V, step->vUv_V);
    self->interp_step.vUv_V = step->vUv_V * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vUv_Y += interp_step.vUv_Y * chunks;
    vUv_U += interp_step.vUv_U * chunks;
    vUv_V += interp_step.vUv_V * chunks;
  }

  // brush_fs (brush_yuv_image.glsl:97-120) → sample_yuv (yuv.glsl:183-246) + brush.glsl main
  void main() {
    vec3 s3;
    switch (this->vFormat.x) {
      case 3: {
        vec2 uy = clamp(vUv_Y, vec2(this->vUvBounds_Y.sel(X, Y)), vec2(this->vUvBounds_Y.sel(Z, W)));
        vec2 uu = clamp(vUv_U, vec2(this->vUvBounds_U.sel(X, Y)), vec2(this->vUvBounds_U.sel(Z, W)));
        vec2 uv = clamp(vUv_V, vec2(this->vUvBounds_V.sel(X, Y)), vec2(this->vUvBounds_V.sel(Z, W)));
        s3.x = texture(this->sColor0, uy).x;
        s3.y = texture(this->sColor1, uu).x;
        s3.z = texture(this->sColor2, uv).x;
        break;
      }
      case 0: case 1: case 2: {
        vec2 uy = clamp(vUv_Y, vec2(this->vUvBounds_Y.sel(X, Y)), vec2(this->vUvBounds_Y.sel(Z, W)));
        vec2 uu = clamp(vUv_U, vec2(this->vUvBounds_U.sel(X, Y)), vec2(this->vUvBounds_U.sel(Z, W)));
        s3.x = texture(this->sColor0, uy).x;
        vec4 t = texture(this->sColor1, uu);
        s3.y = t.x;
        s3.z = t.y;
        break;
      }
      case 4: {
        vec2 uy = clamp(vUv_Y, vec2(this->vUvBounds_Y.sel(X, Y)), vec2(this->vUvBounds_Y.sel(Z, W)));
        vec4 t = texture(this->sColor0, uy);
        s3 = vec3(t.y, t.z, t.x);
        break;
      }
      default:
        s3 = vec3(Float(0.0f), Float(0.0f), Float(0.0f));
        break;
    }
    vec3 rgb = this->vRgbFromDebiasedYcbcr * (s3 - vec3(this->vYcbcrBias));
    if (VARIANT == 1) rgb = clamp(rgb, vec3(Float(0.0f)), vec3(Float(1.0f)));  // ALPHA_PASS && SWGL_CLIP_MASK
    vec4 color = vec4(rgb.x, rgb.y, rgb.z, Float(1.0f));
    if (VARIANT == 1) {
      color *= Float(1.0f);  // antialias_brush()
      color *= Float(1.0f);  // do_clip()
    }
    this->gl_FragColor = color;
  }

  // brush_yuv_image.glsl:122-143
  void swgl_drawSpanRGBA8() {
    if (this->vFormat.x == 3) {
      swgl_commitTextureLinearYUV(this->sColor0, vUv_Y, this->vUvBounds_Y, this->sColor1, vUv_U, this->vUvBounds_U,
                                  this->sColor2, vUv_V, this->vUvBounds_V, this->vYcbcrBias,
                                  this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    } else if (this->vFormat.x == 0 || this->vFormat.x == 1) {
      swgl_commitTextureLinearYUV(this->sColor0, vUv_Y, this->vUvBounds_Y, this->sColor1, vUv_U, this->vUvBounds_U,
                                  this->vYcbcrBias, this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    } else if (this->vFormat.x == 4) {
      swgl_commitTextureLinearYUV(this->sColor0, vUv_Y, this->vUvBounds_Y, this->vYcbcrBias,
                                  this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    }
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  brush_yuv_image_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

typedef brush_yuv_image_frag_t<0> brush_yuv_image_TEXTURE_2D_YUV_frag;
typedef brush_yuv_image_frag_t<1> brush_yuv_image_ALPHA_PASS_TEXTURE_2D_YUV_frag;
typedef brush_yuv_image_frag_t<1> brush_yuv_image_ALPHA_PASS_ANTIALIASING_TEXTURE_2D_YUV_frag;
WR_PROGRAM(brush_yuv_image_TEXTURE_2D_YUV, "brush_yuv_image TEXTURE_2D,YUV")
WR_PROGRAM(brush_yuv_image_ALPHA_PASS_TEXTURE_2D_YUV, "brush_yuv_image ALPHA_PASS,TEXTURE_2D,YUV")
WR_PROGRAM(brush_yuv_image_ALPHA_PASS_ANTIALIASING_TEXTURE_2D_YUV, "brush_yuv_image ALPHA_PASS,ANTIALIASING,TEXTURE_2D,YUV")
