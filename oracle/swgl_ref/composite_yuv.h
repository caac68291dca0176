// TEST INFRASTRUCTURE — hand-instantiated SWGL program "composite TEXTURE_2D,YUV":
// webrender/res/composite.glsl with WR_FEATURE_YUV + webrender/res/yuv.glsl restated
// in the glsl.h vocabulary (the shape glsl-to-cxx emits, glsl-to-cxx/src/lib.rs:200-245).
#pragma once

struct composite_yuv_vert : VertexShaderImpl, WrCommon {
  typedef composite_yuv_vert Self;
  vec2 aPosition;
  vec4_scalar aDeviceRect, aDeviceClipRect, aColor, aParams, aUvRect0, aUvRect1, aUvRect2;
  vec2_scalar aFlip;
  int a_loc[9];

  // flat varyings (composite.glsl:18-33)
  vec3_scalar vYcbcrBias;
  mat3_scalar vRgbFromDebiasedYcbcr;
  ivec2_scalar vYuvFormat;
  int32_t vRescaleFactor;
  vec4_scalar vUVBounds_y, vUVBounds_u, vUVBounds_v;
  vec2 vUV_y, vUV_u, vUV_v;

  struct InterpOutputs {
    vec2_scalar vUV_y, vUV_u, vUV_v;
  };

  composite_yuv_vert() {
    static const char* names[9] = {"aPosition", "aDeviceRect", "aDeviceClipRect", "aColor", "aParams",
                                   "aUvRect0",  "aUvRect1",    "aUvRect2",        "aFlip"};
    for (int i = 0; i < 9; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_Color0 | WR_S_Color1 | WR_S_Color2;
    init_vertex_abi();
  }

  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aDeviceRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aDeviceClipRect, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aColor, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aParams, attribs[L[self->a_loc[4]]], start, instance, count);
    load_flat_attrib(self->aUvRect0, attribs[L[self->a_loc[5]]], start, instance, count);
    load_flat_attrib(self->aUvRect1, attribs[L[self->a_loc[6]]], start, instance, count);
    load_flat_attrib(self->aUvRect2, attribs[L[self->a_loc[7]]], start, instance, count);
    load_flat_attrib(self->aFlip, attribs[L[self->a_loc[8]]], start, instance, count);
  }

  // yuv.glsl:79-96
  static vec4_scalar zero_one_identity(int bit_depth, float channel_max) {
    float all_ones_normalized = float((1 << bit_depth) - 1) / channel_max;
    return vec4_scalar(0.0f, 0.0f, all_ones_normalized, all_ones_normalized);
  }
  static vec4_scalar zero_one_narrow_range(int bit_depth, float channel_max) {
    ivec4_scalar zero_one_ints = ivec4_scalar(16, 128, 235, 240) << (bit_depth - 8);
    return vec4_scalar(float(zero_one_ints.x), float(zero_one_ints.y), float(zero_one_ints.z),
                       float(zero_one_ints.w)) /
           channel_max;
  }
  static vec4_scalar zero_one_full_range(int bit_depth, float channel_max) {
    vec4_scalar narrow = zero_one_narrow_range(bit_depth, channel_max);
    vec4_scalar identity = zero_one_identity(bit_depth, channel_max);
    return vec4_scalar(0.0f, narrow.y, identity.z, identity.w);
  }

  // get_yuv_color_info + get_rgb_from_ycbcr_info (yuv.glsl:98-161)
  static void color_matrix(int color_space, int yuv_format, int channel_bit_depth, vec3_scalar& vYcbcrBias,
                           mat3_scalar& vRgbFromDebiasedYcbcr) {
    float channel_max = 255.0f;
    if (channel_bit_depth > 8) {
      if (yuv_format == 1) channel_max = float((1 << channel_bit_depth) - 1);
      else channel_max = 65535.0f;
    }
    const mat3_scalar rec601(vec3_scalar(1.00000f, 1.00000f, 1.00000f), vec3_scalar(0.00000f, -0.17207f, 0.88600f),
                             vec3_scalar(0.70100f, -0.35707f, 0.00000f));
    const mat3_scalar rec709(vec3_scalar(1.00000f, 1.00000f, 1.00000f), vec3_scalar(0.00000f, -0.09366f, 0.92780f),
                             vec3_scalar(0.78740f, -0.23406f, 0.00000f));
    const mat3_scalar rec2020(vec3_scalar(1.00000f, 1.00000f, 1.00000f), vec3_scalar(0.00000f, -0.08228f, 0.94070f),
                              vec3_scalar(0.73730f, -0.28568f, 0.00000f));
    const mat3_scalar gbr(vec3_scalar(0.0f, 1.0f, 0.0f), vec3_scalar(0.0f, 0.0f, 1.0f), vec3_scalar(1.0f, 0.0f, 0.0f));
    mat3_scalar rgb_from_yuv;
    vec4_scalar zero_one;
    switch (color_space) {
      case 0: rgb_from_yuv = rec601; zero_one = zero_one_narrow_range(channel_bit_depth, channel_max); break;
      case 1: rgb_from_yuv = rec601; zero_one = zero_one_full_range(channel_bit_depth, channel_max); break;
      case 2: rgb_from_yuv = rec709; zero_one = zero_one_narrow_range(channel_bit_depth, channel_max); break;
      case 3: rgb_from_yuv = rec709; zero_one = zero_one_full_range(channel_bit_depth, channel_max); break;
      case 4: rgb_from_yuv = rec2020; zero_one = zero_one_narrow_range(channel_bit_depth, channel_max); break;
      case 5: rgb_from_yuv = rec2020; zero_one = zero_one_full_range(channel_bit_depth, channel_max); break;
      default: rgb_from_yuv = gbr; zero_one = zero_one_identity(channel_bit_depth, channel_max); break;
    }
    // get_rgb_from_ycbcr_info (yuv.glsl:144-161)
    vec2_scalar zero = zero_one.sel(X, Y);
    vec2_scalar one = zero_one.sel(Z, W);
    vec2_scalar scale = 1.0f / (one - zero);
    vYcbcrBias = vec3_scalar(zero.x, zero.y, zero.y);
    mat3_scalar yuv_from_debiased_ycbcr(vec3_scalar(scale.x, 0.0f, 0.0f), vec3_scalar(0.0f, scale.y, 0.0f),
                                        vec3_scalar(0.0f, 0.0f, scale.y));
    vRgbFromDebiasedYcbcr = rgb_from_yuv * yuv_from_debiased_ycbcr;
  }

  // yuv.glsl:163-178
  void write_uv_rect(vec2_scalar uv0, vec2_scalar uv1, vec2 f, vec2_scalar texture_size, vec2& uv,
                     vec4_scalar& uv_bounds) {
    uv = mix(vec2(uv0), vec2(uv1), f);
    uv_bounds = make_vec4(uv0 + vec2_scalar(0.5f), uv1 - vec2_scalar(0.5f));
    uv /= vec2(texture_size);
    uv_bounds /= texture_size.sel(X, Y, X, Y);
  }

  // composite.glsl:73-130 (YUV branch)
  void main() {
    vec4_scalar device_rect = mix(aDeviceRect, aDeviceRect.sel(Z, W, X, Y), aFlip.sel(X, Y, X, Y));
    vec2 world_pos = mix(device_rect.sel(X, Y), device_rect.sel(Z, W), aPosition);
    vec2 clipped_world_pos = clamp(world_pos, vec2(aDeviceClipRect.sel(X, Y)), vec2(aDeviceClipRect.sel(Z, W)));
    vec2 uv = (clipped_world_pos - vec2(device_rect.sel(X, Y))) / vec2(device_rect.sel(Z, W) - device_rect.sel(X, Y));

    // fetch_yuv_primitive (composite.glsl:63-70)
    int color_space = int(aParams.y);
    int yuv_format = int(aParams.z);
    int channel_bit_depth = int(aParams.w);

    vRescaleFactor = 0;
    if (channel_bit_depth > 8 && yuv_format != 1 /* YUV_FORMAT_P010 */) {
      vRescaleFactor = 16 - channel_bit_depth;
    }

    color_matrix(color_space, yuv_format, channel_bit_depth, vYcbcrBias, vRgbFromDebiasedYcbcr);
    vYuvFormat.x = yuv_format;

    write_uv_rect(aUvRect0.sel(X, Y), aUvRect0.sel(Z, W), uv, make_vec2(textureSize(sColor0, 0)), vUV_y, vUVBounds_y);
    write_uv_rect(aUvRect1.sel(X, Y), aUvRect1.sel(Z, W), uv, make_vec2(textureSize(sColor1, 0)), vUV_u, vUVBounds_u);
    write_uv_rect(aUvRect2.sel(X, Y), aUvRect2.sel(Z, W), uv, make_vec2(textureSize(sColor2, 0)), vUV_v, vUVBounds_v);

    gl_Position = uTransform * vec4(clipped_world_pos, Float(0.0f), Float(1.0f));
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUV_y = get_nth(vUV_y, n);
      dest->vUV_u = get_nth(vUV_u, n);
      dest->vUV_v = get_nth(vUV_v, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(composite_yuv)
};

struct composite_TEXTURE_2D_YUV_frag : FragmentShaderImpl, composite_yuv_vert {
  typedef composite_TEXTURE_2D_YUV_frag Self;
  typedef composite_yuv_vert::InterpOutputs InterpInputs;
  typedef composite_yuv_vert::InterpOutputs InterpOutputs;
  vec2 vUV_y, vUV_u, vUV_v;
  InterpInputs interp_step;

  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vUV_y = init_interp(init->vUV_y, step->vUV_y);
    self->interp_step.vUV_y = step->vUV_y * 4.0f;
    self->vUV_u = init_interp(init->vUV_u, step->vUV_u);
    self->interp_step.vUV_u = step->vUV_u * 4.0f;
    self->vUV_v = init_interp(init->vUV_v, step->vUV_v);
    self->interp_step.vUV_v = step->vUV_v * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vUV_y += interp_step.vUV_y * chunks;
    vUV_u += interp_step.vUV_u * chunks;
    vUV_v += interp_step.vUV_v * chunks;
  }

  // sample_yuv (yuv.glsl:183-246) through composite.glsl:163-176
  void main() {
    vec3 ycbcr_sample;
    switch (this->vYuvFormat.x) {
      case 3: {  // YUV_FORMAT_PLANAR
        vec2 uv_y = clamp(vUV_y, vec2(this->vUVBounds_y.sel(X, Y)), vec2(this->vUVBounds_y.sel(Z, W)));
        vec2 uv_u = clamp(vUV_u, vec2(this->vUVBounds_u.sel(X, Y)), vec2(this->vUVBounds_u.sel(Z, W)));
        vec2 uv_v = clamp(vUV_v, vec2(this->vUVBounds_v.sel(X, Y)), vec2(this->vUVBounds_v.sel(Z, W)));
        ycbcr_sample.x = texture(this->sColor0, uv_y).x;
        ycbcr_sample.y = texture(this->sColor1, uv_u).x;
        ycbcr_sample.z = texture(this->sColor2, uv_v).x;
        break;
      }
      case 0:    // NV12
      case 1:    // P010
      case 2: {  // NV16
        vec2 uv_y = clamp(vUV_y, vec2(this->vUVBounds_y.sel(X, Y)), vec2(this->vUVBounds_y.sel(Z, W)));
        vec2 uv_uv = clamp(vUV_u, vec2(this->vUVBounds_u.sel(X, Y)), vec2(this->vUVBounds_u.sel(Z, W)));
        ycbcr_sample.x = texture(this->sColor0, uv_y).x;
        vec4 t = texture(this->sColor1, uv_uv);
        ycbcr_sample.y = t.x;
        ycbcr_sample.z = t.y;
        break;
      }
      case 4: {  // YUV_FORMAT_INTERLEAVED
        vec2 uv_y = clamp(vUV_y, vec2(this->vUVBounds_y.sel(X, Y)), vec2(this->vUVBounds_y.sel(Z, W)));
        vec4 t = texture(this->sColor0, uv_y);
        ycbcr_sample = vec3(t.y, t.z, t.x);
        break;
      }
      default:
        ycbcr_sample = vec3(Float(0.0f), Float(0.0f), Float(0.0f));
        break;
    }
    vec3 rgb = this->vRgbFromDebiasedYcbcr * (ycbcr_sample - vec3(this->vYcbcrBias));
    this->gl_FragColor = vec4(rgb.x, rgb.y, rgb.z, Float(1.0f));
  }

  // composite.glsl:195-214
  void swgl_drawSpanRGBA8() {
    if (this->vYuvFormat.x == 3) {
      swgl_commitTextureLinearYUV(this->sColor0, vUV_y, this->vUVBounds_y, this->sColor1, vUV_u, this->vUVBounds_u,
                                  this->sColor2, vUV_v, this->vUVBounds_v, this->vYcbcrBias,
                                  this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    } else if (this->vYuvFormat.x == 0 || this->vYuvFormat.x == 1) {
      swgl_commitTextureLinearYUV(this->sColor0, vUV_y, this->vUVBounds_y, this->sColor1, vUV_u, this->vUVBounds_u,
                                  this->vYcbcrBias, this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    } else if (this->vYuvFormat.x == 4) {
      swgl_commitTextureLinearYUV(this->sColor0, vUV_y, this->vUVBounds_y, this->vYcbcrBias,
                                  this->vRgbFromDebiasedYcbcr, this->vRescaleFactor);
    }
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  composite_TEXTURE_2D_YUV_frag() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
typedef composite_yuv_vert composite_TEXTURE_2D_YUV_vert;
WR_PROGRAM(composite_TEXTURE_2D_YUV, "composite TEXTURE_2D,YUV")
