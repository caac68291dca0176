// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "composite TEXTURE_2D"
// and "composite FAST_PATH,TEXTURE_2D": webrender/res/composite.glsl (RGB
// path, no YUV) restated in the glsl.h vocabulary.
#pragma once

template <bool FAST>
struct composite_vert_t : VertexShaderImpl, WrCommon {
  typedef composite_vert_t Self;
  vec2 aPosition;
  vec4_scalar aDeviceRect, aDeviceClipRect, aColor, aParams, aUvRect0;
  vec2_scalar aFlip;
  int a_loc[7];

  vec2 vUv;
  vec4_scalar vColor, vUVBounds;

  struct InterpOutputs {
    vec2_scalar vUv;
  };

  composite_vert_t() {
    static const char* names[7] = {"aPosition", "aDeviceRect", "aDeviceClipRect", "aColor", "aParams", "aUvRect0",
                                   "aFlip"};
    for (int i = 0; i < 7; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_Color0;
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
    load_flat_attrib(self->aFlip, attribs[L[self->a_loc[6]]], start, instance, count);
  }

  // composite.glsl:73-159
  void main() {
    vec4_scalar device_rect = mix(aDeviceRect, aDeviceRect.sel(Z, W, X, Y), aFlip.sel(X, Y, X, Y));
    vec2 world_pos = mix(device_rect.sel(X, Y), device_rect.sel(Z, W), aPosition);
    vec2 clipped_world_pos = clamp(world_pos, vec2(aDeviceClipRect.sel(X, Y)), vec2(aDeviceClipRect.sel(Z, W)));
    vec2 uv = (clipped_world_pos - vec2(device_rect.sel(X, Y))) / vec2(device_rect.sel(Z, W) - device_rect.sel(X, Y));
    uv = mix(vec2(aUvRect0.sel(X, Y)), vec2(aUvRect0.sel(Z, W)), uv);
    vec4_scalar uvBounds = make_vec4(min(aUvRect0.sel(X, Y), aUvRect0.sel(Z, W)), max(aUvRect0.sel(X, Y), aUvRect0.sel(Z, W)));
    if (int(aParams.y) == 1) {  // UV_TYPE_UNNORMALIZED
      vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
      uvBounds += vec4_scalar(0.5f, 0.5f, -0.5f, -0.5f);
      uv /= vec2(texture_size);
      uvBounds /= texture_size.sel(X, Y, X, Y);
    }
    vUv = uv;
    if (!FAST) {
      vUVBounds = uvBounds;
      vColor = aColor;
    }
    gl_Position = uTransform * vec4(clipped_world_pos, Float(0.0f), Float(1.0f));
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUv = get_nth(vUv, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(composite)
};

template <bool FAST>
struct composite_frag_t : FragmentShaderImpl, composite_vert_t<FAST> {
  typedef composite_frag_t Self;
  typedef typename composite_vert_t<FAST>::InterpOutputs InterpInputs;
  typedef typename composite_vert_t<FAST>::InterpOutputs InterpOutputs;
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

  // composite.glsl:163-192
  void main() {
    vec2 uv;
    if (FAST) uv = vUv;
    else uv = clamp(vUv, vec2(this->vUVBounds.sel(X, Y)), vec2(this->vUVBounds.sel(Z, W)));
    vec4 texel = texture(this->sColor0, uv);
    vec4 color;
    if (FAST) color = texel;
    else color = vec4(this->vColor) * texel;
    this->gl_FragColor = color;
  }

  // composite.glsl:195-234
  void swgl_drawSpanRGBA8() {
    vec4_scalar color, uvBounds;
    if (FAST) {
      color = vec4_scalar(1.0f);
      uvBounds = vec4_scalar(0.0f, 0.0f, 1.0f, 1.0f);
    } else {
      color = this->vColor;
      uvBounds = this->vUVBounds;
    }
    if (color != vec4_scalar(1.0f)) {
      swgl_commitTextureColorRGBA8(this->sColor0, vUv, uvBounds, color);
    } else {
      swgl_commitTextureRGBA8(this->sColor0, vUv, uvBounds);
    }
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  composite_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

typedef composite_frag_t<false> composite_TEXTURE_2D_frag;
typedef composite_frag_t<true> composite_FAST_PATH_TEXTURE_2D_frag;
WR_PROGRAM(composite_TEXTURE_2D, "composite TEXTURE_2D")
WR_PROGRAM(composite_FAST_PATH_TEXTURE_2D, "composite FAST_PATH,TEXTURE_2D")
