// TEST INFRASTRUCTURE — hand-instantiated SWGL program "cs_scale TEXTURE_2D"
// (webrender/res/cs_scale.glsl): scaled copy of a source rect (down-scaling steps
// of the blur pipeline, external surfaces).
#pragma once

struct cs_scale_vert : VertexShaderImpl, WrCommon {
  typedef cs_scale_vert Self;
  vec2 aPosition;
  vec4_scalar aScaleTargetRect, aScaleSourceRect;
  float aSourceRectType;
  int a_loc[4];
  vec2 vUv;
  vec4_scalar vUvRect;
  struct InterpOutputs {
    vec2_scalar vUv;
  };
  cs_scale_vert() {
    static const char* names[4] = {"aPosition", "aScaleTargetRect", "aScaleSourceRect", "aSourceRectType"};
    for (int i = 0; i < 4; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_Color0;
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aScaleTargetRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aScaleSourceRect, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aSourceRectType, attribs[L[self->a_loc[3]]], start, instance, count);
  }
  // cs_scale.glsl:24-53
  void main() {
    vec2_scalar src_offset = aScaleSourceRect.sel(X, Y);
    vec2_scalar src_size = aScaleSourceRect.sel(Z, W) - aScaleSourceRect.sel(X, Y);
    vUvRect = make_vec4(min(aScaleSourceRect.sel(X, Y), aScaleSourceRect.sel(Z, W)),
                        max(aScaleSourceRect.sel(X, Y), aScaleSourceRect.sel(Z, W)));
    vUv = vec2(src_offset) + vec2(src_size) * aPosition;
    if (int(aSourceRectType) == 1) {  // UV_TYPE_UNNORMALIZED
      vUvRect = make_vec4(vUvRect.sel(X, Y) + vec2_scalar(0.5f), vUvRect.sel(Z, W) - vec2_scalar(0.5f));
      vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
      vUvRect /= texture_size.sel(X, Y, X, Y);
      vUv /= vec2(texture_size);
    }
    vec2 pos = mix(aScaleTargetRect.sel(X, Y), aScaleTargetRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vUv = get_nth(vUv, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_scale)
};

struct cs_scale_frag : FragmentShaderImpl, cs_scale_vert {
  typedef cs_scale_frag Self;
  typedef cs_scale_vert::InterpOutputs InterpInputs;
  typedef cs_scale_vert::InterpOutputs InterpOutputs;
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
  void main() {
    vec2 st = clamp(vUv, vec2(vUvRect.sel(X, Y)), vec2(vUvRect.sel(Z, W)));
    gl_FragColor = texture(sColor0, st);
  }
  void swgl_drawSpanRGBA8() { swgl_commitTextureLinearRGBA8(sColor0, vUv, vUvRect); }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  cs_scale_frag() {
    init_fragment_abi();
    draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
typedef cs_scale_frag cs_scale_TEXTURE_2D_frag;
WR_PROGRAM(cs_scale_TEXTURE_2D, "cs_scale TEXTURE_2D")
