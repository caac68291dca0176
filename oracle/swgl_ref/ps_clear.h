// TEST INFRASTRUCTURE — hand-instantiated SWGL program "ps_clear"
// (webrender/res/ps_clear.glsl).  vColor is a real (non-flat) varying.
#pragma once

struct ps_clear_vert : VertexShaderImpl, WrCommon {
  typedef ps_clear_vert Self;
  vec2 aPosition;
  vec4_scalar aRect, aColor;
  int a_loc[3];
  vec4 vColor;
  struct InterpOutputs {
    vec4_scalar vColor;
  };
  ps_clear_vert() {
    static const char* names[3] = {"aPosition", "aRect", "aColor"};
    for (int i = 0; i < 3; i++) a_loc[i] = attrib_locations.add(names[i]);
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aColor, attribs[L[self->a_loc[2]]], start, instance, count);
  }
  // ps_clear.glsl:13-18
  void main() {
    vec2 pos = mix(aRect.sel(X, Y), aRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
    gl_Position.z = gl_Position.w;
    vColor = vec4(aColor);
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vColor = get_nth(vColor, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(ps_clear)
};

struct ps_clear_frag : FragmentShaderImpl, ps_clear_vert {
  typedef ps_clear_frag Self;
  typedef ps_clear_vert::InterpOutputs InterpInputs;
  typedef ps_clear_vert::InterpOutputs InterpOutputs;
  vec4 vColor;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vColor = init_interp(init->vColor, step->vColor);
    self->interp_step.vColor = step->vColor * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vColor += interp_step.vColor * chunks;
  }
  void main() { gl_FragColor = vColor; }
  WR_FRAGMENT_ABI()
  ps_clear_frag() { init_fragment_abi(); }
};
WR_PROGRAM(ps_clear, "ps_clear")
