// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "ps_quad_radial_gradient"
// and "ps_quad_conic_gradient" (webrender/res/ps_quad_radial_gradient.glsl,
// ps_quad_conic_gradient.glsl + ps_quad.glsl main(), gradient.glsl; DITHERING off).
#pragma once

#define WR_QUAD_GRADIENT_FRAG_COMMON(NAME, VARY)                                                   \
  typedef NAME##_frag Self;                                                                        \
  typedef NAME##_vert::InterpOutputs InterpInputs;                                                 \
  vec2 VARY;                                                                                       \
  InterpInputs interp_step;                                                                        \
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) { \
    Self* self = (Self*)impl;                                                                      \
    const InterpInputs* init = (const InterpInputs*)init_;                                         \
    const InterpInputs* step = (const InterpInputs*)step_;                                         \
    self->VARY = init_interp(init->VARY, step->VARY);                                              \
    self->interp_step.VARY = step->VARY * 4.0f;                                                    \
  }                                                                                                \
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {                                           \
    float chunks = steps * 0.25f;                                                                  \
    VARY += interp_step.VARY * chunks;                                                             \
  }                                                                                                \
  WR_SAMPLE_GRADIENT()

struct ps_quad_radial_gradient_vert : PsQuadVertBase {
  typedef ps_quad_radial_gradient_vert Self;
  vec2_scalar v_start_radius, v_gradient_repeat;
  ivec2_scalar v_gradient_address;
  vec2 v_pos;
  struct InterpOutputs {
    vec2_scalar v_pos;
  };
  ps_quad_radial_gradient_vert() {
    sampler_mask |= WR_S_GpuBufferF;
    init_vertex_abi();
  }
  // ps_quad_radial_gradient.glsl:37-58
  void pattern_vertex(PrimitiveInfo& info) {
    vec4_scalar d0 = fetch_gpu_buffer_f(info.pattern_input.x, 0);
    vec4_scalar d1 = fetch_gpu_buffer_f(info.pattern_input.x, 1);
    vec2_scalar center = d0.sel(X, Y), scale = d0.sel(Z, W);
    float start_radius = d1.x, end_radius = d1.y, xy_ratio = d1.z, repeat = d1.w;
    v_gradient_address.x = info.pattern_input.y;
    float rd = end_radius - start_radius;
    float radius_scale = rd != 0.0f ? 1.0f / rd : 0.0f;
    v_start_radius.x = start_radius * radius_scale;
    v_pos = ((info.local_pos - vec2(info.local_prim_rect.p0)) * vec2(scale) - vec2(center)) * radius_scale;
    v_pos.y *= xy_ratio;
    v_gradient_repeat.x = repeat;
  }
  void main() {
    PrimitiveInfo prim = quad_primive_info();
    v_flags.z = (prim.quad_flags & WR_QF_IS_MASK) != 0 ? 1 : 0;
    pattern_vertex(prim);
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_pos = get_nth(v_pos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(ps_quad_radial_gradient)
};

struct ps_quad_radial_gradient_frag : FragmentShaderImpl, ps_quad_radial_gradient_vert {
  WR_QUAD_GRADIENT_FRAG_COMMON(ps_quad_radial_gradient, v_pos)
  // ps_quad_radial_gradient.glsl:62-68, ps_quad.glsl:406-417
  void main() {
    vec4 base_color = v_color;
    base_color *= Float(1.0f);
    Float offset = length(v_pos) - v_start_radius.x;
    vec4 output_color = base_color;
    output_color *= sample_gradient(offset);
    if (v_flags.z != 0) output_color = output_color.sel(X, X, X, X);
    gl_FragColor = output_color;
  }
  // ps_quad_radial_gradient.glsl:71-79
  void swgl_drawSpanRGBA8() {
    int a = v_gradient_address.x;
    ivec2_scalar uv(int(uint32_t(a) % 1024U), int(uint32_t(a) / 1024U));
    int address = swgl_validateGradient(sGpuBufferF, uv, int(128.0f + 2.0f));
    if (address < 0) return;
    swgl_commitRadialGradientRGBA8(sGpuBufferF, address, 128.0f, v_gradient_repeat.x != 0.0f, v_pos,
                                   v_start_radius.x);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  ps_quad_radial_gradient_frag() {
    init_fragment_abi();
    draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
WR_PROGRAM(ps_quad_radial_gradient, "ps_quad_radial_gradient")

struct ps_quad_conic_gradient_vert : PsQuadVertBase {
  typedef ps_quad_conic_gradient_vert Self;
  vec3_scalar v_start_offset_offset_scale_angle_vec;
  vec2_scalar v_gradient_repeat;
  ivec2_scalar v_gradient_address;
  vec2 v_dir;
  struct InterpOutputs {
    vec2_scalar v_dir;
  };
  ps_quad_conic_gradient_vert() {
    sampler_mask |= WR_S_GpuBufferF;
    init_vertex_abi();
  }
  // ps_quad_conic_gradient.glsl:46-60
  void pattern_vertex(PrimitiveInfo& info) {
    vec4_scalar d0 = fetch_gpu_buffer_f(info.pattern_input.x, 0);
    vec4_scalar d1 = fetch_gpu_buffer_f(info.pattern_input.x, 1);
    vec2_scalar center = d0.sel(X, Y), scale = d0.sel(Z, W);
    float start_offset = d1.x, end_offset = d1.y, angle = d1.z, repeat = d1.w;
    v_gradient_address.x = info.pattern_input.y;
    v_gradient_repeat.x = repeat;
    float d = end_offset - start_offset;
    float offset_scale = d != 0.0f ? 1.0f / d : 0.0f;
    v_start_offset_offset_scale_angle_vec.y = offset_scale;
    v_start_offset_offset_scale_angle_vec.z = 3.141592653589793f / 2.0f - angle;
    v_start_offset_offset_scale_angle_vec.x = start_offset * offset_scale;
    v_dir = (info.local_pos - vec2(info.local_prim_rect.p0)) * vec2(scale) - vec2(center);
  }
  void main() {
    PrimitiveInfo prim = quad_primive_info();
    v_flags.z = (prim.quad_flags & WR_QF_IS_MASK) != 0 ? 1 : 0;
    pattern_vertex(prim);
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_dir = get_nth(v_dir, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(ps_quad_conic_gradient)
};

struct ps_quad_conic_gradient_frag : FragmentShaderImpl, ps_quad_conic_gradient_vert {
  WR_QUAD_GRADIENT_FRAG_COMMON(ps_quad_conic_gradient, v_dir)
  // ps_quad_conic_gradient.glsl:67-82; if_then_else(c, a, b) = mix(b, a, c) (shared.glsl:205)
  static Float approx_atan2(Float y, Float x) {
    vec2 a = abs(vec2(x, y));
    Float slope = min(a.x, a.y) / max(a.x, a.y);
    Float s2 = slope * slope;
    Float r = ((-0.0464964749f * s2 + 0.15931422f) * s2 - 0.327622764f) * s2 * slope + slope;
    r = mix(r, 1.57079637f - r, if_then_else(a.y > a.x, Float(1.0f), Float(0.0f)));
    r = mix(r, 3.14159274f - r, if_then_else(x < 0.0f, Float(1.0f), Float(0.0f)));
    r = r * sign(y);
    return r;
  }
  // ps_quad_conic_gradient.glsl:84-92, ps_quad.glsl:406-417
  void main() {
    vec4 base_color = v_color;
    base_color *= Float(1.0f);
    vec2 current_dir = v_dir;
    Float current_angle = approx_atan2(current_dir.y, current_dir.x) + v_start_offset_offset_scale_angle_vec.z;
    Float offset = fract(current_angle / (2.0f * 3.141592653589793f)) * v_start_offset_offset_scale_angle_vec.y -
                   v_start_offset_offset_scale_angle_vec.x;
    vec4 output_color = base_color;
    output_color *= sample_gradient(offset);
    if (v_flags.z != 0) output_color = output_color.sel(X, X, X, X);
    gl_FragColor = output_color;
  }
  WR_FRAGMENT_ABI()
  ps_quad_conic_gradient_frag() { init_fragment_abi(); }
};
WR_PROGRAM(ps_quad_conic_gradient, "ps_quad_conic_gradient")
