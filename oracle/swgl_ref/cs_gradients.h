// TEST INFRASTRUCTURE — hand-instantiated SWGL programs for the cached gradient
// render tasks drawn by draw_texture_cache_target (renderer/mod.rs:4085-4183):
//   "cs_fast_linear_gradient" (webrender/res/cs_fast_linear_gradient.glsl)
//   "cs_linear_gradient"      (cs_linear_gradient.glsl + gradient.glsl)
//   "cs_radial_gradient"      (cs_radial_gradient.glsl)
//   "cs_conic_gradient"       (cs_conic_gradient.glsl)
// DITHERING off.
#pragma once

// Attribute plumbing shared by the four programs: aPosition + N flat attributes.
#define WR_CS_ATTRIB_NAMES(...) static const char* names[] = {"aPosition", __VA_ARGS__}

// gradient.glsl:30-61
#define WR_SAMPLE_GRADIENT()                                                                              \
  vec4 sample_gradient(Float offset) {                                                                    \
    offset -= floor(offset) * this->v_gradient_repeat.x;                                                  \
    Float x = clamp(1.0f + offset * 128.0f, Float(0.0f), Float(1.0f + 128.0f));                           \
    Float entry_index = floor(x);                                                                         \
    Float entry_fract = x - entry_index;                                                                  \
    I32 addr = this->v_gradient_address.x + 2 * cast(entry_index);                                        \
    ivec2 uv =                                                                                            \
        ivec2(I32(CONVERT(CONVERT(addr, U32) % 1024U, I32)), I32(CONVERT(CONVERT(addr, U32) / 1024U, I32))); \
    vec4 t0 = texelFetch(this->sGpuBufferF, uv, 0);                                                       \
    vec4 t1 = texelFetch(this->sGpuBufferF, uv + ivec2_scalar(1, 0), 0);                                  \
    return t0 + t1 * entry_fract;                                                                         \
  }

// ---------------------------------------------------------------------------
struct cs_fast_linear_gradient_vert : VertexShaderImpl, WrCommon {
  typedef cs_fast_linear_gradient_vert Self;
  vec2 aPosition;
  vec4_scalar aTaskRect, aColor0, aColor1;
  float aAxisSelect;
  int a_loc[5];
  Float vPos;
  vec4_scalar vColor0, vColor1;
  struct InterpOutputs {
    float vPos;
  };
  cs_fast_linear_gradient_vert() {
    WR_CS_ATTRIB_NAMES("aTaskRect", "aColor0", "aColor1", "aAxisSelect");
    for (int i = 0; i < 5; i++) a_loc[i] = attrib_locations.add(names[i]);
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aTaskRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aColor0, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aColor1, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aAxisSelect, attribs[L[self->a_loc[4]]], start, instance, count);
  }
  // cs_fast_linear_gradient.glsl:18-25
  void main() {
    vPos = mix(Float(0.0f), Float(1.0f), mix(aPosition.x, aPosition.y, Float(aAxisSelect)));
    vColor0 = aColor0;
    vColor1 = aColor1;
    vec2 pos = mix(aTaskRect.sel(X, Y), aTaskRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vPos = get_nth(vPos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_fast_linear_gradient)
};

struct cs_fast_linear_gradient_frag : FragmentShaderImpl, cs_fast_linear_gradient_vert {
  typedef cs_fast_linear_gradient_frag Self;
  typedef cs_fast_linear_gradient_vert::InterpOutputs InterpInputs;
  typedef cs_fast_linear_gradient_vert::InterpOutputs InterpOutputs;
  Float vPos;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vPos = init_interp(init->vPos, step->vPos);
    self->interp_step.vPos = step->vPos * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vPos += interp_step.vPos * chunks;
  }
  // cs_fast_linear_gradient.glsl:28-31
  void main() { gl_FragColor = mix(vec4(vColor0), vec4(vColor1), vPos); }
  WR_FRAGMENT_ABI()
  cs_fast_linear_gradient_frag() { init_fragment_abi(); }
};
WR_PROGRAM(cs_fast_linear_gradient, "cs_fast_linear_gradient")

// ---------------------------------------------------------------------------
struct cs_linear_gradient_vert : VertexShaderImpl, WrCommon {
  typedef cs_linear_gradient_vert Self;
  vec2 aPosition;
  vec4_scalar aTaskRect;
  vec2_scalar aStartPoint, aEndPoint, aScale;
  int aExtendMode, aGradientStopsAddress;
  int a_loc[7];
  vec2 v_pos;
  vec2_scalar v_scale_dir, v_start_offset, v_gradient_repeat;
  ivec2_scalar v_gradient_address;
  struct InterpOutputs {
    vec2_scalar v_pos;
  };
  cs_linear_gradient_vert() {
    WR_CS_ATTRIB_NAMES("aTaskRect", "aStartPoint", "aEndPoint", "aScale", "aExtendMode", "aGradientStopsAddress");
    for (int i = 0; i < 7; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_GpuBufferF;
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aTaskRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aStartPoint, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aEndPoint, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aScale, attribs[L[self->a_loc[4]]], start, instance, count);
    load_flat_attrib(self->aExtendMode, attribs[L[self->a_loc[5]]], start, instance, count);
    load_flat_attrib(self->aGradientStopsAddress, attribs[L[self->a_loc[6]]], start, instance, count);
  }
  // cs_linear_gradient.glsl:25-42
  void main() {
    vec2 pos = mix(aTaskRect.sel(X, Y), aTaskRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
    v_pos = aPosition * vec2(aScale);
    vec2_scalar dir = aEndPoint - aStartPoint;
    v_scale_dir = dir / dot(dir, dir);
    v_start_offset.x = dot(aStartPoint, v_scale_dir);
    v_scale_dir *= (aTaskRect.sel(Z, W) - aTaskRect.sel(X, Y));
    v_gradient_repeat.x = float(aExtendMode == 1);
    v_gradient_address.x = aGradientStopsAddress;
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_pos = get_nth(v_pos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_linear_gradient)
};

#define WR_CS_GRADIENT_FRAG_COMMON(NAME)                                                           \
  typedef NAME##_frag Self;                                                                        \
  typedef NAME##_vert::InterpOutputs InterpInputs;                                                 \
  typedef NAME##_vert::InterpOutputs InterpOutputs;                                                \
  vec2 v_pos;                                                                                      \
  InterpInputs interp_step;                                                                        \
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) { \
    Self* self = (Self*)impl;                                                                      \
    const InterpInputs* init = (const InterpInputs*)init_;                                         \
    const InterpInputs* step = (const InterpInputs*)step_;                                         \
    self->v_pos = init_interp(init->v_pos, step->v_pos);                                           \
    self->interp_step.v_pos = step->v_pos * 4.0f;                                                  \
  }                                                                                                \
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {                                           \
    float chunks = steps * 0.25f;                                                                  \
    v_pos += interp_step.v_pos * chunks;                                                           \
  }                                                                                                \
  WR_SAMPLE_GRADIENT()

struct cs_linear_gradient_frag : FragmentShaderImpl, cs_linear_gradient_vert {
  WR_CS_GRADIENT_FRAG_COMMON(cs_linear_gradient)
  // cs_linear_gradient.glsl:48-53
  void main() {
    Float offset = dot(v_pos, vec2(v_scale_dir)) - v_start_offset.x;
    gl_FragColor = sample_gradient(offset);
  }
  // cs_linear_gradient.glsl:57-65
  void swgl_drawSpanRGBA8() {
    int a = v_gradient_address.x;
    ivec2_scalar uv(int(uint32_t(a) % 1024U), int(uint32_t(a) / 1024U));
    int address = swgl_validateGradient(sGpuBufferF, uv, int(128.0f + 2.0f));
    if (address < 0) return;
    swgl_commitLinearGradientRGBA8(sGpuBufferF, address, 128.0f, false, v_gradient_repeat.x != 0.0f, v_pos,
                                   v_scale_dir, v_start_offset.x);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  cs_linear_gradient_frag() {
    init_fragment_abi();
    draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
WR_PROGRAM(cs_linear_gradient, "cs_linear_gradient")

// ---------------------------------------------------------------------------
struct cs_radial_gradient_vert : VertexShaderImpl, WrCommon {
  typedef cs_radial_gradient_vert Self;
  vec2 aPosition;
  vec4_scalar aTaskRect;
  vec2_scalar aCenter, aScale;
  float aStartRadius, aEndRadius, aXYRatio;
  int aExtendMode, aGradientStopsAddress;
  int a_loc[9];
  vec2 v_pos;
  vec2_scalar v_start_radius, v_gradient_repeat;
  ivec2_scalar v_gradient_address;
  struct InterpOutputs {
    vec2_scalar v_pos;
  };
  cs_radial_gradient_vert() {
    WR_CS_ATTRIB_NAMES("aTaskRect", "aCenter", "aScale", "aStartRadius", "aEndRadius", "aXYRatio", "aExtendMode",
                       "aGradientStopsAddress");
    for (int i = 0; i < 9; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_GpuBufferF;
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aTaskRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aCenter, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aScale, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aStartRadius, attribs[L[self->a_loc[4]]], start, instance, count);
    load_flat_attrib(self->aEndRadius, attribs[L[self->a_loc[5]]], start, instance, count);
    load_flat_attrib(self->aXYRatio, attribs[L[self->a_loc[6]]], start, instance, count);
    load_flat_attrib(self->aExtendMode, attribs[L[self->a_loc[7]]], start, instance, count);
    load_flat_attrib(self->aGradientStopsAddress, attribs[L[self->a_loc[8]]], start, instance, count);
  }
  // cs_radial_gradient.glsl:26-46
  void main() {
    float rd = aEndRadius - aStartRadius;
    float radius_scale = rd != 0.0f ? 1.0f / rd : 0.0f;
    vec2 pos = mix(aTaskRect.sel(X, Y), aTaskRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
    v_start_radius.x = aStartRadius * radius_scale;
    v_pos = (vec2(aTaskRect.sel(Z, W) - aTaskRect.sel(X, Y)) * aPosition * vec2(aScale) - vec2(aCenter)) * radius_scale;
    v_pos.y *= aXYRatio;
    v_gradient_repeat.x = float(aExtendMode == 1);
    v_gradient_address.x = aGradientStopsAddress;
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_pos = get_nth(v_pos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_radial_gradient)
};

struct cs_radial_gradient_frag : FragmentShaderImpl, cs_radial_gradient_vert {
  WR_CS_GRADIENT_FRAG_COMMON(cs_radial_gradient)
  // cs_radial_gradient.glsl:52-57
  void main() {
    Float offset = length(v_pos) - v_start_radius.x;
    gl_FragColor = sample_gradient(offset);
  }
  // cs_radial_gradient.glsl:60-68
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
  cs_radial_gradient_frag() {
    init_fragment_abi();
    draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
WR_PROGRAM(cs_radial_gradient, "cs_radial_gradient")

// ---------------------------------------------------------------------------
struct cs_conic_gradient_vert : VertexShaderImpl, WrCommon {
  typedef cs_conic_gradient_vert Self;
  vec2 aPosition;
  vec4_scalar aTaskRect;
  vec2_scalar aCenter, aScale;
  float aStartOffset, aEndOffset, aAngle;
  int aExtendMode, aGradientStopsAddress;
  int a_loc[9];
  vec2 v_pos;
  vec2_scalar v_center, v_gradient_repeat;
  vec3_scalar v_start_offset_offset_scale_angle_vec;
  ivec2_scalar v_gradient_address;
  struct InterpOutputs {
    vec2_scalar v_pos;
  };
  cs_conic_gradient_vert() {
    WR_CS_ATTRIB_NAMES("aTaskRect", "aCenter", "aScale", "aStartOffset", "aEndOffset", "aAngle", "aExtendMode",
                       "aGradientStopsAddress");
    for (int i = 0; i < 9; i++) a_loc[i] = attrib_locations.add(names[i]);
    sampler_mask |= WR_S_GpuBufferF;
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aTaskRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aCenter, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aScale, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aStartOffset, attribs[L[self->a_loc[4]]], start, instance, count);
    load_flat_attrib(self->aEndOffset, attribs[L[self->a_loc[5]]], start, instance, count);
    load_flat_attrib(self->aAngle, attribs[L[self->a_loc[6]]], start, instance, count);
    load_flat_attrib(self->aExtendMode, attribs[L[self->a_loc[7]]], start, instance, count);
    load_flat_attrib(self->aGradientStopsAddress, attribs[L[self->a_loc[8]]], start, instance, count);
  }
  // cs_conic_gradient.glsl:34-52
  void main() {
    float d = aEndOffset - aStartOffset;
    float offset_scale = d != 0.0f ? 1.0f / d : 0.0f;
    v_start_offset_offset_scale_angle_vec.y = offset_scale;
    vec2 pos = mix(aTaskRect.sel(X, Y), aTaskRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
    v_start_offset_offset_scale_angle_vec.z = 3.141592653589793f / 2.0f - aAngle;
    v_start_offset_offset_scale_angle_vec.x = aStartOffset * offset_scale;
    v_center = aCenter * offset_scale;
    v_pos = vec2(aTaskRect.sel(Z, W) - aTaskRect.sel(X, Y)) * aPosition * offset_scale * vec2(aScale);
    v_gradient_repeat.x = float(aExtendMode == 1);
    v_gradient_address.x = aGradientStopsAddress;
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_pos = get_nth(v_pos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_conic_gradient)
};

struct cs_conic_gradient_frag : FragmentShaderImpl, cs_conic_gradient_vert {
  WR_CS_GRADIENT_FRAG_COMMON(cs_conic_gradient)
  // cs_conic_gradient.glsl:58-66
  void main() {
    vec2 current_dir = v_pos - vec2(v_center);
    Float current_angle = atan(current_dir.y, current_dir.x) + v_start_offset_offset_scale_angle_vec.z;
    Float offset = fract(current_angle / (2.0f * 3.141592653589793f)) * v_start_offset_offset_scale_angle_vec.y -
                   v_start_offset_offset_scale_angle_vec.x;
    gl_FragColor = sample_gradient(offset);
  }
  WR_FRAGMENT_ABI()
  cs_conic_gradient_frag() { init_fragment_abi(); }
};
WR_PROGRAM(cs_conic_gradient, "cs_conic_gradient")
