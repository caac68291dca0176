// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_linear_gradient"
// and "brush_linear_gradient ALPHA_PASS" (webrender/res/brush_linear_gradient.glsl,
// gradient_shared.glsl, gradient.glsl; DITHERING off).  Under SWGL_ANTIALIAS the
// two variants run the same code.
#pragma once

template <int VARIANT>
struct brush_linear_gradient_vert_t : BrushVertBase<brush_linear_gradient_vert_t<VARIANT>> {
  typedef brush_linear_gradient_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 2;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2_scalar v_start_offset, v_scale_dir, v_repeated_size, v_tile_repeat, v_gradient_repeat;
  ivec2_scalar v_gradient_address;
  vec2 v_pos;
  struct InterpOutputs {
    vec2_scalar v_pos;
  };
  brush_linear_gradient_vert_t() {
    this->sampler_mask |= WR_S_GpuBufferF;
    this->init_vertex_abi();
  }
  void brush_vs(VertexInfo& vi, int prim_address, RectWithEndpoint local_rect, RectWithEndpoint segment_rect,
                ivec4_scalar prim_user_data, int, mat4_scalar, PictureTask&, int brush_flags,
                vec4_scalar texel_rect) {
    vec4_scalar g0 = this->fetch_gpu_cache(prim_address, 0);
    vec4_scalar g1 = this->fetch_gpu_cache(prim_address, 1);
    int extend_mode = int(g1.x);
    vec2_scalar stretch_size = g1.sel(Y, Z);
    // write_gradient_vertex, gradient_shared.glsl:19-51
    if ((brush_flags & WR_BRUSH_FLAG_SEGMENT_RELATIVE) != 0) {
      v_pos = (vi.local_pos - vec2(segment_rect.p0)) / vec2(segment_rect.p1 - segment_rect.p0);
      v_pos = v_pos * vec2(texel_rect.sel(Z, W) - texel_rect.sel(X, Y)) + vec2(texel_rect.sel(X, Y));
      v_pos = v_pos * vec2(local_rect.p1 - local_rect.p0);
    } else {
      v_pos = vi.local_pos - vec2(local_rect.p0);
    }
    vec2_scalar tile_repeat = (local_rect.p1 - local_rect.p0) / stretch_size;
    v_repeated_size = stretch_size;
    v_pos /= vec2(v_repeated_size);
    v_gradient_address.x = prim_user_data.x;
    v_gradient_repeat.x = float(extend_mode == 1);
    v_tile_repeat = tile_repeat;
    vec2_scalar start_point = g0.sel(X, Y);
    vec2_scalar end_point = g0.sel(Z, W);
    vec2_scalar dir = end_point - start_point;
    v_scale_dir = dir / dot(dir, dir);
    v_start_offset.x = dot(start_point, v_scale_dir);
    v_scale_dir *= v_repeated_size;
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_pos = get_nth(v_pos, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_linear_gradient)
};

template <int VARIANT>
struct brush_linear_gradient_frag_t : FragmentShaderImpl, brush_linear_gradient_vert_t<VARIANT> {
  typedef brush_linear_gradient_frag_t Self;
  typedef typename brush_linear_gradient_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_linear_gradient_vert_t<VARIANT>::InterpOutputs InterpOutputs;
  vec2 v_pos;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->v_pos = init_interp(init->v_pos, step->v_pos);
    self->interp_step.v_pos = step->v_pos * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    v_pos += interp_step.v_pos * chunks;
  }
  // gradient.glsl:30-61
  vec4 sample_gradient(Float offset) {
    offset -= floor(offset) * this->v_gradient_repeat.x;
    Float x = clamp(1.0f + offset * 128.0f, Float(0.0f), Float(1.0f + 128.0f));
    Float entry_index = floor(x);
    Float entry_fract = x - entry_index;
    I32 addr = this->v_gradient_address.x + 2 * cast(entry_index);
    ivec2 uv = ivec2(I32(CONVERT(CONVERT(addr, U32) % 1024U, I32)), I32(CONVERT(CONVERT(addr, U32) / 1024U, I32)));
    vec4 t0 = texelFetch(this->sGpuBufferF, uv, 0);
    vec4 t1 = texelFetch(this->sGpuBufferF, uv + ivec2_scalar(1, 0), 0);
    return t0 + t1 * entry_fract;
  }
  void main() {
    vec2 pos = fract(v_pos);  // compute_repeated_pos under SWGL_ANTIALIAS
    Float offset = dot(pos, vec2(this->v_scale_dir)) - this->v_start_offset.x;
    vec4 color = sample_gradient(offset);
    if (VARIANT == 1) color *= Float(1.0f);
    this->gl_FragColor = color;
  }
  void swgl_drawSpanRGBA8() {
    int a = this->v_gradient_address.x;
    ivec2_scalar uv(int(uint32_t(a) % 1024U), int(uint32_t(a) / 1024U));
    int address = swgl_validateGradient(this->sGpuBufferF, uv, int(128.0f + 2.0f));
    if (address < 0) return;
    swgl_commitLinearGradientRGBA8(this->sGpuBufferF, address, 128.0f, true, this->v_gradient_repeat.x != 0.0f,
                                   v_pos, this->v_scale_dir, this->v_start_offset.x);
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }
  WR_FRAGMENT_ABI()
  brush_linear_gradient_frag_t() {
    this->init_fragment_abi();
    this->draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};
typedef brush_linear_gradient_frag_t<0> brush_linear_gradient_frag;
typedef brush_linear_gradient_frag_t<1> brush_linear_gradient_ALPHA_PASS_frag;
WR_PROGRAM(brush_linear_gradient, "brush_linear_gradient")
WR_PROGRAM(brush_linear_gradient_ALPHA_PASS, "brush_linear_gradient ALPHA_PASS")
