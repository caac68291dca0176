// TEST INFRASTRUCTURE — hand-instantiated SWGL program "ps_quad_textured".
//
// Restates webrender/res/ps_quad_textured.glsl (+ ps_quad.glsl main(),
// sample_color0.glsl) for the SWGL feature set.  Shape follows what
// glsl-to-cxx emits (glsl-to-cxx/src/lib.rs:195-245): <name>_vert holds the
// vertex stage, <name>_frag the fragment stage + swgl_drawSpan*, and
// <name>_program the ProgramImpl glue.

#pragma once

struct ps_quad_textured_vert : PsQuadVertBase {
  typedef ps_quad_textured_vert Self;

  // sample_color0.glsl:7-8
  vec4_scalar v_uv0_sample_bounds;
  vec2 v_uv0;

  struct InterpOutputs {
    vec2_scalar v_uv0;
  };

  ps_quad_textured_vert() {
    sampler_mask |= WR_S_Color0;
    init_vertex_abi();
  }

  // sample_color0.glsl:12-21
  void vs_init_sample_color0(vec2 sample_pos, RectWithEndpoint uv_rect) {
    vec2 uv = mix(uv_rect.p0, uv_rect.p1, sample_pos);
    vec2_scalar texture_size = make_vec2(textureSize(sColor0, 0));
    v_uv0 = uv / vec2(texture_size);
    v_uv0_sample_bounds =
        make_vec4(uv_rect.p0 + make_vec2(0.5f), uv_rect.p1 - make_vec2(0.5f)) /
        texture_size.sel(X, Y, X, Y);
  }

  // ps_quad_textured.glsl:15-36
  void pattern_vertex(PrimitiveInfo& info) {
    if (info.segment.uv_rect.p0 != info.segment.uv_rect.p1) {
      v_flags.x = 1;
      v_color = make_vec4(1.0f);
      vec2 f = (info.local_pos - vec2(info.segment.rect.p0)) /
               vec2(info.segment.rect.p1 - info.segment.rect.p0);
      vs_init_sample_color0(f, info.segment.uv_rect);
    } else {
      v_flags.x = 0;
    }
  }

  // ps_quad.glsl:378-389
  void main() {
    PrimitiveInfo prim = quad_primive_info();
    if ((prim.quad_flags & WR_QF_IS_MASK) != 0) {
      v_flags.z = 1;
    } else {
      v_flags.z = 0;
    }
    pattern_vertex(prim);
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_uv0 = get_nth(v_uv0, n);
      dest_ptr += stride;
    }
  }

  WR_VERTEX_ABI(ps_quad_textured)
};

struct ps_quad_textured_frag : FragmentShaderImpl, ps_quad_textured_vert {
  typedef ps_quad_textured_frag Self;
  typedef ps_quad_textured_vert::InterpOutputs InterpInputs;

  vec2 v_uv0;  // shadows the vertex-side output, as in generated code
  InterpInputs interp_step;

  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_,
                                 const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->v_uv0 = init_interp(init->v_uv0, step->v_uv0);
    self->interp_step.v_uv0 = step->v_uv0 * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    v_uv0 += interp_step.v_uv0 * chunks;
  }
  struct InterpPerspective {
    vec2 v_uv0;
  };
  InterpPerspective interp_perspective;
  static void read_perspective_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    Float w = 1.0f / self->gl_FragCoord.w;
    self->interp_perspective.v_uv0 = init_interp(init->v_uv0, step->v_uv0);
    self->v_uv0 = self->interp_perspective.v_uv0 * w;
    self->interp_step.v_uv0 = step->v_uv0 * 4.0f;
  }
  ALWAYS_INLINE void step_perspective_inputs(int steps = 4) {
    this->step_perspective(steps);
    float chunks = steps * 0.25f;
    Float w = 1.0f / this->gl_FragCoord.w;
    interp_perspective.v_uv0 += interp_step.v_uv0 * chunks;
    v_uv0 = w * interp_perspective.v_uv0;
  }

  // sample_color0.glsl:25-31
  vec4 fs_sample_color0() {
    vec2 uv = clamp(v_uv0, vec2(v_uv0_sample_bounds.sel(X, Y)),
                    vec2(v_uv0_sample_bounds.sel(Z, W)));
    return texture(sColor0, uv);
  }

  // ps_quad_textured.glsl:42-49, ps_quad.glsl:406-417
  void main() {
    vec4 base_color = v_color;
    base_color *= Float(1.0f);  // antialiasing_fragment() under SWGL_ANTIALIAS
    vec4 output_color = base_color;
    if (v_flags.x != 0) {
      vec4 texel = fs_sample_color0();
      output_color *= texel;
    }
    if (v_flags.z != 0) {
      output_color = output_color.sel(X, X, X, X);
    }
    gl_FragColor = output_color;
  }

  // ps_quad_textured.glsl:52-64
  void swgl_drawSpanRGBA8() {
    if (v_flags.x != 0) {
      if (v_flags.z != 0) {
        // falls back to the fragment shader
      } else {
        swgl_commitTextureLinearColorRGBA8(sColor0, v_uv0, v_uv0_sample_bounds,
                                           v_color);
      }
    } else {
      swgl_commitSolidRGBA8(v_color);
    }
  }
  static int draw_span_RGBA8(FragmentShaderImpl* impl) {
    Self* self = (Self*)impl;
    DISPATCH_DRAW_SPAN(self, RGBA8);
  }

  WR_FRAGMENT_ABI_W()

  ps_quad_textured_frag() {
    init_fragment_abi();
    draw_span_RGBA8_func = &draw_span_RGBA8;
  }
};

WR_PROGRAM(ps_quad_textured, "ps_quad_textured")
