// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_mix_blend" and
// "brush_mix_blend ALPHA_PASS" (webrender/res/brush_mix_blend.glsl).  No span
// shader.  The translator vectorises the scalar blend functions with masks;
// here the four lanes are evaluated one by one with the same float ops.
#pragma once

namespace wr_mix {
static inline float Lum(const float* c) { return c[0] * 0.3f + c[1] * 0.59f + c[2] * 0.11f; }
static inline float Sat(const float* c) {
  return fmaxf(c[0], fmaxf(c[1], c[2])) - fminf(c[0], fminf(c[1], c[2]));
}
static inline void ClipColor(float* C) {
  float L = Lum(C);
  float n = fminf(C[0], fminf(C[1], C[2]));
  float x = fmaxf(C[0], fmaxf(C[1], C[2]));
  if (n < 0.0f)
    for (int i = 0; i < 3; i++) C[i] = L + (((C[i] - L) * L) / (L - n));
  if (x > 1.0f)
    for (int i = 0; i < 3; i++) C[i] = L + (((C[i] - L) * (1.0f - L)) / (x - L));
}
static inline void SetLum(const float* C, float l, float* out) {
  float d = l - Lum(C);
  for (int i = 0; i < 3; i++) out[i] = C[i] + d;
  ClipColor(out);
}
static inline void SetSatInner(float& Cmin, float& Cmid, float& Cmax, float s) {
  if (Cmax > Cmin) {
    Cmid = (((Cmid - Cmin) * s) / (Cmax - Cmin));
    Cmax = s;
  } else {
    Cmid = 0.0f;
    Cmax = 0.0f;
  }
  Cmin = 0.0f;
}
static inline void SetSat(float* C, float s) {
  if (C[0] <= C[1]) {
    if (C[1] <= C[2]) SetSatInner(C[0], C[1], C[2], s);
    else if (C[0] <= C[2]) SetSatInner(C[0], C[2], C[1], s);
    else SetSatInner(C[2], C[0], C[1], s);
  } else {
    if (C[0] <= C[2]) SetSatInner(C[1], C[0], C[2], s);
    else if (C[1] <= C[2]) SetSatInner(C[1], C[2], C[0], s);
    else SetSatInner(C[2], C[1], C[0], s);
  }
}
static inline float ColorDodge(float Cb, float Cs) {
  if (Cb == 0.0f) return 0.0f;
  else if (Cs == 1.0f) return 1.0f;
  else return fminf(1.0f, Cb / (1.0f - Cs));
}
static inline float ColorBurn(float Cb, float Cs) {
  if (Cb == 1.0f) return 1.0f;
  else if (Cs == 0.0f) return 0.0f;
  else return 1.0f - fminf(1.0f, (1.0f - Cb) / Cs);
}
static inline float SoftLight(float Cb, float Cs) {
  if (Cs <= 0.5f) {
    return Cb - (1.0f - 2.0f * Cs) * Cb * (1.0f - Cb);
  } else {
    float D;
    if (Cb <= 0.25f) D = ((16.0f * Cb - 12.0f) * Cb + 4.0f) * Cb;
    else D = sqrtf(Cb);
    return Cb + (2.0f * Cs - 1.0f) * (D - Cb);
  }
}
static inline void HardLight(const float* Cb, const float* Cs, float* out) {
  for (int i = 0; i < 3; i++) {
    float m = Cb[i] * (2.0f * Cs[i]);
    float s2 = 2.0f * Cs[i] - 1.0f;
    float s = Cb[i] + s2 - (Cb[i] * s2);
    float st = Cs[i] >= 0.5f ? 1.0f : 0.0f;
    out[i] = (s - m) * st + m;  // mix(m, s, step(edge, Cs))
  }
}
// brush_fs (brush_mix_blend.glsl:230-331) for one lane; Cb/Cs premultiplied RGBA in, result out
static inline void fragment(int op, float* Cb, float* Cs, float* result) {
  if (Cb[3] != 0.0f) for (int i = 0; i < 3; i++) Cb[i] /= Cb[3];
  if (Cs[3] != 0.0f) for (int i = 0; i < 3; i++) Cs[i] /= Cs[3];
  result[0] = 1.0f; result[1] = 1.0f; result[2] = 0.0f; result[3] = 1.0f;
  float t[3];
  switch (op & 0xFF) {
    case 1: for (int i = 0; i < 3; i++) result[i] = Cb[i] * Cs[i]; break;
    case 3: HardLight(Cs, Cb, result); break;
    case 4: for (int i = 0; i < 3; i++) result[i] = fminf(Cs[i], Cb[i]); break;
    case 5: for (int i = 0; i < 3; i++) result[i] = fmaxf(Cs[i], Cb[i]); break;
    case 6: for (int i = 0; i < 3; i++) result[i] = ColorDodge(Cb[i], Cs[i]); break;
    case 7: for (int i = 0; i < 3; i++) result[i] = ColorBurn(Cb[i], Cs[i]); break;
    case 8: HardLight(Cb, Cs, result); break;
    case 9: for (int i = 0; i < 3; i++) result[i] = SoftLight(Cb[i], Cs[i]); break;
    case 10: for (int i = 0; i < 3; i++) result[i] = fabsf(Cb[i] - Cs[i]); break;
    case 12: t[0] = Cs[0]; t[1] = Cs[1]; t[2] = Cs[2]; SetSat(t, Sat(Cb)); SetLum(t, Lum(Cb), result); break;
    case 13: t[0] = Cb[0]; t[1] = Cb[1]; t[2] = Cb[2]; SetSat(t, Sat(Cs)); SetLum(t, Lum(Cb), result); break;
    case 14: SetLum(Cs, Lum(Cb), result); break;
    case 15: SetLum(Cb, Lum(Cs), result); break;
    default: break;
  }
  for (int i = 0; i < 3; i++) result[i] = (1.0f - Cb[3]) * Cs[i] + Cb[3] * result[i];
  result[3] = Cs[3];
  for (int i = 0; i < 3; i++) result[i] *= result[3];
}
}  // namespace wr_mix

template <int VARIANT>
struct brush_mix_blend_vert_t : BrushVertBase<brush_mix_blend_vert_t<VARIANT>> {
  typedef brush_mix_blend_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 3;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2 v_src_uv, v_backdrop_uv;
  vec4_scalar v_src_uv_sample_bounds, v_backdrop_uv_sample_bounds;
  vec2_scalar v_perspective;
  ivec2_scalar v_op;
  struct InterpOutputs {
    vec2_scalar v_src_uv;
    vec2_scalar v_backdrop_uv;
  };

  brush_mix_blend_vert_t() {
    this->sampler_mask |= WR_S_Color0 | WR_S_Color1;
    this->init_vertex_abi();
  }

  // brush_mix_blend.glsl:25-45
  void get_uv(int res_address, vec2 f, ivec2_scalar texture_size, Float perspective_f, vec2& out_uv,
              vec4_scalar& out_bounds) {
    vec4_scalar r0 = this->fetch_gpu_cache(res_address, 0);
    vec2_scalar uv0 = r0.sel(X, Y), uv1 = r0.sel(Z, W);
    vec2_scalar inv_texture_size = vec2_scalar(1.0f) / make_vec2(texture_size);
    vec4_scalar st_tl = this->fetch_gpu_cache(res_address + 2, 0);
    vec4_scalar st_tr = this->fetch_gpu_cache(res_address + 2, 1);
    vec4_scalar st_bl = this->fetch_gpu_cache(res_address + 2, 2);
    vec4_scalar st_br = this->fetch_gpu_cache(res_address + 2, 3);
    vec4 x = mix(st_tl, st_tr, f.x);
    vec4 y = mix(st_bl, st_br, f.x);
    vec4 z = mix(x, y, f.y);
    f = z.sel(X, Y) / z.w;
    vec2 uv = mix(uv0, uv1, f);
    out_uv = uv * vec2(inv_texture_size) * perspective_f;
    out_bounds = make_vec4(uv0 + make_vec2(0.5f), uv1 - make_vec2(0.5f)) * inv_texture_size.sel(X, Y, X, Y);
  }

  // brush_mix_blend.glsl:47-83
  void brush_vs(VertexInfo& vi, int, RectWithEndpoint local_rect, RectWithEndpoint, ivec4_scalar prim_user_data,
                int, mat4_scalar, PictureTask&, int brush_flags, vec4_scalar) {
    vec2 f = (vi.local_pos - vec2(local_rect.p0)) / vec2(local_rect.p1 - local_rect.p0);
    float perspective_interpolate = (brush_flags & WR_BRUSH_FLAG_PERSPECTIVE_INTERPOLATION) != 0 ? 1.0f : 0.0f;
    Float perspective_f = mix(vi.world_pos.w, Float(1.0f), Float(perspective_interpolate));
    v_perspective.x = perspective_interpolate;
    v_op.x = prim_user_data.x;
    get_uv(prim_user_data.y, f, textureSize(this->sColor0, 0), Float(1.0f), v_backdrop_uv, v_backdrop_uv_sample_bounds);
    get_uv(prim_user_data.z, f, textureSize(this->sColor1, 0), perspective_f, v_src_uv, v_src_uv_sample_bounds);
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_src_uv = get_nth(v_src_uv, n);
      dest->v_backdrop_uv = get_nth(v_backdrop_uv, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_mix_blend)
};

template <int VARIANT>
struct brush_mix_blend_frag_t : FragmentShaderImpl, brush_mix_blend_vert_t<VARIANT> {
  typedef brush_mix_blend_frag_t Self;
  typedef typename brush_mix_blend_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_mix_blend_vert_t<VARIANT>::InterpOutputs InterpOutputs;
  vec2 v_src_uv, v_backdrop_uv;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->v_src_uv = init_interp(init->v_src_uv, step->v_src_uv);
    self->interp_step.v_src_uv = step->v_src_uv * 4.0f;
    self->v_backdrop_uv = init_interp(init->v_backdrop_uv, step->v_backdrop_uv);
    self->interp_step.v_backdrop_uv = step->v_backdrop_uv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    v_src_uv += interp_step.v_src_uv * chunks;
    v_backdrop_uv += interp_step.v_backdrop_uv * chunks;
  }
  // draw_perspective: the varyings arrive divided by w and are interpolated linearly in screen space; each chunk
  // multiplies them back by w = 1 / gl_FragCoord.w (what glsl-to-cxx generates next to the plain pair)
  struct InterpPerspective {
    vec2 v_src_uv;
    vec2 v_backdrop_uv;
  };
  InterpPerspective interp_perspective;
  static void read_perspective_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    Float w = 1.0f / self->gl_FragCoord.w;
    self->interp_perspective.v_src_uv = init_interp(init->v_src_uv, step->v_src_uv);
    self->v_src_uv = self->interp_perspective.v_src_uv * w;
    self->interp_step.v_src_uv = step->v_src_uv * 4.0f;
    self->interp_perspective.v_backdrop_uv = init_interp(init->v_backdrop_uv, step->v_backdrop_uv);
    self->v_backdrop_uv = self->interp_perspective.v_backdrop_uv * w;
    self->interp_step.v_backdrop_uv = step->v_backdrop_uv * 4.0f;
  }
  ALWAYS_INLINE void step_perspective_inputs(int steps = 4) {
    this->step_perspective(steps);
    float chunks = steps * 0.25f;
    Float w = 1.0f / this->gl_FragCoord.w;
    interp_perspective.v_src_uv += interp_step.v_src_uv * chunks;
    v_src_uv = w * interp_perspective.v_src_uv;
    interp_perspective.v_backdrop_uv += interp_step.v_backdrop_uv * chunks;
    v_backdrop_uv = w * interp_perspective.v_backdrop_uv;
  }

  void main() {
    Float perspective_divisor = mix(this->gl_FragCoord.w, Float(1.0f), Float(this->v_perspective.x));
    vec2 src_uv = v_src_uv * perspective_divisor;
    src_uv = clamp(src_uv, vec2(this->v_src_uv_sample_bounds.sel(X, Y)), vec2(this->v_src_uv_sample_bounds.sel(Z, W)));
    vec2 backdrop_uv = clamp(v_backdrop_uv, vec2(this->v_backdrop_uv_sample_bounds.sel(X, Y)),
                             vec2(this->v_backdrop_uv_sample_bounds.sel(Z, W)));
    vec4 Cb = texture(this->sColor0, backdrop_uv);
    vec4 Cs = texture(this->sColor1, src_uv);
    vec4 result;
    for (int lane = 0; lane < 4; lane++) {
      float cb[4] = {Cb.x[lane], Cb.y[lane], Cb.z[lane], Cb.w[lane]};
      float cs[4] = {Cs.x[lane], Cs.y[lane], Cs.z[lane], Cs.w[lane]};
      float r[4];
      wr_mix::fragment(this->v_op.x, cb, cs, r);
      result.x[lane] = r[0]; result.y[lane] = r[1]; result.z[lane] = r[2]; result.w[lane] = r[3];
    }
    if (VARIANT == 1) result *= Float(1.0f);  // do_clip()
    this->gl_FragColor = result;
  }
  WR_FRAGMENT_ABI_W()
  brush_mix_blend_frag_t() { this->init_fragment_abi(); }
};

typedef brush_mix_blend_frag_t<0> brush_mix_blend_frag;
typedef brush_mix_blend_frag_t<1> brush_mix_blend_ALPHA_PASS_frag;
WR_PROGRAM(brush_mix_blend, "brush_mix_blend")
WR_PROGRAM(brush_mix_blend_ALPHA_PASS, "brush_mix_blend ALPHA_PASS")
