// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "brush_blend" and
// "brush_blend ALPHA_PASS" (webrender/res/brush_blend.glsl + blend.glsl).
// No span shader: every chunk runs the fragment main.
#pragma once

template <int VARIANT>
struct brush_blend_vert_t : BrushVertBase<brush_blend_vert_t<VARIANT>> {
  typedef brush_blend_vert_t Self;
  static const int VECS_PER_SPECIFIC_BRUSH = 3;
  typedef typename PrimVertBase::VertexInfo VertexInfo;
  typedef WrCommon::RectWithEndpoint RectWithEndpoint;
  typedef WrCommon::PictureTask PictureTask;

  vec2 v_uv;
  vec4_scalar v_uv_sample_bounds;
  vec2_scalar v_perspective_amount;
  ivec2_scalar v_op_table_address_vec;
  mat4_scalar v_color_mat;
  vec4_scalar v_funcs, v_color_offset;
  struct InterpOutputs {
    vec2_scalar v_uv;
  };

  brush_blend_vert_t() {
    this->sampler_mask |= WR_S_Color0;
    this->init_vertex_abi();
  }

  // blend.glsl:26-91
  void SetupFilterParams(int op, float amount, int gpu_data_address) {
    float lumR = 0.2126f, lumG = 0.7152f, lumB = 0.0722f;
    float oneMinusLumR = 1.0f - lumR, oneMinusLumG = 1.0f - lumG, oneMinusLumB = 1.0f - lumB;
    float invAmount = 1.0f - amount;
    if (op == 1) {
      v_color_mat = mat4_scalar(
          vec4_scalar(lumR + oneMinusLumR * invAmount, lumR - lumR * invAmount, lumR - lumR * invAmount, 0.0f),
          vec4_scalar(lumG - lumG * invAmount, lumG + oneMinusLumG * invAmount, lumG - lumG * invAmount, 0.0f),
          vec4_scalar(lumB - lumB * invAmount, lumB - lumB * invAmount, lumB + oneMinusLumB * invAmount, 0.0f),
          vec4_scalar(0.0f, 0.0f, 0.0f, 1.0f));
      v_color_offset = vec4_scalar(0.0f);
    } else if (op == 2) {
      float c = cosf(amount);
      float s = sinf(amount);
      v_color_mat = mat4_scalar(
          vec4_scalar(lumR + oneMinusLumR * c - lumR * s, lumR - lumR * c + 0.143f * s, lumR - lumR * c - oneMinusLumR * s, 0.0f),
          vec4_scalar(lumG - lumG * c - lumG * s, lumG + oneMinusLumG * c + 0.140f * s, lumG - lumG * c + lumG * s, 0.0f),
          vec4_scalar(lumB - lumB * c + oneMinusLumB * s, lumB - lumB * c - 0.283f * s, lumB + oneMinusLumB * c + lumB * s, 0.0f),
          vec4_scalar(0.0f, 0.0f, 0.0f, 1.0f));
      v_color_offset = vec4_scalar(0.0f);
    } else if (op == 4) {
      v_color_mat = mat4_scalar(
          vec4_scalar(invAmount * lumR + amount, invAmount * lumR, invAmount * lumR, 0.0f),
          vec4_scalar(invAmount * lumG, invAmount * lumG + amount, invAmount * lumG, 0.0f),
          vec4_scalar(invAmount * lumB, invAmount * lumB, invAmount * lumB + amount, 0.0f),
          vec4_scalar(0.0f, 0.0f, 0.0f, 1.0f));
      v_color_offset = vec4_scalar(0.0f);
    } else if (op == 5) {
      v_color_mat = mat4_scalar(
          vec4_scalar(0.393f + 0.607f * invAmount, 0.349f - 0.349f * invAmount, 0.272f - 0.272f * invAmount, 0.0f),
          vec4_scalar(0.769f - 0.769f * invAmount, 0.686f + 0.314f * invAmount, 0.534f - 0.534f * invAmount, 0.0f),
          vec4_scalar(0.189f - 0.189f * invAmount, 0.168f - 0.168f * invAmount, 0.131f + 0.869f * invAmount, 0.0f),
          vec4_scalar(0.0f, 0.0f, 0.0f, 1.0f));
      v_color_offset = vec4_scalar(0.0f);
    } else if (op == 7) {
      v_color_mat = mat4_scalar(this->fetch_gpu_cache(gpu_data_address, 0), this->fetch_gpu_cache(gpu_data_address, 1),
                                this->fetch_gpu_cache(gpu_data_address, 2), this->fetch_gpu_cache(gpu_data_address, 3));
      v_color_offset = this->fetch_from_gpu_cache_1(gpu_data_address + 4);
    } else if (op == 11) {
      v_op_table_address_vec.y = gpu_data_address;
    } else if (op == 10) {
      v_color_offset = this->fetch_from_gpu_cache_1(gpu_data_address);
    }
  }

  // brush_blend.glsl:43-89
  void brush_vs(VertexInfo& vi, int, RectWithEndpoint local_rect, RectWithEndpoint, ivec4_scalar prim_user_data,
                int, mat4_scalar, PictureTask&, int brush_flags, vec4_scalar) {
    vec4_scalar r0 = this->fetch_gpu_cache(prim_user_data.x, 0);
    vec2_scalar uv0 = r0.sel(X, Y);
    vec2_scalar uv1 = r0.sel(Z, W);
    vec2_scalar inv_texture_size = vec2_scalar(1.0f) / make_vec2(textureSize(this->sColor0, 0));
    vec2 f = (vi.local_pos - vec2(local_rect.p0)) / vec2(local_rect.p1 - local_rect.p0);
    {
      vec4_scalar st_tl = this->fetch_gpu_cache(prim_user_data.x + 2, 0);
      vec4_scalar st_tr = this->fetch_gpu_cache(prim_user_data.x + 2, 1);
      vec4_scalar st_bl = this->fetch_gpu_cache(prim_user_data.x + 2, 2);
      vec4_scalar st_br = this->fetch_gpu_cache(prim_user_data.x + 2, 3);
      vec4 x = mix(st_tl, st_tr, f.x);
      vec4 y = mix(st_bl, st_br, f.x);
      vec4 z = mix(x, y, f.y);
      f = z.sel(X, Y) / z.w;
    }
    vec2 uv = mix(uv0, uv1, f);
    float perspective_interpolate = (brush_flags & WR_BRUSH_FLAG_PERSPECTIVE_INTERPOLATION) != 0 ? 1.0f : 0.0f;
    v_uv = uv * vec2(inv_texture_size) * mix(vi.world_pos.w, Float(1.0f), Float(perspective_interpolate));
    v_perspective_amount.x = perspective_interpolate;
    v_uv_sample_bounds = make_vec4(uv0 + make_vec2(0.5f), uv1 - make_vec2(0.5f)) * inv_texture_size.sel(X, Y, X, Y);
    float amount = float(prim_user_data.z) / 65536.0f;
    v_op_table_address_vec.x = prim_user_data.y & 0xffff;
    v_perspective_amount.y = amount;
    v_funcs.x = float((prim_user_data.y >> 28) & 0xf);
    v_funcs.y = float((prim_user_data.y >> 24) & 0xf);
    v_funcs.z = float((prim_user_data.y >> 20) & 0xf);
    v_funcs.w = float((prim_user_data.y >> 16) & 0xf);
    SetupFilterParams(v_op_table_address_vec.x, amount, prim_user_data.z);
  }

  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->v_uv = get_nth(v_uv, n);
      dest_ptr += stride;
    }
  }
  using PrimVertBase::load_attribs;
  WR_VERTEX_ABI(brush_blend)
};

template <int VARIANT>
struct brush_blend_frag_t : FragmentShaderImpl, brush_blend_vert_t<VARIANT> {
  typedef brush_blend_frag_t Self;
  typedef typename brush_blend_vert_t<VARIANT>::InterpOutputs InterpInputs;
  typedef typename brush_blend_vert_t<VARIANT>::InterpOutputs InterpOutputs;
  vec2 v_uv;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->v_uv = init_interp(init->v_uv, step->v_uv);
    self->interp_step.v_uv = step->v_uv * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    v_uv += interp_step.v_uv * chunks;
  }
  // draw_perspective: the varyings arrive divided by w and are interpolated linearly in screen space; each chunk
  // multiplies them back by w = 1 / gl_FragCoord.w (what glsl-to-cxx generates next to the plain pair)
  struct InterpPerspective {
    vec2 v_uv;
  };
  InterpPerspective interp_perspective;
  static void read_perspective_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    Float w = 1.0f / self->gl_FragCoord.w;
    self->interp_perspective.v_uv = init_interp(init->v_uv, step->v_uv);
    self->v_uv = self->interp_perspective.v_uv * w;
    self->interp_step.v_uv = step->v_uv * 4.0f;
  }
  ALWAYS_INLINE void step_perspective_inputs(int steps = 4) {
    this->step_perspective(steps);
    float chunks = steps * 0.25f;
    Float w = 1.0f / this->gl_FragCoord.w;
    interp_perspective.v_uv += interp_step.v_uv * chunks;
    v_uv = w * interp_perspective.v_uv;
  }

  // blend.glsl:143-194; per-lane gathers written out lane by lane
  vec4 ComponentTransfer(vec4 colora) {
    int offset = 0;
    int funcs[4] = {int(this->v_funcs.x), int(this->v_funcs.y), int(this->v_funcs.z), int(this->v_funcs.w)};
    int table_address = this->v_op_table_address_vec.y;
    for (int i = 0; i < 4; i++) {
      Float& ci = i == 0 ? colora.x : i == 1 ? colora.y : i == 2 ? colora.z : colora.w;
      switch (funcs[i]) {
        case 0:
          break;
        case 1:
        case 2: {
          I32 k = cast(floor(ci * 255.0f + 0.5f));
          Float out;
          for (int lane = 0; lane < 4; lane++) {
            int kk = k[lane];
            vec4_scalar texel = this->fetch_from_gpu_cache_1(table_address + offset + kk / 4);
            float v = (kk % 4) == 0 ? texel.x : (kk % 4) == 1 ? texel.y : (kk % 4) == 2 ? texel.z : texel.w;
            out[lane] = v;
          }
          ci = clamp(out, Float(0.0f), Float(1.0f));
          offset = offset + 64;
          break;
        }
        case 3: {
          vec4_scalar texel = this->fetch_from_gpu_cache_1(table_address + offset);
          ci = clamp(texel.x * ci + texel.y, Float(0.0f), Float(1.0f));
          offset = offset + 1;
          break;
        }
        case 4: {
          vec4_scalar texel = this->fetch_from_gpu_cache_1(table_address + offset);
          ci = clamp(texel.x * pow(ci, Float(texel.y)) + texel.z, Float(0.0f), Float(1.0f));
          offset = offset + 1;
          break;
        }
        default:
          break;
      }
    }
    return colora;
  }

  // brush_blend.glsl:92-120 + blend.glsl:196-237
  void main() {
    Float perspective_divisor = mix(this->gl_FragCoord.w, Float(1.0f), Float(this->v_perspective_amount.x));
    vec2 uv = v_uv * perspective_divisor;
    uv = clamp(uv, vec2(this->v_uv_sample_bounds.sel(X, Y)), vec2(this->v_uv_sample_bounds.sel(Z, W)));
    vec4 Cs = texture(this->sColor0, uv);
    Float alpha = Cs.w;
    vec3 color = if_then_else(alpha != 0.0f, Cs.sel(X, Y, Z) / alpha, Cs.sel(X, Y, Z));
    float amount = this->v_perspective_amount.y;
    switch (this->v_op_table_address_vec.x) {
      case 0:
        color = clamp(color * amount - 0.5f * amount + 0.5f, Float(0.0f), Float(1.0f));
        break;
      case 3:
        color = mix(color, vec3(Float(1.0f)) - color, Float(amount));
        break;
      case 6:
        color = clamp(color * amount, vec3(Float(0.0f)), vec3(Float(1.0f)));
        break;
      case 8: {
        vec3 c1 = color / 12.92f;
        vec3 c2 = pow(color / 1.055f + vec3(Float(0.055f / 1.055f)), vec3(Float(2.4f)));
        color = if_then_else(lessThanEqual(color, vec3(Float(0.04045f))), c1, c2);
        break;
      }
      case 9: {
        vec3 c1 = color * 12.92f;
        vec3 c2 = vec3(Float(1.055f)) * pow(color, vec3(Float(1.0f / 2.4f))) - vec3(Float(0.055f));
        color = if_then_else(lessThanEqual(color, vec3(Float(0.0031308f))), c1, c2);
        break;
      }
      case 11: {
        vec4 colora = vec4(color, alpha);
        colora = ComponentTransfer(colora);
        color = colora.sel(X, Y, Z);
        alpha = colora.w;
        break;
      }
      case 10:
        color = vec3(this->v_color_offset.sel(X, Y, Z));
        alpha = this->v_color_offset.w;
        break;
      default: {
        vec4 result = this->v_color_mat * vec4(color, alpha) + vec4(this->v_color_offset);
        result = clamp(result, vec4(Float(0.0f)), vec4(Float(1.0f)));
        color = result.sel(X, Y, Z);
        alpha = result.w;
      }
    }
    if (VARIANT == 1) alpha *= 1.0f;  // antialias_brush()
    vec4 frag = alpha * vec4(color, Float(1.0f));
    if (VARIANT == 1) frag *= Float(1.0f);  // do_clip()
    this->gl_FragColor = frag;
  }
  WR_FRAGMENT_ABI_W()
  brush_blend_frag_t() { this->init_fragment_abi(); }
};

typedef brush_blend_frag_t<0> brush_blend_frag;
typedef brush_blend_frag_t<1> brush_blend_ALPHA_PASS_frag;
WR_PROGRAM(brush_blend, "brush_blend")
WR_PROGRAM(brush_blend_ALPHA_PASS, "brush_blend ALPHA_PASS")
