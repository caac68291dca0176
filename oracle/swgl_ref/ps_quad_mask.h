// TEST INFRASTRUCTURE — hand-instantiated SWGL programs "ps_quad_mask" and
// "ps_quad_mask FAST_PATH" (webrender/res/ps_quad_mask.glsl + ps_quad.glsl main
// + ellipse.glsl).  No swgl_drawSpan*: every chunk runs the fragment shader.
#pragma once

template <bool FAST>
struct ps_quad_mask_vert_t : PsQuadVertBase {
  typedef ps_quad_mask_vert_t Self;
  ivec4_scalar aClipData;
  int a_aClipData;
  vec4 vClipLocalPos;
  vec3_scalar v_clip_params;
  vec4_scalar vClipCenter_Radius_TL, vClipCenter_Radius_TR, vClipCenter_Radius_BR, vClipCenter_Radius_BL;
  vec4_scalar vClipPlane_A, vClipPlane_B, vClipPlane_C;
  vec4_scalar vTransformBounds;
  vec2_scalar vClipMode;
  struct InterpOutputs {
    vec4_scalar vClipLocalPos;
  };

  ps_quad_mask_vert_t() {
    a_aClipData = attrib_locations.add("aClipData");
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    PsQuadVertBase::load_attribs(impl, attribs, start, instance, count);
    load_flat_attrib(self->aClipData, attribs[self->attrib_locations.locs[self->a_aClipData]], start, instance,
                     count);
  }

  // ps_quad_mask.glsl:66-137
  void pattern_vertex(PrimitiveInfo& prim_info) {
    int index = aClipData.y;
    int space = aClipData.z;
    RectWithEndpoint rect;
    vec4_scalar radii, radii_top, radii_bottom;
    float mode;
    vec4_scalar t0 = fetch_gpu_buffer_f(index, 0);
    rect = RectWithEndpoint{t0.sel(X, Y), t0.sel(Z, W)};
    if (FAST) {
      radii = fetch_gpu_buffer_f(index, 1);
      mode = fetch_gpu_buffer_f(index, 2).x;
    } else {
      radii_top = fetch_gpu_buffer_f(index, 1);
      radii_bottom = fetch_gpu_buffer_f(index, 2);
      mode = fetch_gpu_buffer_f(index, 3).x;
    }
    Transform clip_transform = fetch_transform(aClipData.x);
    vClipLocalPos = clip_transform.m * vec4(prim_info.local_pos, Float(0.0f), Float(1.0f));
    if (!FAST) {
      if (space == 0) {
        vTransformBounds = make_vec4(rect.p0, rect.p1);
      } else {
        RectWithEndpoint xf = RectWithEndpoint{max(rect.p0, prim_info.local_clip_rect.p0),
                                               min(rect.p1, prim_info.local_clip_rect.p1)};
        vTransformBounds = make_vec4(xf.p0, xf.p1);
      }
    }
    vClipMode.x = mode;
    if (FAST) {
      vec2_scalar half_size = 0.5f * (rect.p1 - rect.p0);
      float radius = radii.x;
      vec2 sub = (half_size + rect.p0) * vClipLocalPos.w;
      vClipLocalPos.x -= sub.x;
      vClipLocalPos.y -= sub.y;
      vec2_scalar hs = half_size - vec2_scalar(radius);
      v_clip_params = vec3_scalar(hs.x, hs.y, radius);
    } else {
      vec2_scalar r_tl = radii_top.sel(X, Y), r_tr = radii_top.sel(Z, W);
      vec2_scalar r_br = radii_bottom.sel(Z, W), r_bl = radii_bottom.sel(X, Y);
      vClipCenter_Radius_TL = make_vec4(rect.p0 + r_tl, wr_inverse_radii_squared(r_tl));
      vClipCenter_Radius_TR = make_vec4(rect.p1.x - r_tr.x, rect.p0.y + r_tr.y, wr_inverse_radii_squared(r_tr));
      vClipCenter_Radius_BR = make_vec4(rect.p1 - r_br, wr_inverse_radii_squared(r_br));
      vClipCenter_Radius_BL = make_vec4(rect.p0.x + r_bl.x, rect.p1.y - r_bl.y, wr_inverse_radii_squared(r_bl));
      vec2_scalar n_tl = -r_tl.sel(Y, X);
      vec2_scalar n_tr = vec2_scalar(r_tr.y, -r_tr.x);
      vec2_scalar n_br = r_br.sel(Y, X);
      vec2_scalar n_bl = vec2_scalar(-r_bl.y, r_bl.x);
      vec3_scalar tl(n_tl.x, n_tl.y, dot(n_tl, vec2_scalar(rect.p0.x, rect.p0.y + r_tl.y)));
      vec3_scalar tr(n_tr.x, n_tr.y, dot(n_tr, vec2_scalar(rect.p1.x - r_tr.x, rect.p0.y)));
      vec3_scalar br(n_br.x, n_br.y, dot(n_br, vec2_scalar(rect.p1.x, rect.p1.y - r_br.y)));
      vec3_scalar bl(n_bl.x, n_bl.y, dot(n_bl, vec2_scalar(rect.p0.x + r_bl.x, rect.p1.y)));
      vClipPlane_A = vec4_scalar(tl.x, tl.y, tl.z, tr.x);
      vClipPlane_B = vec4_scalar(tr.y, tr.z, br.x, br.y);
      vClipPlane_C = vec4_scalar(br.z, bl.x, bl.y, bl.z);
    }
  }

  void main() {
    PrimitiveInfo prim = quad_primive_info();
    v_flags.z = (prim.quad_flags & WR_QF_IS_MASK) != 0 ? 1 : 0;
    pattern_vertex(prim);
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vClipLocalPos = get_nth(vClipLocalPos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(ps_quad_mask)
};

template <bool FAST>
struct ps_quad_mask_frag_t : FragmentShaderImpl, ps_quad_mask_vert_t<FAST> {
  typedef ps_quad_mask_frag_t Self;
  typedef typename ps_quad_mask_vert_t<FAST>::InterpOutputs InterpInputs;
  typedef typename ps_quad_mask_vert_t<FAST>::InterpOutputs InterpOutputs;
  vec4 vClipLocalPos;
  InterpInputs interp_step;
  static void read_interp_inputs(FragmentShaderImpl* impl, const void* init_, const void* step_) {
    Self* self = (Self*)impl;
    const InterpInputs* init = (const InterpInputs*)init_;
    const InterpInputs* step = (const InterpInputs*)step_;
    self->vClipLocalPos = init_interp(init->vClipLocalPos, step->vClipLocalPos);
    self->interp_step.vClipLocalPos = step->vClipLocalPos * 4.0f;
  }
  ALWAYS_INLINE void step_interp_inputs(int steps = 4) {
    float chunks = steps * 0.25f;
    vClipLocalPos += interp_step.vClipLocalPos * chunks;
  }
  typedef cs_clip_rectangle_frag_t<FAST> CR;  // shares the distance helpers

  Float distance_to_rounded_rect(vec2 pos) {
    vec3_scalar ptl(this->vClipPlane_A.x, this->vClipPlane_A.y, this->vClipPlane_A.z);
    vec3_scalar ptr_(this->vClipPlane_A.w, this->vClipPlane_B.x, this->vClipPlane_B.y);
    vec3_scalar pbr(this->vClipPlane_B.z, this->vClipPlane_B.w, this->vClipPlane_C.x);
    vec3_scalar pbl(this->vClipPlane_C.y, this->vClipPlane_C.z, this->vClipPlane_C.w);
    vec4 corner = vec4(vec4_scalar(1.0e-6f, 1.0e-6f, 1.0f, 1.0f));
    vec4 crtl = vec4(this->vClipCenter_Radius_TL), crtr = vec4(this->vClipCenter_Radius_TR);
    vec4 crbr = vec4(this->vClipCenter_Radius_BR), crbl = vec4(this->vClipCenter_Radius_BL);
    vec2 t;
    t = crtl.sel(X, Y) - pos; crtl.x = t.x; crtl.y = t.y;
    t = (crtr.sel(X, Y) - pos) * vec2_scalar(-1.0f, 1.0f); crtr.x = t.x; crtr.y = t.y;
    t = pos - crbr.sel(X, Y); crbr.x = t.x; crbr.y = t.y;
    t = (crbl.sel(X, Y) - pos) * vec2_scalar(1.0f, -1.0f); crbl.x = t.x; crbl.y = t.y;
    corner = if_then_else(dot(pos, vec2(ptl.sel(X, Y))) > ptl.z, crtl, corner);
    corner = if_then_else(dot(pos, vec2(ptr_.sel(X, Y))) > ptr_.z, crtr, corner);
    corner = if_then_else(dot(pos, vec2(pbr.sel(X, Y))) > pbr.z, crbr, corner);
    corner = if_then_else(dot(pos, vec2(pbl.sel(X, Y))) > pbl.z, crbl, corner);
    return max(CR::distance_to_ellipse_approx(corner.sel(X, Y), corner.sel(Z, W), 1.0f),
               CR::signed_distance_rect(pos, this->vTransformBounds.sel(X, Y), this->vTransformBounds.sel(Z, W)));
  }

  // ps_quad.glsl:406-417 + ps_quad_mask.glsl:152-180
  void main() {
    vec2 clip_local_pos = vClipLocalPos.sel(X, Y) / vClipLocalPos.w;
    float aa_range = CR::compute_aa_range(clip_local_pos);
    Float dist;
    if (FAST) dist = CR::sd_rounded_box(clip_local_pos, this->v_clip_params.sel(X, Y), this->v_clip_params.z);
    else dist = distance_to_rounded_rect(clip_local_pos);
    Float alpha = CR::distance_aa(aa_range, dist);
    Float final_alpha = mix(alpha, 1.0f - alpha, Float(this->vClipMode.x));
    vec4 output_color = vec4(final_alpha);
    if (this->v_flags.z != 0) output_color = output_color.sel(X, X, X, X);
    this->gl_FragColor = output_color;
  }
  WR_FRAGMENT_ABI()
  ps_quad_mask_frag_t() { this->init_fragment_abi(); }
};
typedef ps_quad_mask_frag_t<false> ps_quad_mask_frag;
typedef ps_quad_mask_frag_t<true> ps_quad_mask_FAST_PATH_frag;
WR_PROGRAM(ps_quad_mask, "ps_quad_mask")
WR_PROGRAM(ps_quad_mask_FAST_PATH, "ps_quad_mask FAST_PATH")
