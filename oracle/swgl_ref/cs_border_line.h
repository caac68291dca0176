// TEST INFRASTRUCTURE — hand-instantiated SWGL programs for the border and line
// decoration render tasks drawn by draw_texture_cache_target
// (renderer/mod.rs:4015-4083):
//   "cs_line_decoration" (webrender/res/cs_line_decoration.glsl)
//   "cs_border_solid"    (cs_border_solid.glsl)
//   "cs_border_segment"  (cs_border_segment.glsl)
// Fragment shaders only (no swgl_drawSpan).  Flat varyings stay scalar, as the
// translator keeps them; `if` on a varying condition becomes a per-lane select.
#pragma once

// shared.glsl:110-113, 145-148 (SWGL), 184-189; ellipse.glsl:7-46
struct WrBorderMath {
  static float compute_aa_range(vec2 position) { return recip(fwidth(position).x); }
  static Float distance_aa(float aa_range, Float signed_distance) {
    Float dist = signed_distance * aa_range;
    return clamp(0.5f - dist, Float(0.0f), Float(1.0f));
  }
  // flat point and direction, varying sample position
  static Float distance_to_line(vec2_scalar p0, vec2_scalar perp_dir, vec2 p) {
    vec2 dir_to_p0 = vec2(p0) - p;
    return dot(vec2(normalize(perp_dir)), dir_to_p0);
  }
  // everything varying (cs_line_decoration's wave)
  static Float distance_to_line_v(vec2 p0, vec2 perp_dir, vec2 p) {
    vec2 dir_to_p0 = p0 - p;
    return dot(normalize(perp_dir), dir_to_p0);
  }
  static vec2_scalar inverse_radii_squared(vec2_scalar radii) { return 1.0f / max(radii * radii, 1.0e-6f); }
  static Float distance_to_ellipse_approx(vec2 p, vec2_scalar inv_radii_sq, float scale) {
    vec2 p_r = p * vec2(inv_radii_sq);
    Float g = dot(p, p_r) - scale;
    vec2 dG = (1.0f + scale) * p_r;
    return g * inversesqrt(dot(dG, dG));
  }
  static Float distance_to_ellipse(vec2 p, vec2_scalar radii) {
    return distance_to_ellipse_approx(p, inverse_radii_squared(radii), float(radii.x > 0.0f && radii.y > 0.0f));
  }
};

#define WR_CS_POS_FRAG_COMMON(NAME, VARY)                                                          \
  typedef NAME##_frag Self;                                                                        \
  typedef NAME##_vert::InterpOutputs InterpInputs;                                                 \
  typedef NAME##_vert::InterpOutputs InterpOutputs;                                                \
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
  }

// ---------------------------------------------------------------------------
struct cs_line_decoration_vert : VertexShaderImpl, WrCommon {
  typedef cs_line_decoration_vert Self;
  vec2 aPosition;
  vec4_scalar aTaskRect;
  vec2_scalar aLocalSize;
  float aWavyLineThickness, aAxisSelect;
  int aStyle;
  int a_loc[6];
  vec2 vLocalPos;
  ivec2_scalar vStyle;
  vec4_scalar vParams;
  struct InterpOutputs {
    vec2_scalar vLocalPos;
  };
  cs_line_decoration_vert() {
    static const char* names[6] = {"aPosition", "aTaskRect", "aLocalSize", "aWavyLineThickness", "aStyle",
                                   "aAxisSelect"};
    for (int i = 0; i < 6; i++) a_loc[i] = attrib_locations.add(names[i]);
    init_vertex_abi();
  }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    Self* self = (Self*)impl;
    auto& L = self->attrib_locations.locs;
    load_attrib(self->aPosition, attribs[L[self->a_loc[0]]], start, instance, count);
    load_flat_attrib(self->aTaskRect, attribs[L[self->a_loc[1]]], start, instance, count);
    load_flat_attrib(self->aLocalSize, attribs[L[self->a_loc[2]]], start, instance, count);
    load_flat_attrib(self->aWavyLineThickness, attribs[L[self->a_loc[3]]], start, instance, count);
    load_flat_attrib(self->aStyle, attribs[L[self->a_loc[4]]], start, instance, count);
    load_flat_attrib(self->aAxisSelect, attribs[L[self->a_loc[5]]], start, instance, count);
  }
  // cs_line_decoration.glsl:46-93
  void main() {
    vec2_scalar size = mix(aLocalSize, aLocalSize.sel(Y, X), aAxisSelect);
    vStyle.x = aStyle;
    switch (vStyle.x) {
      case 0:
        break;
      case 2:
        vParams = vec4_scalar(size.x, 0.5f * size.x, 0.0f, 0.0f);
        break;
      case 1: {
        float diameter = size.y;
        float period = diameter * 2.0f;
        float center_line = 0.5f * size.y;
        vParams = vec4_scalar(period, diameter / 2.0f, center_line, 0.0f);
        break;
      }
      case 3: {
        float line_thickness = max(aWavyLineThickness, 1.0f);
        float slope_length = size.y - line_thickness;
        float flat_length = max((line_thickness - 1.0f) * 2.0f, 1.0f);
        vParams = vec4_scalar(line_thickness / 2.0f, slope_length, flat_length, size.y);
        break;
      }
      default:
        vParams = vec4_scalar(0.0f);
    }
    vLocalPos = mix(aPosition, aPosition.sel(Y, X), Float(aAxisSelect)) * vec2(size);
    vec2 pos = mix(aTaskRect.sel(X, Y), aTaskRect.sel(Z, W), aPosition);
    gl_Position = uTransform * vec4(pos, Float(0.0f), Float(1.0f));
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vLocalPos = get_nth(vLocalPos, n);
      dest_ptr += stride;
    }
  }
  WR_VERTEX_ABI(cs_line_decoration)
};

struct cs_line_decoration_frag : FragmentShaderImpl, cs_line_decoration_vert, WrBorderMath {
  WR_CS_POS_FRAG_COMMON(cs_line_decoration, vLocalPos)
  // cs_line_decoration.glsl:100-162
  void main() {
    vec2 pos = vLocalPos;
    float aa_range = compute_aa_range(pos);
    Float alpha = 1.0f;
    switch (vStyle.x) {
      case 0:
        break;
      case 2:
        alpha = step(floor(pos.x + 0.5f), Float(vParams.y));
        break;
      case 1: {
        vec2 dot_relative_pos = pos - vec2(vParams.sel(Y, Z));
        Float dot_distance = length(dot_relative_pos) - vParams.y;
        alpha = distance_aa(aa_range, dot_distance);
        break;
      }
      case 3: {
        float half_line_thickness = vParams.x;
        float slope_length = vParams.y;
        float flat_length = vParams.z;
        float vertical_bounds = vParams.w;
        float half_period = slope_length + flat_length;
        float mid_height = vertical_bounds / 2.0f;
        Float peak_offset = mid_height - half_line_thickness;
        Float flip = -2.0f * (step(mod(pos.x, Float(2.0f * half_period)), Float(half_period)) - 0.5f);
        peak_offset *= flip;
        Float peak_height = mid_height + peak_offset;
        pos.x = mod(pos.x, Float(half_period));
        Float dist1 = distance_to_line_v(vec2(Float(0.0f), peak_height), vec2(Float(1.0f), -flip), pos);
        Float dist2 = distance_to_line_v(vec2(Float(0.0f), peak_height), vec2(Float(0.0f), -flip), pos);
        Float dist3 = distance_to_line_v(vec2(Float(flat_length), peak_height), vec2(Float(-1.0f), -flip), pos);
        Float dist = abs(max(max(dist1, dist2), dist3));
        alpha = distance_aa(aa_range, dist - half_line_thickness);
        if (half_line_thickness <= 1.0f) {
          alpha = 1.0f - step(alpha, Float(0.5f));
        }
        break;
      }
      default:
        break;
    }
    gl_FragColor = vec4(alpha);
  }
  WR_FRAGMENT_ABI()
  cs_line_decoration_frag() { init_fragment_abi(); }
};
WR_PROGRAM(cs_line_decoration, "cs_line_decoration")

// ---------------------------------------------------------------------------
// BorderInstance attributes (renderer/vertex.rs desc::BORDER, gpu_types.rs:193-202)
struct BorderVertBase : VertexShaderImpl, WrCommon {
  vec2 aPosition;
  vec2_scalar aTaskOrigin, aWidths, aRadii;
  vec4_scalar aRect, aColor0, aColor1, aClipParams1, aClipParams2;
  int aFlags;
  int a_loc[10];
  vec2 vPos;
  struct InterpOutputs {
    vec2_scalar vPos;
  };
  BorderVertBase() {
    static const char* names[10] = {"aPosition", "aTaskOrigin", "aRect", "aColor0", "aColor1", "aFlags",
                                    "aWidths", "aRadii", "aClipParams1", "aClipParams2"};
    for (int i = 0; i < 10; i++) a_loc[i] = attrib_locations.add(names[i]);
  }
  void load_border_attribs(VertexAttrib* attribs, uint32_t start, int instance, int count) {
    auto& L = attrib_locations.locs;
    load_attrib(aPosition, attribs[L[a_loc[0]]], start, instance, count);
    load_flat_attrib(aTaskOrigin, attribs[L[a_loc[1]]], start, instance, count);
    load_flat_attrib(aRect, attribs[L[a_loc[2]]], start, instance, count);
    load_flat_attrib(aColor0, attribs[L[a_loc[3]]], start, instance, count);
    load_flat_attrib(aColor1, attribs[L[a_loc[4]]], start, instance, count);
    load_flat_attrib(aFlags, attribs[L[a_loc[5]]], start, instance, count);
    load_flat_attrib(aWidths, attribs[L[a_loc[6]]], start, instance, count);
    load_flat_attrib(aRadii, attribs[L[a_loc[7]]], start, instance, count);
    load_flat_attrib(aClipParams1, attribs[L[a_loc[8]]], start, instance, count);
    load_flat_attrib(aClipParams2, attribs[L[a_loc[9]]], start, instance, count);
  }
  static vec2_scalar get_outer_corner_scale(int segment) {
    switch (segment) {
      case 0: return vec2_scalar(0.0f, 0.0f);
      case 1: return vec2_scalar(1.0f, 0.0f);
      case 2: return vec2_scalar(1.0f, 1.0f);
      case 3: return vec2_scalar(0.0f, 1.0f);
      default: return vec2_scalar(0.0f);
    }
  }
  ALWAYS_INLINE void store_interp_outputs(char* dest_ptr, size_t stride) {
    for (int n = 0; n < 4; n++) {
      auto* dest = reinterpret_cast<InterpOutputs*>(dest_ptr);
      dest->vPos = get_nth(vPos, n);
      dest_ptr += stride;
    }
  }
};

struct cs_border_solid_vert : BorderVertBase {
  typedef cs_border_solid_vert Self;
  vec4_scalar vColor0, vColor1, vColorLine, vClipCenter_Sign, vClipRadii;
  vec4_scalar vHorizontalClipCenter_Sign, vVerticalClipCenter_Sign;
  vec2_scalar vHorizontalClipRadii, vVerticalClipRadii;
  ivec2_scalar vMixColors;
  cs_border_solid_vert() { init_vertex_abi(); }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    ((Self*)impl)->load_border_attribs(attribs, start, instance, count);
  }
  // cs_border_solid.glsl:85-133
  void main() {
    int segment = aFlags & 0xff;
    bool do_aa = ((aFlags >> 24) & 0xf0) != 0;
    vec2_scalar outer_scale = get_outer_corner_scale(segment);
    vec2_scalar size = aRect.sel(Z, W) - aRect.sel(X, Y);
    vec2_scalar outer = outer_scale * size;
    vec2_scalar clip_sign = 1.0f - 2.0f * outer_scale;
    int mix_colors;
    switch (segment) {
      case 0: case 1: case 2: case 3:
        mix_colors = do_aa ? 1 : 2;
        break;
      default:
        mix_colors = 0;
        break;
    }
    vMixColors.x = mix_colors;
    vPos = vec2(size) * aPosition;
    vColor0 = aColor0;
    vColor1 = aColor1;
    vClipCenter_Sign = make_vec4(outer + clip_sign * aRadii, clip_sign);
    vClipRadii = make_vec4(aRadii, max(aRadii - aWidths, 0.0f));
    vColorLine = make_vec4(outer, vec2_scalar(aWidths.y * -clip_sign.y, aWidths.x * clip_sign.x));
    vec2_scalar horizontal_clip_sign = vec2_scalar(-clip_sign.x, clip_sign.y);
    vHorizontalClipCenter_Sign =
        make_vec4(aClipParams1.sel(X, Y) + horizontal_clip_sign * aClipParams1.sel(Z, W), horizontal_clip_sign);
    vHorizontalClipRadii = aClipParams1.sel(Z, W);
    vec2_scalar vertical_clip_sign = vec2_scalar(clip_sign.x, -clip_sign.y);
    vVerticalClipCenter_Sign =
        make_vec4(aClipParams2.sel(X, Y) + vertical_clip_sign * aClipParams2.sel(Z, W), vertical_clip_sign);
    vVerticalClipRadii = aClipParams2.sel(Z, W);
    gl_Position = uTransform * vec4(vec2(aTaskOrigin + aRect.sel(X, Y)) + vPos, Float(0.0f), Float(1.0f));
  }
  WR_VERTEX_ABI(cs_border_solid)
};

struct cs_border_solid_frag : FragmentShaderImpl, cs_border_solid_vert, WrBorderMath {
  WR_CS_POS_FRAG_COMMON(cs_border_solid, vPos)
  // cs_border_solid.glsl:137-177
  void main() {
    float aa_range = compute_aa_range(vPos);
    bool do_aa = vMixColors.x != 2;
    Float mix_factor = 0.0f;
    if (vMixColors.x != 0) {
      Float d_line = distance_to_line(vColorLine.sel(X, Y), vColorLine.sel(Z, W), vPos);
      if (do_aa) {
        mix_factor = distance_aa(aa_range, -d_line);
      } else {
        mix_factor = if_then_else(d_line + 0.0001f >= 0.0f, Float(1.0f), Float(0.0f));
      }
    }
    vec2 clip_relative_pos = vPos - vec2(vClipCenter_Sign.sel(X, Y));
    auto in_clip_region = (vClipCenter_Sign.z * clip_relative_pos.x < 0.0f) &
                          (vClipCenter_Sign.w * clip_relative_pos.y < 0.0f);
    Float d = -1.0f;
    {
      Float d_radii_a = distance_to_ellipse(clip_relative_pos, vClipRadii.sel(X, Y));
      Float d_radii_b = distance_to_ellipse(clip_relative_pos, vClipRadii.sel(Z, W));
      d = if_then_else(in_clip_region, max(d_radii_a, -d_radii_b), d);
    }
    clip_relative_pos = vPos - vec2(vHorizontalClipCenter_Sign.sel(X, Y));
    in_clip_region = (vHorizontalClipCenter_Sign.z * clip_relative_pos.x < 0.0f) &
                     (vHorizontalClipCenter_Sign.w * clip_relative_pos.y < 0.0f);
    {
      Float d_radii = distance_to_ellipse(clip_relative_pos, vHorizontalClipRadii);
      d = if_then_else(in_clip_region, max(d_radii, d), d);
    }
    clip_relative_pos = vPos - vec2(vVerticalClipCenter_Sign.sel(X, Y));
    in_clip_region = (vVerticalClipCenter_Sign.z * clip_relative_pos.x < 0.0f) &
                     (vVerticalClipCenter_Sign.w * clip_relative_pos.y < 0.0f);
    {
      Float d_radii = distance_to_ellipse(clip_relative_pos, vVerticalClipRadii);
      d = if_then_else(in_clip_region, max(d_radii, d), d);
    }
    Float alpha = do_aa ? distance_aa(aa_range, d) : Float(1.0f);
    vec4 color = mix(vec4(vColor0), vec4(vColor1), mix_factor);
    gl_FragColor = color * alpha;
  }
  WR_FRAGMENT_ABI()
  cs_border_solid_frag() { init_fragment_abi(); }
};
WR_PROGRAM(cs_border_solid, "cs_border_solid")

// ---------------------------------------------------------------------------
struct cs_border_segment_vert : BorderVertBase {
  typedef cs_border_segment_vert Self;
  vec4_scalar vColor00, vColor01, vColor10, vColor11, vColorLine, vStyleEdgeAxis, vClipCenter_Sign, vClipRadii;
  vec4_scalar vEdgeReference, vPartialWidths, vClipParams1, vClipParams2;
  vec2_scalar vSegmentClipMode;
  cs_border_segment_vert() { init_vertex_abi(); }
  static void load_attribs(VertexShaderImpl* impl, VertexAttrib* attribs, uint32_t start, int instance,
                           int count) {
    ((Self*)impl)->load_border_attribs(attribs, start, instance, count);
  }
  // cs_border_segment.glsl:113-133
  static vec4_scalar mod_color(vec4_scalar color, bool is_black, bool lighter) {
    const float light_black = 0.7f, dark_black = 0.3f, dark_scale = 0.66666666f, light_scale = 1.0f;
    if (is_black) {
      if (lighter) return vec4_scalar(light_black, light_black, light_black, color.w);
      return vec4_scalar(dark_black, dark_black, dark_black, color.w);
    }
    if (lighter) return vec4_scalar(color.x * light_scale, color.y * light_scale, color.z * light_scale, color.w);
    return vec4_scalar(color.x * dark_scale, color.y * dark_scale, color.z * dark_scale, color.w);
  }
  // cs_border_segment.glsl:135-157
  static void get_colors_for_side(vec4_scalar color, int style, vec4_scalar* result) {
    bool is_black = color.x == 0.0f && color.y == 0.0f && color.z == 0.0f;
    switch (style) {
      case 6:
        result[0] = mod_color(color, is_black, true);
        result[1] = mod_color(color, is_black, false);
        break;
      case 7:
        result[0] = mod_color(color, is_black, false);
        result[1] = mod_color(color, is_black, true);
        break;
      default:
        result[0] = color;
        result[1] = color;
        break;
    }
  }
  // cs_border_segment.glsl:159-254
  void main() {
    int segment = aFlags & 0xff;
    int style0 = (aFlags >> 8) & 0xff;
    int style1 = (aFlags >> 16) & 0xff;
    int clip_mode = (aFlags >> 24) & 0x0f;
    vec2_scalar size = aRect.sel(Z, W) - aRect.sel(X, Y);
    vec2_scalar outer_scale = get_outer_corner_scale(segment);
    vec2_scalar outer = outer_scale * size;
    vec2_scalar clip_sign = 1.0f - 2.0f * outer_scale;
    ivec2_scalar edge_axis = ivec2_scalar(0, 0);
    vec2_scalar edge_reference = vec2_scalar(0.0f);
    switch (segment) {
      case 0:
        edge_axis = ivec2_scalar(0, 1);
        edge_reference = outer;
        break;
      case 1:
        edge_axis = ivec2_scalar(1, 0);
        edge_reference = vec2_scalar(outer.x - aWidths.x, outer.y);
        break;
      case 2:
        edge_axis = ivec2_scalar(0, 1);
        edge_reference = outer - aWidths;
        break;
      case 3:
        edge_axis = ivec2_scalar(1, 0);
        edge_reference = vec2_scalar(outer.x, outer.y - aWidths.y);
        break;
      case 5: case 7:
        edge_axis = ivec2_scalar(1, 1);
        break;
      default:
        break;
    }
    vSegmentClipMode = vec2_scalar(float(segment), float(clip_mode));
    vStyleEdgeAxis = vec4_scalar(float(style0), float(style1), float(edge_axis.x), float(edge_axis.y));
    vPartialWidths = make_vec4(aWidths / 3.0f, aWidths / 2.0f);
    vPos = vec2(size) * aPosition;
    vec4_scalar c[2];
    get_colors_for_side(aColor0, style0, c);
    vColor00 = c[0];
    vColor01 = c[1];
    get_colors_for_side(aColor1, style1, c);
    vColor10 = c[0];
    vColor11 = c[1];
    vClipCenter_Sign = make_vec4(outer + clip_sign * aRadii, clip_sign);
    vClipRadii = make_vec4(aRadii, max(aRadii - aWidths, 0.0f));
    vColorLine = make_vec4(outer, vec2_scalar(aWidths.y * -clip_sign.y, aWidths.x * clip_sign.x));
    vEdgeReference = make_vec4(edge_reference, edge_reference + aWidths);
    vClipParams1 = aClipParams1;
    vClipParams2 = aClipParams2;
    if (clip_mode == 3) {
      float radius = aClipParams1.z;
      if (radius > 0.5f) radius += 2.0f;
      vPos = vec2(vClipParams1.sel(X, Y)) + radius * (2.0f * aPosition - 1.0f);
      vPos = clamp(vPos, vec2(vec2_scalar(0.0f)), vec2(size));
    } else if (clip_mode == 1) {
      vec2_scalar center = (aClipParams1.sel(X, Y) + aClipParams2.sel(X, Y)) * 0.5f;
      float dash_length = length(aClipParams1.sel(X, Y) - aClipParams2.sel(X, Y));
      float width = max(aWidths.x, aWidths.y);
      vec2_scalar r = vec2_scalar(max(dash_length, width)) + 2.0f;
      vPos = clamp(vPos, vec2(center - r), vec2(center + r));
    }
    gl_Position = uTransform * vec4(vec2(aTaskOrigin + aRect.sel(X, Y)) + vPos, Float(0.0f), Float(1.0f));
  }
  WR_VERTEX_ABI(cs_border_segment)
};

struct cs_border_segment_frag : FragmentShaderImpl, cs_border_segment_vert, WrBorderMath {
  WR_CS_POS_FRAG_COMMON(cs_border_segment, vPos)
  // cs_border_segment.glsl:258-313
  vec4 evaluate_color_for_style_in_corner(vec2 clip_relative_pos, int style, vec4_scalar color0s, vec4_scalar color1s,
                                          vec4_scalar clip_radii, Float mix_factor, int segment, float aa_range) {
    vec4 color0 = vec4(color0s), color1 = vec4(color1s);
    switch (style) {
      case 2: {
        Float d_radii_a = distance_to_ellipse(clip_relative_pos, clip_radii.sel(X, Y) - vPartialWidths.sel(X, Y));
        Float d_radii_b =
            distance_to_ellipse(clip_relative_pos, clip_radii.sel(X, Y) - 2.0f * vPartialWidths.sel(X, Y));
        Float d = min(-d_radii_a, d_radii_b);
        color0 *= distance_aa(aa_range, d);
        break;
      }
      case 6:
      case 7: {
        Float d = distance_to_ellipse(clip_relative_pos, clip_radii.sel(X, Y) - vPartialWidths.sel(Z, W));
        Float alpha = distance_aa(aa_range, d);
        Float swizzled_factor;
        switch (segment) {
          case 0: swizzled_factor = 0.0f; break;
          case 1: swizzled_factor = mix_factor; break;
          case 2: swizzled_factor = 1.0f; break;
          case 3: swizzled_factor = 1.0f - mix_factor; break;
          default: swizzled_factor = 0.0f; break;
        }
        vec4 c0 = mix(color1, color0, swizzled_factor);
        vec4 c1 = mix(color0, color1, swizzled_factor);
        color0 = mix(c0, c1, alpha);
        break;
      }
      default:
        break;
    }
    return color0;
  }
  // cs_border_segment.glsl:315-355
  vec4 evaluate_color_for_style_in_edge(vec2 pos_vec, int style, vec4_scalar color0s, vec4_scalar color1s,
                                        float aa_range, int edge_axis_id) {
    vec4 color0 = vec4(color0s), color1 = vec4(color1s);
    vec2_scalar edge_axis = edge_axis_id != 0 ? vec2_scalar(0.0f, 1.0f) : vec2_scalar(1.0f, 0.0f);
    Float pos = dot(pos_vec, vec2(edge_axis));
    switch (style) {
      case 2: {
        Float d = -1.0f;
        float partial_width = dot(vPartialWidths.sel(X, Y), edge_axis);
        if (partial_width >= 1.0f) {
          vec2_scalar ref = vec2_scalar(dot(vEdgeReference.sel(X, Y), edge_axis) + partial_width,
                                        dot(vEdgeReference.sel(Z, W), edge_axis) - partial_width);
          d = min(pos - ref.x, ref.y - pos);
        }
        color0 *= distance_aa(aa_range, d);
        break;
      }
      case 6:
      case 7: {
        float ref = dot(vEdgeReference.sel(X, Y) + vPartialWidths.sel(Z, W), edge_axis);
        Float d = pos - ref;
        Float alpha = distance_aa(aa_range, d);
        color0 = mix(color0, color1, alpha);
        break;
      }
      default:
        break;
    }
    return color0;
  }
  // cs_border_segment.glsl:357-449
  void main() {
    float aa_range = compute_aa_range(vPos);
    int segment = int(vSegmentClipMode.x);
    int clip_mode = int(vSegmentClipMode.y);
    ivec2_scalar style = ivec2_scalar(int(vStyleEdgeAxis.x), int(vStyleEdgeAxis.y));
    ivec2_scalar edge_axis = ivec2_scalar(int(vStyleEdgeAxis.z), int(vStyleEdgeAxis.w));
    Float mix_factor = 0.0f;
    if (edge_axis.x != edge_axis.y) {
      Float d_line = distance_to_line(vColorLine.sel(X, Y), vColorLine.sel(Z, W), vPos);
      mix_factor = distance_aa(aa_range, -d_line);
    }
    vec2 clip_relative_pos = vPos - vec2(vClipCenter_Sign.sel(X, Y));
    auto in_clip_region = (vClipCenter_Sign.z * clip_relative_pos.x < 0.0f) &
                          (vClipCenter_Sign.w * clip_relative_pos.y < 0.0f);
    Float d = -1.0f;
    switch (clip_mode) {
      case 3:
        d = length(vec2(vClipParams1.sel(X, Y)) - vPos) - vClipParams1.z;
        break;
      case 2: {
        bool is_vertical = vClipParams1.x == 0.0f;
        float half_dash = is_vertical ? vClipParams1.y : vClipParams1.x;
        Float pos = is_vertical ? vPos.y : vPos.x;
        auto in_dash = (pos < half_dash) | (pos > 3.0f * half_dash);
        d = if_then_else(in_dash, d, Float(1.0f));
        break;
      }
      case 1: {
        Float d0 = distance_to_line(vClipParams1.sel(X, Y), vClipParams1.sel(Z, W), vPos);
        Float d1 = distance_to_line(vClipParams2.sel(X, Y), vClipParams2.sel(Z, W), vPos);
        d = max(d0, -d1);
        break;
      }
      default:
        break;
    }
    Float d_radii_a = distance_to_ellipse(clip_relative_pos, vClipRadii.sel(X, Y));
    Float d_radii_b = distance_to_ellipse(clip_relative_pos, vClipRadii.sel(Z, W));
    Float d_radii = max(d_radii_a, -d_radii_b);
    d = if_then_else(in_clip_region, max(d, d_radii), d);
    vec4 c0c = evaluate_color_for_style_in_corner(clip_relative_pos, style.x, vColor00, vColor01, vClipRadii,
                                                  mix_factor, segment, aa_range);
    vec4 c1c = evaluate_color_for_style_in_corner(clip_relative_pos, style.y, vColor10, vColor11, vClipRadii,
                                                  mix_factor, segment, aa_range);
    vec4 c0e = evaluate_color_for_style_in_edge(vPos, style.x, vColor00, vColor01, aa_range, edge_axis.x);
    vec4 c1e = evaluate_color_for_style_in_edge(vPos, style.y, vColor10, vColor11, aa_range, edge_axis.y);
    vec4 color0 = if_then_else(in_clip_region, c0c, c0e);
    vec4 color1 = if_then_else(in_clip_region, c1c, c1e);
    Float alpha = distance_aa(aa_range, d);
    vec4 color = mix(color0, color1, mix_factor);
    gl_FragColor = color * alpha;
  }
  WR_FRAGMENT_ABI()
  cs_border_segment_frag() { init_fragment_abi(); }
};
WR_PROGRAM(cs_border_segment, "cs_border_segment")
