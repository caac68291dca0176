// TEST INFRASTRUCTURE — stands in for the `load_shader.h` that swgl/build.rs:19-31
// generates: maps a program key ("name feat,feat") to its loader.
#pragma once
#include "wr_common.h"
#include "ps_quad.h"
#include "ps_quad_textured.h"
#include "brush.h"
#include "brush_solid.h"
#include "cs_clip_rectangle.h"
#include "ps_quad_mask.h"
#include "brush_image.h"
#include "brush_image_repeat.h"
#include "ps_text_run.h"
#include "ps_text_run_gt.h"
#include "brush_linear_gradient.h"
#include "cs_clip_box_shadow.h"
#include "composite.h"
#include "composite_yuv.h"
#include "brush_yuv_image.h"
#include "brush_opacity.h"
#include "ps_clear.h"
#include "brush_blend.h"
#include "brush_mix_blend.h"
#include "cs_blur.h"
#include "cs_scale.h"
#include "cs_gradients.h"
#include "cs_border_line.h"
#include "ps_quad_gradients.h"
#include "ps_split_composite.h"

ProgramLoader load_shader(const char* name) {
  if (!strcmp(name, "ps_quad_textured")) return ps_quad_textured_program::loader;
  if (!strcmp(name, "brush_solid")) return brush_solid_program::loader;
  if (!strcmp(name, "brush_solid ALPHA_PASS")) return brush_solid_ALPHA_PASS_program::loader;
  if (!strcmp(name, "cs_clip_rectangle")) return cs_clip_rectangle_program::loader;
  if (!strcmp(name, "cs_clip_rectangle FAST_PATH")) return cs_clip_rectangle_FAST_PATH_program::loader;
  if (!strcmp(name, "ps_quad_mask")) return ps_quad_mask_program::loader;
  if (!strcmp(name, "ps_quad_mask FAST_PATH")) return ps_quad_mask_FAST_PATH_program::loader;
  if (!strcmp(name, "brush_image TEXTURE_2D")) return brush_image_TEXTURE_2D_program::loader;
  if (!strcmp(name, "brush_image ALPHA_PASS,TEXTURE_2D")) return brush_image_ALPHA_PASS_TEXTURE_2D_program::loader;
  if (!strcmp(name, "brush_image ADVANCED_BLEND,ALPHA_PASS,TEXTURE_2D"))
    return brush_image_ADVANCED_BLEND_ALPHA_PASS_TEXTURE_2D_program::loader;
  if (!strcmp(name, "ps_text_run ALPHA_PASS,TEXTURE_2D")) return ps_text_run_ALPHA_PASS_TEXTURE_2D_program::loader;
  if (!strcmp(name, "ps_text_run ALPHA_PASS,DUAL_SOURCE_BLENDING,TEXTURE_2D"))
    return ps_text_run_ALPHA_PASS_DUAL_SOURCE_BLENDING_TEXTURE_2D_program::loader;
  if (!strcmp(name, "brush_linear_gradient")) return brush_linear_gradient_program::loader;
  if (!strcmp(name, "brush_linear_gradient ALPHA_PASS")) return brush_linear_gradient_ALPHA_PASS_program::loader;
  if (!strcmp(name, "cs_clip_box_shadow TEXTURE_2D")) return cs_clip_box_shadow_TEXTURE_2D_program::loader;
  if (!strcmp(name, "composite TEXTURE_2D")) return composite_TEXTURE_2D_program::loader;
  if (!strcmp(name, "composite FAST_PATH,TEXTURE_2D")) return composite_FAST_PATH_TEXTURE_2D_program::loader;
  if (!strcmp(name, "composite TEXTURE_2D,YUV")) return composite_TEXTURE_2D_YUV_program::loader;
  if (!strcmp(name, "brush_yuv_image TEXTURE_2D,YUV")) return brush_yuv_image_TEXTURE_2D_YUV_program::loader;
  if (!strcmp(name, "brush_yuv_image ALPHA_PASS,TEXTURE_2D,YUV")) return brush_yuv_image_ALPHA_PASS_TEXTURE_2D_YUV_program::loader;
  if (!strcmp(name, "brush_yuv_image ALPHA_PASS,ANTIALIASING,TEXTURE_2D,YUV")) return brush_yuv_image_ALPHA_PASS_ANTIALIASING_TEXTURE_2D_YUV_program::loader;
  if (!strcmp(name, "brush_opacity")) return brush_opacity_program::loader;
  if (!strcmp(name, "brush_opacity ALPHA_PASS")) return brush_opacity_ALPHA_PASS_program::loader;
  if (!strcmp(name, "brush_opacity ALPHA_PASS,ANTIALIASING")) return brush_opacity_ALPHA_PASS_ANTIALIASING_program::loader;
  if (!strcmp(name, "ps_clear")) return ps_clear_program::loader;
  if (!strcmp(name, "brush_blend")) return brush_blend_program::loader;
  if (!strcmp(name, "brush_blend ALPHA_PASS")) return brush_blend_ALPHA_PASS_program::loader;
  if (!strcmp(name, "brush_mix_blend")) return brush_mix_blend_program::loader;
  if (!strcmp(name, "brush_mix_blend ALPHA_PASS")) return brush_mix_blend_ALPHA_PASS_program::loader;
  if (!strcmp(name, "cs_blur ALPHA_TARGET")) return cs_blur_ALPHA_TARGET_program::loader;
  if (!strcmp(name, "cs_blur COLOR_TARGET")) return cs_blur_COLOR_TARGET_program::loader;
  if (!strcmp(name, "cs_scale TEXTURE_2D")) return cs_scale_TEXTURE_2D_program::loader;
  if (!strcmp(name, "cs_fast_linear_gradient")) return cs_fast_linear_gradient_program::loader;
  if (!strcmp(name, "cs_linear_gradient")) return cs_linear_gradient_program::loader;
  if (!strcmp(name, "cs_radial_gradient")) return cs_radial_gradient_program::loader;
  if (!strcmp(name, "cs_conic_gradient")) return cs_conic_gradient_program::loader;
  if (!strcmp(name, "cs_line_decoration")) return cs_line_decoration_program::loader;
  if (!strcmp(name, "cs_border_solid")) return cs_border_solid_program::loader;
  if (!strcmp(name, "cs_border_segment")) return cs_border_segment_program::loader;
  if (!strcmp(name, "ps_quad_radial_gradient")) return ps_quad_radial_gradient_program::loader;
  if (!strcmp(name, "ps_quad_conic_gradient")) return ps_quad_conic_gradient_program::loader;
  if (!strcmp(name, "brush_image ANTIALIASING,REPETITION,TEXTURE_2D"))
    return brush_image_ANTIALIASING_REPETITION_TEXTURE_2D_program::loader;
  if (!strcmp(name, "brush_image ALPHA_PASS,ANTIALIASING,REPETITION,TEXTURE_2D"))
    return brush_image_ALPHA_PASS_ANTIALIASING_REPETITION_TEXTURE_2D_program::loader;
  if (!strcmp(name, "ps_text_run ALPHA_PASS,GLYPH_TRANSFORM,TEXTURE_2D"))
    return ps_text_run_ALPHA_PASS_GLYPH_TRANSFORM_TEXTURE_2D_program::loader;
  if (!strcmp(name, "ps_text_run ALPHA_PASS,DUAL_SOURCE_BLENDING,GLYPH_TRANSFORM,TEXTURE_2D"))
    return ps_text_run_ALPHA_PASS_DUAL_SOURCE_BLENDING_GLYPH_TRANSFORM_TEXTURE_2D_program::loader;
  if (!strcmp(name, "ps_split_composite")) return ps_split_composite_program::loader;
  return nullptr;
}
