"""Binary layouts of the batch inputs the frame-draw path consumes.

Host-side encoders for the `#[repr(C)]` records of webrender/src/gpu_types.rs
and the per-frame data tables.  These are what the
reference's frame builder (CPU, out of scope) hands to `Renderer::draw_frame`;
tests and benchmarks use them to synthesise frames.
"""
import numpy as np

# QuadFlags, webrender/src/command_buffer.rs:76 / ps_quad.glsl:62-66
QF_IS_OPAQUE = 1
QF_APPLY_DEVICE_CLIP = 2
QF_IGNORE_DEVICE_SCALE = 4
QF_USE_AA_SEGMENTS = 8
QF_IS_MASK = 16

EDGE_AA_LEFT, EDGE_AA_TOP, EDGE_AA_RIGHT, EDGE_AA_BOTTOM = 1, 2, 4, 8
PART_CENTER, PART_LEFT, PART_TOP, PART_RIGHT, PART_BOTTOM, PART_ALL = range(6)
INVALID_SEGMENT_INDEX = 0xFF

CLIP_TASK_EMPTY = 0x7FFFFFFF  # render_task.glsl:75

IDENTITY = np.eye(4, dtype=np.float32)


def ortho(width, height, near=-float(1 << 22), far=float((1 << 22) - 1)):
    """Transform3D::ortho(0, w, 0, h, near, far) as the reference passes it to
    uTransform (renderer/mod.rs:4705-4712, device/gl.rs:2040-2046); returned
    column-major, 16 floats (euclid: m11..m44 row-vector convention == GL
    column-major upload)."""
    left, right, bottom, top = np.float32(0), np.float32(width), np.float32(0), np.float32(height)
    near, far = np.float32(near), np.float32(far)
    tx = -((right + left) / (right - left))
    ty = -((top + bottom) / (top - bottom))
    tz = -((far + near) / (far - near))
    m = np.zeros(16, dtype=np.float32)
    m[0] = np.float32(2) / (right - left)
    m[5] = np.float32(2) / (top - bottom)
    m[10] = np.float32(-2) / (far - near)
    m[12], m[13], m[14], m[15] = tx, ty, tz, 1
    return m


def scale_offset_transform(sx=1.0, sy=1.0, tx=0.0, ty=0.0):
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 1], m[0, 3], m[1, 3] = sx, sy, tx, ty
    return m


class FrameTables:
    """The per-frame data tables (`Frame.prim_headers`, `transform_palette`,
    `render_tasks`, gpu cache, `gpu_buffer_f/i`; frame_builder.rs:1129-1180).
    All tables are arrays of 16-byte texels."""

    def __init__(self):
        self.prim_headers_f = []   # 2 texels / prim
        self.prim_headers_i = []   # 2 texels / prim
        self.transforms = []       # 8 texels / transform
        self.render_tasks = []     # 2 texels / task
        self.gpu_cache = []
        self.gpu_buffer_f = []
        self.gpu_buffer_i = []
        self.add_transform(IDENTITY)  # TransformPaletteId::IDENTITY == 0

    # -- transform palette (gpu_types.rs:736-768, transform.glsl:22-46) -----
    def add_transform(self, m, inv=None, axis_aligned=True):
        """m: 4x4 matrix in math (row, col) convention; stored column-major."""
        m = np.asarray(m, dtype=np.float32).reshape(4, 4)
        if inv is None:
            inv = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
        idx = len(self.transforms) // 8
        for mat in (m, inv):
            for col in range(4):
                self.transforms.append(mat[:, col].astype(np.float32).copy())
        return idx | (0 if axis_aligned else (1 << 23))

    # -- render tasks (render_task.rs:723-836, render_task.glsl) ------------
    def add_render_task(self, rect, device_pixel_scale=1.0, content_origin=(0.0, 0.0)):
        addr = len(self.render_tasks) // 2
        self.render_tasks.append(np.array(rect, dtype=np.float32))
        self.render_tasks.append(np.array([device_pixel_scale, content_origin[0], content_origin[1], 0.0],
                                          dtype=np.float32))
        return addr

    def push_gpu_cache(self, blocks):
        addr = len(self.gpu_cache)
        for b in blocks:
            self.gpu_cache.append(np.asarray(b, dtype=np.float32).reshape(4))
        return addr

    def push_gpu_buffer_f(self, blocks):
        addr = len(self.gpu_buffer_f)
        for b in blocks:
            self.gpu_buffer_f.append(np.asarray(b, dtype=np.float32).reshape(4))
        return addr

    def push_gpu_buffer_i(self, block):
        addr = len(self.gpu_buffer_i)
        self.gpu_buffer_i.append(np.asarray(block, dtype=np.int32).reshape(4))
        return addr

    # -- prim headers (gpu_types.rs:476-493, prim_shared.glsl:77-96) --------
    def add_prim_header(self, local_rect, local_clip_rect, z, specific_prim_address, transform_id,
                        render_task_address, user_data=(0, 0, 0, 0)):
        idx = len(self.prim_headers_f) // 2
        self.prim_headers_f.append(np.array(local_rect, dtype=np.float32))
        self.prim_headers_f.append(np.array(local_clip_rect, dtype=np.float32))
        self.prim_headers_i.append(np.array([z, specific_prim_address, transform_id, render_task_address],
                                            dtype=np.int32))
        self.prim_headers_i.append(np.array(user_data, dtype=np.int64).astype(np.int32))
        return idx

    # -- quads (quad.rs:941-1001, ps_quad.glsl:96-145) -------------------------
    def add_quad_prim(self, bounds, clip, color, uv_rect=(0, 0, 0, 0), scale_offset=(1, 1, 0, 0),
                      segments=()):
        """write_prim_blocks: 5 blocks + 2 per segment (rect, uv rect)."""
        blocks = [bounds, clip, uv_rect, scale_offset, color]
        for rect, uv in segments:
            blocks += [rect, uv]
        return self.push_gpu_buffer_f(blocks)

    def add_quad_header(self, transform_id, z_id, pattern_input=(0, 0)):
        return self.push_gpu_buffer_i([transform_id, z_id, pattern_input[0], pattern_input[1]])

    def arrays(self):
        def f(lst):
            return (np.ascontiguousarray(np.stack(lst).astype(np.float32)) if lst
                    else np.zeros((0, 4), np.float32))

        def i(lst):
            return (np.ascontiguousarray(np.stack(lst).astype(np.int32)) if lst
                    else np.zeros((0, 4), np.int32))
        return {
            "prim_headers_f": f(self.prim_headers_f), "prim_headers_i": i(self.prim_headers_i),
            "transforms": f(self.transforms), "render_tasks": f(self.render_tasks),
            "gpu_cache": f(self.gpu_cache), "gpu_buffer_f": f(self.gpu_buffer_f),
            "gpu_buffer_i": i(self.gpu_buffer_i),
        }


def quad_instance(prim_address_i, prim_address_f, quad_flags, edge_flags, part_index, segment_index,
                  render_task_address):
    """QuadInstance → PrimitiveInstanceData (gpu_types.rs:554-587)."""
    z = ((quad_flags & 0xFF) << 24) | ((edge_flags & 0xFF) << 16) | ((part_index & 0xFF) << 8) | (segment_index & 0xFF)
    return np.array([prim_address_i, prim_address_f, z, render_task_address], dtype=np.int64).astype(np.int32)


def brush_instance(prim_header_index, clip_task_address, segment_index, edge_flags, brush_flags,
                   resource_address):
    """BrushInstance → PrimitiveInstanceData (gpu_types.rs:681-703)."""
    z = (segment_index & 0xFFFF) | ((brush_flags & 0xFFF) << 16) | ((edge_flags & 0xF) << 28)
    return np.array([prim_header_index, clip_task_address, z, resource_address],
                    dtype=np.int64).astype(np.int32)


def split_composite_instance(prim_header_index, polygons_address, z, render_task_address):
    """SplitCompositeInstance → PrimitiveInstanceData (gpu_types.rs:531-552)."""
    return np.array([prim_header_index, polygons_address, z, render_task_address], dtype=np.int64).astype(np.int32)


def clip_rect_instance(sub_rect, task_origin, screen_origin, device_pixel_scale, clip_transform_id,
                       prim_transform_id, local_pos, local_rect, mode, radii):
    """ClipMaskInstanceRect, 200 bytes (gpu_types.rs:208-225, prim_store/mod.rs:774-813).
    radii = ((tl_rx, tl_ry), (tr_rx, tr_ry), (bl_rx, bl_ry), (br_rx, br_ry))."""
    x0, y0, x1, y1 = local_rect
    (tl, tr, bl, br) = radii
    corner_rects = [
        (x0, y0, x0 + tl[0], y0 + tl[1]),
        (x1 - tr[0], y0, x1, y0 + tr[1]),
        (x0, y1 - bl[1], x0 + bl[0], y1),
        (x1 - br[0], y1 - br[1], x1, y1),
    ]
    buf = np.zeros(50, dtype=np.float32)
    buf[0:4] = sub_rect
    buf[4:6] = task_origin
    buf[6:8] = screen_origin
    buf[8] = device_pixel_scale
    buf[9:11] = np.array([clip_transform_id, prim_transform_id], dtype=np.int32).view(np.float32)
    buf[11:13] = local_pos
    buf[13:17] = local_rect
    buf[17] = mode
    for i, (rect, r) in enumerate(zip(corner_rects, (tl, tr, bl, br))):
        buf[18 + 8 * i: 22 + 8 * i] = rect
        buf[22 + 8 * i: 26 + 8 * i] = (r[0], r[1], 0.0, 0.0)
    return buf.view(np.uint8).copy()


def mask_instance(prim_instance, clip_transform_id, clip_address, clip_space=0):
    """MaskInstance, 32 bytes (gpu_types.rs:614-624; vertex.rs:632-650): the quad
    instance followed by aClipData = [clip_transform_id, clip_address, clip_space, 0]."""
    return np.concatenate([np.asarray(prim_instance, dtype=np.int32),
                           np.array([clip_transform_id, clip_address, clip_space, 0], dtype=np.int32)])


def glyph_instance(prim_header_index, clip_task_address, subpx_dir, color_mode, glyph_index, uv_rect_address):
    """GlyphInstance::build → PrimitiveInstanceData (gpu_types.rs:511-528)."""
    z = ((subpx_dir & 0xFF) << 24) | ((color_mode & 0xFF) << 16) | (glyph_index & 0xFFFF)
    return np.array([prim_header_index, clip_task_address, z, uv_rect_address], dtype=np.int64).astype(np.int32)


def build_gradient_table(stops):
    """GradientGpuBlockBuilder::build (prim_store/gradient/mod.rs:166-340), forward
    order: 130 entries of (start_color, step), first/last = clamp entries.
    stops = [(offset, (r, g, b, a) premultiplied)], first offset 0, last 1.
    Returns a (260, 4) float32 array."""
    f32 = np.float32
    N = 130
    start = np.ones((N, 4), dtype=f32)
    step_arr = np.zeros((N, 4), dtype=f32)

    def fill(i0, i1, c0, c1, prev_step):
        inv = f32(1.0) / f32(i1 - i0)
        step = ((c1 - c0) * inv).astype(f32)
        if np.array_equal(step, prev_step):
            a = step[3]
            bits = np.uint32(1) if a == 0.0 else np.float32(a).view(np.uint32) + np.uint32(1)
            step = step.copy()
            step[3] = np.uint32(bits).view(np.float32)
        cur = c0.copy()
        for idx in range(i0, i1):
            start[idx] = cur
            cur = (cur + step).astype(f32)
            step_arr[idx] = step
        return step

    def get_index(off):
        return int(np.floor(f32(min(max(off, 0.0), 1.0)) * f32(128) + f32(1) + f32(0.5)))

    cur_color = np.array(stops[0][1], dtype=f32)
    prev = cur_color
    prev = fill(0, 1, cur_color, cur_color, prev)
    cur_idx = 1
    for off, col in stops[1:]:
        nc = np.array(col, dtype=f32)
        ni = get_index(off)
        if ni > cur_idx:
            prev = fill(cur_idx, ni, cur_color, nc, prev)
            cur_idx = ni
        cur_color = nc
    fill(129, 130, cur_color, cur_color, prev)
    out = np.zeros((260, 4), dtype=f32)
    out[0::2] = start
    out[1::2] = step_arr
    return out


def box_shadow_instance(sub_rect, task_origin, screen_origin, device_pixel_scale, clip_transform_id,
                        prim_transform_id, resource_address, src_rect_size, clip_mode, stretch_mode, dest_rect):
    """ClipMaskInstanceBoxShadow, 84 bytes (gpu_types.rs:226-247; vertex.rs:446-500).
    resource_address = gpu-cache texel index of the shadow mask's uv rect (sent as
    u16 x, u16 y); stretch_mode = (x, y) with 0 = Stretch, 1 = Simple."""
    buf = np.zeros(21, dtype=np.float32)
    buf[0:4] = sub_rect
    buf[4:6] = task_origin
    buf[6:8] = screen_origin
    buf[8] = device_pixel_scale
    ints = buf.view(np.int32)
    ints[9] = clip_transform_id
    ints[10] = prim_transform_id
    buf.view(np.uint16)[22:24] = (resource_address % 1024, resource_address // 1024)
    buf[12:14] = src_rect_size
    ints[14] = clip_mode
    ints[15:17] = stretch_mode
    buf[17:21] = dest_rect
    return buf.view(np.uint8).copy()


def composite_instance(rect, clip_rect, color=(1.0, 1.0, 1.0, 1.0), uv_rect=(0.0, 0.0, 1.0, 1.0), normalized=True,
                       flip=(False, False)):
    """CompositeInstance, 120 bytes (gpu_types.rs:288-356; vertex.rs desc::COMPOSITE):
    rect, clip_rect, premultiplied colour, params [_, uv_type, 0, 0], 3 uv rects, flip."""
    buf = np.zeros(30, dtype=np.float32)
    buf[0:4] = rect
    buf[4:8] = clip_rect
    buf[8:12] = color
    buf[13] = 0.0 if normalized else 1.0   # UV_TYPE_NORMALIZED = 0, UV_TYPE_UNNORMALIZED = 1
    buf[16:20] = uv_rect
    buf[20:24] = uv_rect
    buf[24:28] = uv_rect
    buf[28:30] = (float(flip[0]), float(flip[1]))
    return buf.view(np.uint8).copy()


# YuvFormat / YuvRangedColorSpace (webrender_api/src/image.rs; yuv.glsl:7-11, 27-34)
YUV_FORMAT_NV12, YUV_FORMAT_P010, YUV_FORMAT_NV16, YUV_FORMAT_PLANAR, YUV_FORMAT_INTERLEAVED = 0, 1, 2, 3, 4
(YUV_REC601_NARROW, YUV_REC601_FULL, YUV_REC709_NARROW, YUV_REC709_FULL, YUV_REC2020_NARROW, YUV_REC2020_FULL,
 YUV_GBR_IDENTITY) = range(7)


def composite_yuv_instance(rect, clip_rect, color_space, yuv_format, bit_depth, uv_rects, flip=(False, False)):
    """CompositeInstance::new_yuv (gpu_types.rs:358-378): white colour, params
    [_, colour space, format, channel bit depth], one texel-space uv rect per plane."""
    buf = np.zeros(30, dtype=np.float32)
    buf[0:4] = rect
    buf[4:8] = clip_rect
    buf[8:12] = (1.0, 1.0, 1.0, 1.0)
    buf[13], buf[14], buf[15] = float(color_space), float(yuv_format), float(bit_depth)
    for i in range(3):
        buf[16 + 4 * i:20 + 4 * i] = uv_rects[i]
    buf[28:30] = (float(flip[0]), float(flip[1]))
    return buf.view(np.uint8).copy()


def blur_instance(task_address, src_task_address, direction, std_deviation, blur_region):
    """BlurInstance, 24 bytes (gpu_types.rs:112-118): direction 0 = horizontal, 1 = vertical."""
    buf = np.zeros(6, dtype=np.float32)
    ints = buf.view(np.int32)
    ints[0], ints[1], ints[2] = task_address, src_task_address, direction
    buf[3] = std_deviation
    buf[4:6] = blur_region
    return buf.view(np.uint8).copy()


# ---- cached gradient render tasks (draw_texture_cache_target) -------------------
def fast_linear_gradient_instance(task_rect, color0, color1, axis_select):
    """FastLinearGradientInstance, 52 bytes (prim_store/gradient/linear.rs:689-694)."""
    buf = np.zeros(13, dtype=np.float32)
    buf[0:4], buf[4:8], buf[8:12], buf[12] = task_rect, color0, color1, axis_select
    return buf.view(np.uint8).copy()


def linear_gradient_instance(task_rect, start, end, scale, extend_mode, stops_address):
    """LinearGradientInstance, 48 bytes (prim_store/gradient/linear.rs:727-734)."""
    buf = np.zeros(12, dtype=np.float32)
    buf[0:4], buf[4:6], buf[6:8], buf[8:10] = task_rect, start, end, scale
    ints = buf.view(np.int32)
    ints[10], ints[11] = extend_mode, stops_address
    return buf.view(np.uint8).copy()


def radial_gradient_instance(task_rect, center, scale, start_radius, end_radius, ratio_xy, extend_mode, stops_address):
    """RadialGradientInstance, 52 bytes (prim_store/gradient/radial.rs:375-384)."""
    buf = np.zeros(13, dtype=np.float32)
    buf[0:4], buf[4:6], buf[6:8] = task_rect, center, scale
    buf[8], buf[9], buf[10] = start_radius, end_radius, ratio_xy
    ints = buf.view(np.int32)
    ints[11], ints[12] = extend_mode, stops_address
    return buf.view(np.uint8).copy()


def conic_gradient_instance(task_rect, center, scale, start_offset, end_offset, angle, extend_mode, stops_address):
    """ConicGradientInstance, 52 bytes (prim_store/gradient/conic.rs:409-418)."""
    buf = np.zeros(13, dtype=np.float32)
    buf[0:4], buf[4:6], buf[6:8] = task_rect, center, scale
    buf[8], buf[9], buf[10] = start_offset, end_offset, angle
    ints = buf.view(np.int32)
    ints[11], ints[12] = extend_mode, stops_address
    return buf.view(np.uint8).copy()


# ---- border / line-decoration render tasks ------------------------------------------
SEGMENT_TOP_LEFT, SEGMENT_TOP_RIGHT, SEGMENT_BOTTOM_RIGHT, SEGMENT_BOTTOM_LEFT = 0, 1, 2, 3
SEGMENT_LEFT, SEGMENT_TOP, SEGMENT_RIGHT, SEGMENT_BOTTOM = 4, 5, 6, 7
(BORDER_STYLE_NONE, BORDER_STYLE_SOLID, BORDER_STYLE_DOUBLE, BORDER_STYLE_DOTTED, BORDER_STYLE_DASHED,
 BORDER_STYLE_HIDDEN, BORDER_STYLE_GROOVE, BORDER_STYLE_RIDGE, BORDER_STYLE_INSET, BORDER_STYLE_OUTSET) = range(10)
BORDER_CLIP_NONE, BORDER_CLIP_DASH_CORNER, BORDER_CLIP_DASH_EDGE, BORDER_CLIP_DOT = 0, 1, 2, 3


def border_instance(task_origin, local_rect, color0, color1, segment, style0, style1, do_aa, widths, radius,
                    clip_kind=0, clip_params=(0,) * 8):
    """BorderInstance, 108 bytes (gpu_types.rs:193-202); flags as border.rs:920-923:
    segment | style0 << 8 | style1 << 16 | clip_kind << 24 | do_aa << 28."""
    buf = np.zeros(27, dtype=np.float32)
    buf[0:2], buf[2:6], buf[6:10], buf[10:14] = task_origin, local_rect, color0, color1
    buf.view(np.int32)[14] = segment | (style0 << 8) | (style1 << 16) | (clip_kind << 24) | (int(bool(do_aa)) << 28)
    buf[15:17], buf[17:19], buf[19:27] = widths, radius, clip_params
    return buf.view(np.uint8).copy()


def line_decoration_instance(task_rect, local_size, wavy_line_thickness, style, axis_select):
    """LineDecorationJob, 36 bytes (render_target.rs:1184-1190); style: 0 solid, 1 dotted, 2 dashed, 3 wavy."""
    buf = np.zeros(9, dtype=np.float32)
    buf[0:4], buf[4:6], buf[6] = task_rect, local_size, wavy_line_thickness
    buf.view(np.int32)[7] = style
    buf[8] = axis_select
    return buf.view(np.uint8).copy()
