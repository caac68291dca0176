"""TEST INFRASTRUCTURE — Python drivers for the two CPU checkers.

* `OracleDevice`: ctypes binding of oracle/libwr_oracle.so (the plain-C
  restatement, wr_oracle.c) — same device calls as the CUDA backend.
* `SwglDevice`: replays the same device calls as the GL call sequence
  `Renderer::draw_frame` issues (renderer/mod.rs:2001-2065, 4418-4841;
  device/gl.rs) against oracle/_ref/libswgl_ref.so = the UNMODIFIED reference
  rasteriser /root/reference/swgl/src/gl.cc built by oracle/Makefile.  It
  contains no arithmetic: only the plumbing `Device` does (textures, FBOs,
  VAOs, uniforms, blend state).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

from webrender_b200 import abi
from webrender_b200.device import DeviceBase, WrcuError, bind_prefixed

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "libwr_oracle.so")
SWGL_LIB = os.path.join(_HERE, "_ref", "libswgl_ref.so")


def have_oracle():
    return os.path.exists(ORACLE_LIB)


def have_swgl():
    return os.path.exists(SWGL_LIB)


class OracleDevice(DeviceBase):
    prefix = "wro_"

    def __init__(self):
        self.lib = C.CDLL(ORACLE_LIB)
        bind_prefixed(self.lib, "wro_")
        self.lib.wro_ctx_create.argtypes = [C.POINTER(C.c_void_p)]
        self.lib.wro_ctx_destroy.argtypes = [C.c_void_p]
        self.lib.wro_shaded_pixels.argtypes = [C.c_void_p]
        self.lib.wro_shaded_pixels.restype = C.c_uint64
        ctx = C.c_void_p()
        rc = self.lib.wro_ctx_create(C.byref(ctx))
        if rc:
            raise WrcuError(rc, "wro_ctx_create")
        self.ctx = ctx

    def shaded_pixels(self):
        return int(self.lib.wro_shaded_pixels(self.ctx))

    def finish(self):
        pass

    def close(self):
        if self.ctx:
            self.lib.wro_ctx_destroy(self.ctx)
            self.ctx = None


# ---- GL constants (swgl/src/gl_defs.h) ---------------------------------------
GL = dict(
    RGBA32F=0x8814, RGBA8=0x8058, R8=0x8229, RG8=0x822B, RG=0x8227, RGBA32I=0x8D82, DEPTH_COMPONENT24=0x81A6,
    UNSIGNED_BYTE=0x1401, UNSIGNED_SHORT=0x1403, INT=0x1404, FLOAT=0x1406,
    RED=0x1903, RGBA=0x1908, RGBA_INTEGER=0x8D99, BGRA=0x80E1,
    ARRAY_BUFFER=0x8892, ELEMENT_ARRAY_BUFFER=0x8893,
    READ_FRAMEBUFFER=0x8CA8, DRAW_FRAMEBUFFER=0x8CA9, COLOR_ATTACHMENT0=0x8CE0, DEPTH_ATTACHMENT=0x8D00,
    COLOR_BUFFER_BIT=0x4000, DEPTH_BUFFER_BIT=0x100,
    NEAREST=0x2600, LINEAR=0x2601, TEXTURE_MAG_FILTER=0x2800, TEXTURE_MIN_FILTER=0x2801,
    TEXTURE_2D=0x0DE1, TEXTURE0=0x84C0, VERTEX_SHADER=0x8B31, FRAGMENT_SHADER=0x8B30,
    BLEND=0x0BE2, DEPTH_TEST=0x0B71, SCISSOR_TEST=0x0C11, TRIANGLES=4,
    ZERO=0, ONE=1, SRC_COLOR=0x300, ONE_MINUS_SRC_COLOR=0x301, SRC_ALPHA=0x302, ONE_MINUS_SRC_ALPHA=0x303,
    DST_ALPHA=0x304, ONE_MINUS_DST_ALPHA=0x305, DST_COLOR=0x306, CONSTANT_COLOR=0x8001,
    CONSTANT_ALPHA=0x8003, SRC1_COLOR=0x88F9, ONE_MINUS_SRC1_COLOR=0x88FA, ONE_MINUS_SRC1_ALPHA=0x88FB,
    FUNC_ADD=0x8006, MIN=0x8007, MAX=0x8008, LESS=0x201, LEQUAL=0x203, STATIC_DRAW=0x88E4,
)
_G = GL

# wrcu_blend → (srgb, drgb, sa, da, equation), device/gl.rs:3901-4017
_BLEND_GL = {
    abi.BLEND_ALPHA: (_G["SRC_ALPHA"], _G["ONE_MINUS_SRC_ALPHA"], _G["ONE"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_PREMULTIPLIED_ALPHA: (_G["ONE"], _G["ONE_MINUS_SRC_ALPHA"], _G["ONE"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_SUBPIXEL_PASS0: (_G["ZERO"], _G["ONE_MINUS_SRC_COLOR"], _G["ZERO"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_SUBPIXEL_PASS0_KEEP_A: (_G["ZERO"], _G["ONE_MINUS_SRC_COLOR"], _G["ZERO"], _G["ONE"], _G["FUNC_ADD"]),
    abi.BLEND_PREMULTIPLIED_DEST_OUT: (_G["ZERO"], _G["ONE_MINUS_SRC_ALPHA"], _G["ZERO"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_MULTIPLY: (_G["ZERO"], _G["SRC_COLOR"], _G["ZERO"], _G["SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_PLUS_LIGHTER: (_G["ONE"], _G["ONE"], _G["ONE"], _G["ONE"], _G["FUNC_ADD"]),
    abi.BLEND_ADD_KEEP_ALPHA_OVER: (_G["ONE"], _G["ONE"], _G["ONE"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_DST_ALPHA_ADD: (_G["ONE_MINUS_DST_ALPHA"], _G["ONE"], _G["ZERO"], _G["ONE"], _G["FUNC_ADD"]),
    abi.BLEND_CONSTANT_COLOR: (_G["CONSTANT_COLOR"], _G["ONE_MINUS_SRC_COLOR"], _G["CONSTANT_ALPHA"], _G["ONE_MINUS_SRC_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_SUBPIXEL_DUAL_SOURCE: (_G["ONE"], _G["ONE_MINUS_SRC1_COLOR"], _G["ONE"], _G["ONE_MINUS_SRC1_ALPHA"], _G["FUNC_ADD"]),
    abi.BLEND_MIN: (_G["ONE"], _G["ONE"], _G["ONE"], _G["ONE"], _G["MIN"]),
    abi.BLEND_MAX: (_G["ONE"], _G["ONE"], _G["ONE"], _G["ONE"], _G["MAX"]),
}
for _i, _eq in enumerate([0x9294, 0x9295, 0x9296, 0x9297, 0x9298, 0x9299, 0x929A, 0x929B, 0x929C, 0x929E,
                          0x92A0, 0x92AD, 0x92AE, 0x92AF, 0x92B0]):
    _BLEND_GL[abi.BLEND_ADV_MULTIPLY + _i] = (_G["ONE"], _G["ONE"], _G["ONE"], _G["ONE"], _eq)

# Texture sampler slots, renderer/mod.rs:369-386
SLOTS = dict(sColor0=0, sColor1=1, sColor2=2, sGpuCache=3, sTransformPalette=4, sRenderTasks=5, sDither=6,
             sPrimitiveHeadersF=7, sPrimitiveHeadersI=8, sClipMask=9, sGpuBufferF=10, sGpuBufferI=11)

# Instance attribute descriptors, renderer/vertex.rs desc::* : (name, count, 'f'|'i'|'u16')
_PRIM = [("aData", 4, "i")]
_CLIP_COMMON = [("aClipDeviceArea", 4, "f"), ("aClipOrigins", 4, "f"), ("aDevicePixelScale", 1, "f"),
                ("aTransformIds", 2, "i")]
_BORDER = [("aTaskOrigin", 2, "f"), ("aRect", 4, "f"), ("aColor0", 4, "f"), ("aColor1", 4, "f"), ("aFlags", 1, "i"),
           ("aWidths", 2, "f"), ("aRadii", 2, "f"), ("aClipParams1", 4, "f"), ("aClipParams2", 4, "f")]
ATTRIBS = {
    abi.KIND_QUAD_TEXTURED: _PRIM, abi.KIND_QUAD_RADIAL_GRADIENT: _PRIM, abi.KIND_QUAD_CONIC_GRADIENT: _PRIM, abi.KIND_BRUSH_SOLID: _PRIM, abi.KIND_BRUSH_IMAGE: _PRIM,
    abi.KIND_BRUSH_LINEAR_GRADIENT: _PRIM, abi.KIND_BRUSH_BLEND: _PRIM, abi.KIND_BRUSH_MIX_BLEND: _PRIM,
    abi.KIND_BRUSH_OPACITY: _PRIM, abi.KIND_TEXT_RUN: _PRIM, abi.KIND_BRUSH_YUV_IMAGE: _PRIM, abi.KIND_SPLIT_COMPOSITE: _PRIM,
    abi.KIND_QUAD_MASK: [("aData", 4, "i"), ("aClipData", 4, "i")],
    abi.KIND_CLIP_RECTANGLE: _CLIP_COMMON + [
        ("aClipLocalPos", 2, "f"), ("aClipLocalRect", 4, "f"), ("aClipMode", 1, "f"),
        ("aClipRect_TL", 4, "f"), ("aClipRadii_TL", 4, "f"), ("aClipRect_TR", 4, "f"), ("aClipRadii_TR", 4, "f"),
        ("aClipRect_BL", 4, "f"), ("aClipRadii_BL", 4, "f"), ("aClipRect_BR", 4, "f"), ("aClipRadii_BR", 4, "f")],
    abi.KIND_CLIP_BOX_SHADOW: _CLIP_COMMON + [
        ("aClipDataResourceAddress", 2, "u16"), ("aClipSrcRectSize", 2, "f"), ("aClipMode", 1, "i"),
        ("aStretchMode", 2, "i"), ("aClipDestRect", 4, "f")],
    abi.KIND_COMPOSITE: [("aDeviceRect", 4, "f"), ("aDeviceClipRect", 4, "f"), ("aColor", 4, "f"),
                         ("aParams", 4, "f"), ("aUvRect0", 4, "f"), ("aUvRect1", 4, "f"), ("aUvRect2", 4, "f"),
                         ("aFlip", 2, "f")],
    abi.KIND_CLEAR: [("aRect", 4, "f"), ("aColor", 4, "f")],
    abi.KIND_BLUR: [("aBlurRenderTaskAddress", 1, "i"), ("aBlurSourceTaskAddress", 1, "i"),
                    ("aBlurDirection", 1, "i"), ("aBlurParams", 3, "f")],
    abi.KIND_SCALE: [("aScaleTargetRect", 4, "f"), ("aScaleSourceRect", 4, "f"), ("aSourceRectType", 1, "f")],
    # cached gradient tasks (prim_store/gradient/{linear,radial,conic}.rs instance structs)
    abi.KIND_FAST_LINEAR_GRADIENT: [("aTaskRect", 4, "f"), ("aColor0", 4, "f"), ("aColor1", 4, "f"),
                                    ("aAxisSelect", 1, "f")],
    abi.KIND_LINEAR_GRADIENT: [("aTaskRect", 4, "f"), ("aStartPoint", 2, "f"), ("aEndPoint", 2, "f"),
                               ("aScale", 2, "f"), ("aExtendMode", 1, "i"), ("aGradientStopsAddress", 1, "i")],
    abi.KIND_RADIAL_GRADIENT: [("aTaskRect", 4, "f"), ("aCenter", 2, "f"), ("aScale", 2, "f"),
                               ("aStartRadius", 1, "f"), ("aEndRadius", 1, "f"), ("aXYRatio", 1, "f"),
                               ("aExtendMode", 1, "i"), ("aGradientStopsAddress", 1, "i")],
    abi.KIND_LINE_DECORATION: [("aTaskRect", 4, "f"), ("aLocalSize", 2, "f"), ("aWavyLineThickness", 1, "f"),
                               ("aStyle", 1, "i"), ("aAxisSelect", 1, "f")],
    abi.KIND_BORDER_SOLID: _BORDER, abi.KIND_BORDER_SEGMENT: _BORDER,
    abi.KIND_CONIC_GRADIENT: [("aTaskRect", 4, "f"), ("aCenter", 2, "f"), ("aScale", 2, "f"),
                              ("aStartOffset", 1, "f"), ("aEndOffset", 1, "f"), ("aAngle", 1, "f"),
                              ("aExtendMode", 1, "i"), ("aGradientStopsAddress", 1, "i")],
}


class SwglDevice:
    """The reference rasteriser behind the wrcu device calls."""

    def __init__(self, lib_path=None):
        """lib_path: any library exporting SWGL's `extern "C"` surface (swgl_fns.rs:23-320) — by
        default the reference build; tests also drive webrender_b200/libwrcu_gl.so through it."""
        self.gl = C.CDLL(lib_path or SWGL_LIB)
        g = self.gl
        g.CreateContext.restype = C.c_void_p
        g.MakeCurrent.argtypes = [C.c_void_p]
        g.DestroyContext.argtypes = [C.c_void_p]
        g.CreateProgram.restype = C.c_uint
        g.CreateShader.restype = C.c_uint
        g.ShaderSourceByName.argtypes = [C.c_uint, C.c_char_p]
        g.BindAttribLocation.argtypes = [C.c_uint, C.c_uint, C.c_char_p]
        g.GetUniformLocation.argtypes = [C.c_uint, C.c_char_p]
        g.GetUniformLocation.restype = C.c_int
        g.GetLinkStatus.restype = C.c_int
        g.TexSubImage2D.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint,
                                    C.c_void_p]
        g.ReadPixels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_void_p]
        g.BufferData.argtypes = [C.c_uint, C.c_ssize_t, C.c_void_p, C.c_uint]
        g.UniformMatrix4fv.argtypes = [C.c_int, C.c_int, C.c_ubyte, C.c_void_p]
        g.VertexAttribPointer.argtypes = [C.c_uint, C.c_int, C.c_uint, C.c_bool, C.c_int, C.c_uint]
        g.VertexAttribIPointer.argtypes = [C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_uint]
        g.ClearColor.argtypes = [C.c_float] * 4
        g.ClearDepth.argtypes = [C.c_double]
        g.BlendColor.argtypes = [C.c_float] * 4
        g.DrawElementsInstanced.argtypes = [C.c_uint, C.c_int, C.c_uint, C.c_ssize_t, C.c_int]
        g.BlitFramebuffer.argtypes = [C.c_int] * 8 + [C.c_uint, C.c_uint]
        self.ctx = g.CreateContext()
        if not self.ctx:
            raise WrcuError(abi.ERR_NO_DEVICE, "CreateContext failed")
        g.MakeCurrent(self.ctx)
        self.tex = {}        # handle -> (fmt, w, h)
        self.fbos = {}       # (color, depth) -> fbo
        self.programs = {}   # key -> (program id, attrib list)
        self.table_tex = {}
        self.gpu_cache_shadow = None
        self.cur = None
        # static quad geometry (renderer/vertex.rs:1075-1076)
        self.quad_vbo = self._gen("GenBuffers")
        self.quad_ibo = self._gen("GenBuffers")
        self.inst_vbo = self._gen("GenBuffers")
        verts = np.array([[0, 0], [0xFF, 0], [0, 0xFF], [0xFF, 0xFF]], dtype=np.uint8)
        idx = np.array([0, 1, 2, 2, 1, 3], dtype=np.uint16)
        self.vaos = {}
        self._verts, self._idx = verts, idx
        g.BindBuffer(_G["ARRAY_BUFFER"], self.quad_vbo)
        g.BufferData(_G["ARRAY_BUFFER"], verts.nbytes, verts.ctypes.data, _G["STATIC_DRAW"])

    def _gen(self, fn):
        out = C.c_uint(0)
        getattr(self.gl, fn)(1, C.byref(out))
        return out.value

    def close(self):
        if self.ctx:
            self.gl.DestroyContext(self.ctx)
            self.ctx = None

    def finish(self):
        pass

    # -- textures --------------------------------------------------------------
    _IFMT = {abi.FMT_RGBA8: _G["RGBA8"], abi.FMT_R8: _G["R8"], abi.FMT_RGBAF32: _G["RGBA32F"],
             abi.FMT_RGBAI32: _G["RGBA32I"], abi.FMT_DEPTH24: _G["DEPTH_COMPONENT24"], abi.FMT_RG8: _G["RG8"]}
    _XFER = {abi.FMT_RGBA8: (_G["BGRA"], _G["UNSIGNED_BYTE"]), abi.FMT_R8: (_G["RED"], _G["UNSIGNED_BYTE"]),
             abi.FMT_RGBAF32: (_G["RGBA"], _G["FLOAT"]), abi.FMT_RGBAI32: (_G["RGBA_INTEGER"], _G["INT"]),
             abi.FMT_RG8: (_G["RG"], _G["UNSIGNED_BYTE"])}

    def texture_create(self, fmt, w, h):
        g = self.gl
        t = self._gen("GenTextures")
        g.ActiveTexture(_G["TEXTURE0"] + 15)
        g.BindTexture(_G["TEXTURE_2D"], t)
        g.TexStorage2D(_G["TEXTURE_2D"], 1, self._IFMT[fmt], w, h)
        self.tex[t] = (fmt, w, h)
        if fmt != abi.FMT_DEPTH24:
            # the reference's targets start zeroed by an explicit clear; make
            # creation deterministic the same way for every backend
            z = np.zeros((h, w * abi.FMT_BPP[fmt]), np.uint8)
            self.texture_upload(t, 0, 0, w, h, z)
        return t

    def texture_set_filter(self, tex, filt):
        g = self.gl
        v = _G["LINEAR"] if filt == abi.LINEAR else _G["NEAREST"]
        g.SetTextureParameter(tex, _G["TEXTURE_MAG_FILTER"], v)
        g.SetTextureParameter(tex, _G["TEXTURE_MIN_FILTER"], v)

    def texture_upload(self, tex, x, y, w, h, data):
        g = self.gl
        fmt = self.tex[tex][0]
        rows = np.ascontiguousarray(data).view(np.uint8).reshape(h, -1)
        rows = np.ascontiguousarray(rows[:, : w * abi.FMT_BPP[fmt]])
        ext, ty = self._XFER[fmt]
        g.ActiveTexture(_G["TEXTURE0"] + 15)
        g.BindTexture(_G["TEXTURE_2D"], tex)
        g.TexSubImage2D(_G["TEXTURE_2D"], 0, x, y, w, h, ext, ty, rows.ctypes.data)

    def texture_destroy(self, tex):
        self.gl.DeleteTexture(tex)
        del self.tex[tex]

    # -- SwCompositor's call sequence (compositor/sw_compositor.rs: lock, composite, unlock) -------
    def sw_composite(self, dst, src, src_rect, dst_rect, opaque, flip_x, flip_y, linear, clip_rect):
        g = self.gl
        g.LockTexture.restype = C.c_void_p
        g.LockTexture.argtypes = [C.c_uint]
        g.UnlockResource.argtypes = [C.c_void_p]
        g.Composite.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_ubyte] * 3 + [C.c_uint] + [C.c_int] * 4
        ld, ls = g.LockTexture(dst), g.LockTexture(src)
        g.Composite(ld, ls, *src_rect, *dst_rect, 1 if opaque else 0, 1 if flip_x else 0, 1 if flip_y else 0,
                    _G["LINEAR"] if linear else _G["NEAREST"], *clip_rect)
        g.UnlockResource(ls)
        g.UnlockResource(ld)

    def sw_composite_yuv(self, dst, y, u, v, color_space, src_rect, dst_rect, flip_x, flip_y, clip_rect, color_depth=8):
        """CompositeYUV (swgl/src/composite.h:1335-1384) between locked textures, as SwCompositor calls it."""
        g = self.gl
        g.LockTexture.restype = C.c_void_p
        g.LockTexture.argtypes = [C.c_uint]
        g.UnlockResource.argtypes = [C.c_void_p]
        g.CompositeYUV.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_uint] + [C.c_int] * 8 + [C.c_ubyte] * 2 + [C.c_int] * 4
        locks = [g.LockTexture(t) for t in (dst, y, u, v)]
        g.CompositeYUV(*locks, int(color_space), int(color_depth), *src_rect, *dst_rect, 1 if flip_x else 0,
                       1 if flip_y else 0, *clip_rect)
        for l in locks[::-1]:
            g.UnlockResource(l)

    def locked_pixels(self, tex):
        """GetResourceBuffer on a locked texture → a copy of what the compositor would read"""
        g = self.gl
        g.LockTexture.restype = C.c_void_p
        g.LockTexture.argtypes = [C.c_uint]
        g.UnlockResource.argtypes = [C.c_void_p]
        g.GetResourceBuffer.restype = C.c_void_p
        g.GetResourceBuffer.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l = g.LockTexture(tex)
        w, h, stride = C.c_int32(), C.c_int32(), C.c_int32()
        p = g.GetResourceBuffer(l, C.byref(w), C.byref(h), C.byref(stride))
        buf = (C.c_uint8 * (stride.value * h.value)).from_address(p)
        out = np.frombuffer(buf, dtype=np.uint8).reshape(h.value, stride.value)[:, : w.value * 4].copy()
        g.UnlockResource(l)
        return out

    # -- update path: the reference's plumbing, call for call -------------------------
    def texture_upload_batch(self, tex, rects, staging):
        """upload_to_texture_cache (renderer/upload.rs): one TexSubImage2D per rect."""
        staging = np.ascontiguousarray(staging).view(np.uint8).reshape(-1)
        bpp = abi.FMT_BPP[self.tex[tex][0]]
        for (x, y, w, h, offset, stride) in rects:
            rows = np.stack([staging[offset + r * stride: offset + r * stride + w * bpp] for r in range(h)])
            self.texture_upload(tex, x, y, w, h, rows)

    def texture_copy(self, src, dst, src_rect, dst_x, dst_y):
        """blit_render_target (device/gl.rs:2310-2370 → BlitFramebuffer, gl.cc)."""
        g = self.gl
        x, y, w, h = src_rect
        g.BindFramebuffer(_G["READ_FRAMEBUFFER"], self._fbo(src, 0))
        g.BindFramebuffer(_G["DRAW_FRAMEBUFFER"], self._fbo(dst, 0))
        g.Disable(_G["SCISSOR_TEST"])
        g.BlitFramebuffer(x, y, x + w, y + h, dst_x, dst_y, dst_x + w, dst_y + h, _G["COLOR_BUFFER_BIT"], _G["NEAREST"])

    def gpu_cache_update(self, height, clear, updates, blocks):
        """GpuCacheBus::PixelBuffer (renderer/gpu_cache.rs:256-290, 324-356): a CPU shadow of the
        rows is patched and the dirty rows re-uploaded to the persistent RGBAF32 texture."""
        blocks = np.ascontiguousarray(blocks, dtype=np.float32).reshape(-1, 4)
        if self.gpu_cache_shadow is None or height > self.gpu_cache_shadow.shape[0]:
            n = np.zeros((height, 1024, 4), np.float32)
            if self.gpu_cache_shadow is not None:
                n[: self.gpu_cache_shadow.shape[0]] = self.gpu_cache_shadow
            self.gpu_cache_shadow = n
        if clear:
            self.gpu_cache_shadow[:] = 0
        for (bi, bc, u, v) in updates:
            self.gpu_cache_shadow[v, u:u + bc] = blocks[bi:bi + bc]

    def _fbo(self, color, depth):
        key = (color, depth)
        if key not in self.fbos:
            g = self.gl
            f = self._gen("GenFramebuffers")
            g.BindFramebuffer(_G["DRAW_FRAMEBUFFER"], f)
            g.FramebufferTexture2D(_G["DRAW_FRAMEBUFFER"], _G["COLOR_ATTACHMENT0"], _G["TEXTURE_2D"], color, 0)
            if depth:
                g.FramebufferTexture2D(_G["DRAW_FRAMEBUFFER"], _G["DEPTH_ATTACHMENT"], _G["TEXTURE_2D"], depth, 0)
            self.fbos[key] = f
        return self.fbos[key]

    def read_pixels(self, tex, x, y, w, h, bpp):
        g = self.gl
        fmt = self.tex[tex][0]
        f = self._fbo(tex, 0)
        g.BindFramebuffer(_G["READ_FRAMEBUFFER"], f)
        out = np.empty((h, w * bpp), dtype=np.uint8)
        ext, ty = self._XFER[fmt]
        g.ReadPixels(x, y, w, h, ext, ty, out.ctypes.data)
        return out

    # -- frame tables: 1024-texel-wide data textures (vertex.rs:877-1038) -------
    def frame_begin(self, tables):
        g = self.gl
        names = dict(prim_headers_f=("sPrimitiveHeadersF", abi.FMT_RGBAF32),
                     prim_headers_i=("sPrimitiveHeadersI", abi.FMT_RGBAI32),
                     transforms=("sTransformPalette", abi.FMT_RGBAF32),
                     render_tasks=("sRenderTasks", abi.FMT_RGBAF32),
                     gpu_cache=("sGpuCache", abi.FMT_RGBAF32),
                     gpu_buffer_f=("sGpuBufferF", abi.FMT_RGBAF32),
                     gpu_buffer_i=("sGpuBufferI", abi.FMT_RGBAI32))
        for key, (sampler, fmt) in names.items():
            src = tables[key]
            if src is None and key == "gpu_cache":
                src = self.gpu_cache_shadow.reshape(-1, 4)
            arr = np.ascontiguousarray(src)
            n = arr.size // 4
            rows = max(1, (n + 1023) // 1024)
            buf = np.zeros((rows * 1024, 4), dtype=arr.dtype if n else (np.float32 if fmt == abi.FMT_RGBAF32 else np.int32))
            if n:
                buf[:n] = arr.reshape(n, 4)
            old = self.table_tex.get(key)
            if old is not None:
                self.texture_destroy(old)
            t = self._gen("GenTextures")
            g.ActiveTexture(_G["TEXTURE0"] + SLOTS[sampler])
            g.BindTexture(_G["TEXTURE_2D"], t)
            g.TexStorage2D(_G["TEXTURE_2D"], 1, self._IFMT[fmt], 1024, rows)
            self.tex[t] = (fmt, 1024, rows)
            ext, ty = self._XFER[fmt]
            g.TexSubImage2D(_G["TEXTURE_2D"], 0, 0, 0, 1024, rows, ext, ty, buf.ctypes.data)
            self.table_tex[key] = t

    def frame_end(self):
        pass

    # -- targets ----------------------------------------------------------------
    def target_bind(self, color, depth, projection, viewport):
        g = self.gl
        f = self._fbo(color, depth)
        g.BindFramebuffer(_G["DRAW_FRAMEBUFFER"], f)
        g.SetViewport(*viewport)
        self.cur = (color, depth, np.asarray(projection, np.float32).copy())

    def clear(self, rect, color, depth):
        g = self.gl
        if rect is not None:
            g.Enable(_G["SCISSOR_TEST"])
            g.SetScissor(*rect)
        else:
            g.Disable(_G["SCISSOR_TEST"])
        mask = 0
        if color is not None:
            g.ClearColor(*[float(c) for c in color])
            mask |= _G["COLOR_BUFFER_BIT"]
        if depth is not None and self.cur[1]:
            g.ClearDepth(float(depth))
            g.DepthMask(1)
            mask |= _G["DEPTH_BUFFER_BIT"]
        g.Clear(mask)
        g.Disable(_G["SCISSOR_TEST"])
        # SWGL defers whole-target clears (gl.cc:2342-2353); resolve so later
        # sampling of this texture sees the cleared values, as a later draw or
        # ReadPixels would.
        g.ResolveFramebuffer(self._fbo(self.cur[0], self.cur[1]))

    # -- programs ---------------------------------------------------------------
    def _program(self, kind, features):
        key = abi.program_key(kind, features)
        if key in self.programs:
            return self.programs[key]
        g = self.gl
        p = g.CreateProgram()
        vs = g.CreateShader(_G["VERTEX_SHADER"])
        fs = g.CreateShader(_G["FRAGMENT_SHADER"])
        g.ShaderSourceByName(vs, key.encode())
        g.ShaderSourceByName(fs, key.encode())
        g.AttachShader(p, vs)
        g.AttachShader(p, fs)
        if not g.GetLinkStatus(p):
            raise WrcuError(abi.ERR_UNSUPPORTED, f"SWGL reference has no program '{key}'")
        attribs = [("aPosition", 2, "u8n")] + ATTRIBS[kind]
        for i, (name, _, _) in enumerate(attribs):
            g.BindAttribLocation(p, i, name.encode())
        g.LinkProgram(p)
        g.UseProgram(p)
        for name, slot in SLOTS.items():
            loc = g.GetUniformLocation(p, name.encode())
            if loc >= 0:
                g.Uniform1i(loc, slot)
        self.programs[key] = (p, attribs)
        return self.programs[key]

    def _vao(self, kind, attribs, stride):
        k = (kind, stride)
        if k in self.vaos:
            return self.vaos[k]
        g = self.gl
        v = self._gen("GenVertexArrays")
        g.BindVertexArray(v)
        g.BindBuffer(_G["ELEMENT_ARRAY_BUFFER"], self.quad_ibo)
        g.BufferData(_G["ELEMENT_ARRAY_BUFFER"], self._idx.nbytes, self._idx.ctypes.data, _G["STATIC_DRAW"])
        g.BindBuffer(_G["ARRAY_BUFFER"], self.quad_vbo)
        g.EnableVertexAttribArray(0)
        g.VertexAttribPointer(0, 2, _G["UNSIGNED_BYTE"], True, 2, 0)
        g.VertexAttribDivisor(0, 0)
        g.BindBuffer(_G["ARRAY_BUFFER"], self.inst_vbo)
        off = 0
        for i, (name, count, ty) in enumerate(attribs[1:], start=1):
            g.EnableVertexAttribArray(i)
            if ty == "f":
                g.VertexAttribPointer(i, count, _G["FLOAT"], False, stride, off)
                off += 4 * count
            elif ty == "i":
                g.VertexAttribIPointer(i, count, _G["INT"], stride, off)
                off += 4 * count
            elif ty == "u16":
                g.VertexAttribIPointer(i, count, _G["UNSIGNED_SHORT"], stride, off)
                off += 2 * count
            g.VertexAttribDivisor(i, 1)
        assert off <= stride, (off, stride)
        self.vaos[k] = v
        return v

    def draw_batch(self, kind, features, blend, depth, colors, clip_mask, scissor, blend_color, inst):
        g = self.gl
        p, attribs = self._program(kind, features)
        g.UseProgram(p)
        loc = g.GetUniformLocation(p, b"uTransform")
        proj = np.ascontiguousarray(self.cur[2], dtype=np.float32)
        g.UniformMatrix4fv(loc, 1, 0, proj.ctypes.data)
        # bind_textures (renderer/mod.rs:2001-2020)
        for i in range(3):
            g.ActiveTexture(_G["TEXTURE0"] + i)
            g.BindTexture(_G["TEXTURE_2D"], colors[i])
        g.ActiveTexture(_G["TEXTURE0"] + SLOTS["sClipMask"])
        g.BindTexture(_G["TEXTURE_2D"], clip_mask)
        for key, sampler in (("prim_headers_f", "sPrimitiveHeadersF"), ("prim_headers_i", "sPrimitiveHeadersI"),
                             ("transforms", "sTransformPalette"), ("render_tasks", "sRenderTasks"),
                             ("gpu_cache", "sGpuCache"), ("gpu_buffer_f", "sGpuBufferF"),
                             ("gpu_buffer_i", "sGpuBufferI")):
            g.ActiveTexture(_G["TEXTURE0"] + SLOTS[sampler])
            g.BindTexture(_G["TEXTURE_2D"], self.table_tex[key])
        # blend / depth / scissor state
        if blend == abi.BLEND_NONE:
            g.Disable(_G["BLEND"])
        else:
            s, d, sa, da, eq = _BLEND_GL[blend]
            g.Enable(_G["BLEND"])
            g.BlendEquation(eq)
            g.BlendFunc(s, d, sa, da)
            g.BlendColor(*[float(c) for c in blend_color])
        if depth == abi.DEPTH_OFF or not self.cur[1]:
            g.Disable(_G["DEPTH_TEST"])
        else:
            g.Enable(_G["DEPTH_TEST"])
            g.DepthFunc(_G["LEQUAL"])
            g.DepthMask(1 if depth == abi.DEPTH_TEST_WRITE else 0)
        if scissor is not None:
            g.Enable(_G["SCISSOR_TEST"])
            g.SetScissor(*scissor)
        else:
            g.Disable(_G["SCISSOR_TEST"])
        inst = np.ascontiguousarray(inst)
        n, stride = inst.shape
        v = self._vao(kind, attribs, stride)
        g.BindVertexArray(v)
        g.BindBuffer(_G["ARRAY_BUFFER"], self.inst_vbo)
        g.BufferData(_G["ARRAY_BUFFER"], inst.nbytes, inst.ctypes.data, _G["STATIC_DRAW"])
        g.DrawElementsInstanced(_G["TRIANGLES"], 6, _G["UNSIGNED_SHORT"], 0, n)
