"""The reference's own FFI surface over the CUDA backend: webrender_b200/libwrcu_gl.so exports the
`extern "C"` symbol set WebRender binds for its software rasteriser (swgl/src/swgl_fns.rs:23-320,
defined in swgl/src/gl.cc and composite.h).  CPU tier: every symbol is exported.  GPU tier: the GL call
sequence `Device` issues (the same driver that runs the UNMODIFIED reference rasteriser in
oracle/backends.py::SwglDevice) is replayed against the shim and must reproduce the oracle's pixels."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.backends import OracleDevice, SwglDevice
from webrender_b200 import abi, scenes

from common import assert_same, render

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GL_LIB = os.path.join(ROOT, "webrender_b200", "libwrcu_gl.so")

# the names in swgl_fns.rs's extern block (99)
SWGL_SYMBOLS = [
    "ActiveTexture", "AttachShader", "BeginQuery", "BindAttribLocation", "BindBuffer", "BindFramebuffer",
    "BindRenderbuffer", "BindTexture", "BindVertexArray", "BlendColor", "BlendEquation", "BlendFunc",
    "BlitFramebuffer", "BufferData", "BufferSubData", "CheckFramebufferStatus", "Clear", "ClearColor",
    "ClearColorRect", "ClearDepth", "ClearTexImage", "ClearTexSubImage", "Composite", "CompositeYUV",
    "CopyImageSubData", "CopyTexSubImage2D", "CreateContext", "CreateProgram", "CreateShader", "DeleteBuffer",
    "DeleteFramebuffer", "DeleteProgram", "DeleteQuery", "DeleteRenderbuffer", "DeleteShader",
    "DeleteTexture", "DeleteVertexArray", "DepthFunc", "DepthMask", "DestroyContext", "Disable",
    "DrawElementsInstanced", "Enable", "EnableVertexAttribArray", "EndQuery", "Finish",
    "FramebufferRenderbuffer", "FramebufferTexture2D", "GenBuffers", "GenFramebuffers", "GenQueries",
    "GenRenderbuffers", "GenTextures", "GenVertexArrays", "GenerateMipmap", "GetAttribLocation",
    "GetBooleanv", "GetColorBuffer", "GetError", "GetIntegerv", "GetLinkStatus", "GetQueryObjectui64v",
    "GetResourceBuffer", "GetString", "GetStringi", "GetUniformLocation", "InitDefaultFramebuffer",
    "InvalidateFramebuffer", "LinkProgram", "LockFramebuffer", "LockResource", "LockTexture", "MakeCurrent",
    "MapBuffer", "MapBufferRange", "PixelStorei", "ReadPixels", "ReferenceContext", "RenderbufferStorage",
    "ReportMemory", "ResolveFramebuffer", "SetScissor", "SetTextureBuffer", "SetTextureParameter",
    "SetViewport", "ShaderSourceByName", "TexImage2D", "TexParameteri", "TexStorage2D", "TexSubImage2D",
    "Uniform1i", "Uniform4fv", "UniformMatrix4fv", "UnlockResource", "UnmapBuffer", "UseProgram",
    "VertexAttribDivisor", "VertexAttribIPointer", "VertexAttribPointer"
]


def test_gl_library_exports_the_reference_symbol_set():
    assert len(SWGL_SYMBOLS) == 99
    lib = C.CDLL(GL_LIB)
    for s in SWGL_SYMBOLS:
        assert hasattr(lib, s), s


class GlShimDevice(SwglDevice):
    def __init__(self):
        super().__init__(lib_path=GL_LIB)


CASES = [
    ("alpha_rects", lambda: scenes.alpha_rects_frame(640, 360, 60, random_rects=True, seed=4, color=None), None),
    ("alpha_rects_softlight", lambda: scenes.alpha_rects_frame(320, 200, 30, random_rects=True, seed=7, color=None,
                                                                 blend=abi.BLEND_ADV_SOFT_LIGHT), None),
    ("brush_solid_masks_depth", lambda: scenes.brush_solid_frame(seed=1), ["target"]),
    ("clip_masks", lambda: scenes.clip_mask_frame(seed=2, fractional=True), None),
    ("rounded_rects_indirect", lambda: scenes.rounded_rects_frame(seed=1), None),
    ("images", lambda: scenes.image_frame(seed=1, one_to_one=True), ["target"]),
    ("text_subpixel", lambda: scenes.text_frame(seed=2, width=480, height=270, n_runs=8, glyphs_per_run=20,
                                                atlas="rgba", color_modes=(0, 1, 2, 3)), ["target"]),
    ("gradients", lambda: scenes.gradient_frame(seed=1, blend=abi.BLEND_PREMULTIPLIED_ALPHA), None),
    ("box_shadows", lambda: scenes.box_shadow_frame(seed=1), None),
    ("composite", lambda: scenes.composite_frame(seed=1), ["fb"]),
    ("composite_yuv_nv12", lambda: scenes.yuv_composite_frame("nv12", 3, seed=2), ["fb"]),
    ("brush_yuv_image", lambda: scenes.yuv_image_frame("planar", 1, seed=2, fractional=True), ["target"]),
    ("composite_yuv_planar", lambda: scenes.yuv_composite_frame("planar", 4, seed=3, fractional=True), ["fb"]),
    ("blur", lambda: scenes.blur_frame(seed=1), ["mid", "target"]),
    ("texture_cache_target", lambda: scenes.texture_cache_frame(seed=1), None),
    ("quad_radial", lambda: scenes.quad_gradient_frame(abi.KIND_QUAD_RADIAL_GRADIENT, seed=2), None),
    ("config_a", lambda: scenes.config_a_frame(), None),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,make,targets", CASES, ids=[c[0] for c in CASES])
def test_gl_call_sequence_on_the_cuda_backend(name, make, targets):
    frame = make()
    assert_same(render(GlShimDevice, frame, targets), render(OracleDevice, frame, targets), name)


@pytest.mark.gpu
def test_gl_strings_keep_the_host_on_the_software_path():
    lib = C.CDLL(GL_LIB)
    lib.CreateContext.restype = C.c_void_p
    lib.MakeCurrent.argtypes = [C.c_void_p]
    lib.DestroyContext.argtypes = [C.c_void_p]
    lib.GetString.restype = C.c_char_p
    ctx = lib.CreateContext()
    assert ctx
    lib.MakeCurrent(ctx)
    assert lib.GetString(0x1F01) == b"Software WebRender"   # GL_RENDERER → Device::is_software (device/gl.rs:1645)
    v = C.c_int(0)
    lib.GetIntegerv(0x0D33, C.byref(v))
    assert v.value == 1 << 15                               # GL_MAX_TEXTURE_SIZE, as gl.cc:1158
    assert lib.GetError() == 0
    lib.DestroyContext(ctx)
