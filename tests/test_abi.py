"""The C-ABI library loads and exports every symbol include/wrcu.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes as C
import os
import re

from webrender_b200 import abi
from webrender_b200.device import LIB_PATH, load_library

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "wrcu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wrcu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbol_list_matches_abi_module():
    assert _declared_symbols() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = load_library()
    for s in _declared_symbols():
        assert hasattr(lib, s), s


def test_abi_version_and_renderer_string():
    lib = load_library()
    assert lib.wrcu_abi_version() == abi.ABI_VERSION
    # GetString(GL_RENDERER) must keep the host's is_software behaviour (gl.cc:1214)
    assert lib.wrcu_get_string(0) == b"Software WebRender"


def test_no_device_fails_loudly():
    """Without a CUDA device the context must refuse to exist — no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        return
    lib = load_library()
    ctx = C.c_void_p()
    lib.wrcu_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    rc = lib.wrcu_ctx_create(0, C.byref(ctx))
    assert rc == abi.ERR_NO_DEVICE and not ctx.value


def test_program_key_lookup():
    lib = load_library()
    lib.wrcu_program_from_name.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]
    kind, feats = C.c_int(), C.c_uint32()
    for k in range(1, 16):
        for f in (0, abi.FEAT_ALPHA_PASS, abi.FEAT_ALPHA_PASS | abi.FEAT_TEXTURE_2D | abi.FEAT_REPETITION):
            key = abi.program_key(k, f)
            assert lib.wrcu_program_from_name(key.encode(), C.byref(kind), C.byref(feats)) == 0, key
            assert (kind.value, feats.value) == (k, f)
    assert lib.wrcu_program_from_name(b"cs_svg_filter", C.byref(kind), C.byref(feats)) == abi.ERR_UNSUPPORTED
    assert lib.wrcu_program_from_name(b"brush_image TEXTURE_RECT", C.byref(kind), C.byref(feats)) == abi.ERR_UNSUPPORTED


def test_struct_sizes():
    assert C.sizeof(abi.DrawState) == 4 * 2 + 4 * 3 + 4 + 4 + 16 + 16
    assert C.sizeof(abi.FrameTables) == 14 * 8
    assert C.sizeof(abi.Stats) == 40
    assert os.path.exists(LIB_PATH)
