"""ctypes binding of include/fpb200.h.  The library is mandatory: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libfpb200.so"
_lib = None


class FpbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fpb200 error {code}: {msg}")
        self.code = code


OK, ERR_INVALID, ERR_CUDA, ERR_IO, ERR_STATE, ERR_NO_DEPTH = 0, -1, -2, -3, -4, -5
TAPS_1X1, TAPS_3X3_S1, TAPS_3X3_S2_PHASE, TAPS_STEM7_ROWS = 0, 1, 2, 3
IPC_HANDLE_BYTES = 64


class ConvDesc(C.Structure):
    _fields_ = [
        ("in_", C.c_void_p),
        ("n_img", C.c_int), ("planes", C.c_int), ("in_h", C.c_int), ("in_w", C.c_int), ("cin", C.c_int),
        ("stride_w", C.c_longlong), ("stride_h", C.c_longlong), ("stride_plane", C.c_longlong),
        ("stride_img", C.c_longlong),
        ("out_h", C.c_int), ("out_w", C.c_int), ("bw", C.c_int), ("bh", C.c_int), ("bn", C.c_int),
        ("tap_kind", C.c_int),
        ("weight", C.c_void_p),
        ("cout", C.c_int),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("res_ld", C.c_int),
        ("post_add", C.c_void_p),
        ("out", C.c_void_p),
        ("relu", C.c_int), ("out_phase_split", C.c_int), ("out_ld", C.c_int), ("img_mod", C.c_int),
        ("half_coff", C.c_int), ("max_ctas", C.c_int), ("cluster", C.c_int), ("block_n", C.c_int),
    ]


class EncoderDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("att", "x0", "w_out", "w1", "w2", "b_out", "b1", "b2", "ln1_g", "ln1_b",
                                          "ln2_g", "ln2_b", "partial", "mean_out")] + \
               [("n_hyp", C.c_int), ("mode", C.c_int)] + [(n, C.c_void_p) for n in ("dbg_x1", "dbg_h", "dbg_y", "dbg_clocks")]


class Mesh(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("vertices", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
        ("faces", C.c_void_p), ("texture", C.c_void_p),
        ("num_vertices", C.c_int), ("num_faces", C.c_int), ("tex_h", C.c_int), ("tex_w", C.c_int),
        ("diameter", C.c_float),
        ("center", C.c_float * 3),
    ]


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("max_hyps", C.c_int), ("shard_rank", C.c_int),
                ("shard_world", C.c_int), ("use_cuda_graph", C.c_int)]


def lib() -> C.CDLL:
    """Load libfpb200.so (building it is the job of build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise FpbError(ERR_IO, f"{_LIB_PATH} is missing: run `python -m foundationpose_cpp_b200.build` "
                                   "(there is no CPU fallback)")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.fpb200_last_error.restype = C.c_char_p
    return _lib


def check(code: int) -> None:
    if code != OK:
        raise FpbError(code, lib().fpb200_last_error().decode(errors="replace"))


def debug_conv_gemm(desc: ConvDesc, repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_conv_gemm(C.byref(desc), C.c_int(repeat), C.byref(ms)))
    return ms.value


def debug_encoder_block(desc: EncoderDesc, repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_encoder_block(C.byref(desc), C.c_int(repeat), C.byref(ms)))
    return ms.value


def debug_attention(qkv_ptr: int, out_ptr: int, n_hyp: int, repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_attention(C.c_void_p(qkv_ptr), C.c_void_p(out_ptr), C.c_int(n_hyp), C.c_int(repeat),
                                       C.byref(ms)))
    return ms.value


def debug_small_linear(in_ptr: int, w_ptr: int, b_ptr: int, out_ptr: int, rows: int, k: int, outs: int,
                       repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_small_linear(C.c_void_p(in_ptr), C.c_void_p(w_ptr), C.c_void_p(b_ptr or None),
                                          C.c_void_p(out_ptr), C.c_int(rows), C.c_int(k), C.c_int(outs),
                                          C.c_int(repeat), C.byref(ms)))
    return ms.value


def debug_cross_attention(qkv_ptr: int, out_ptr: int, n: int, repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_cross_attention(C.c_void_p(qkv_ptr), C.c_void_p(out_ptr), C.c_int(n), C.c_int(repeat),
                                             C.byref(ms)))
    return ms.value


def debug_stem_conv(packed_ptr: int, n_img: int, weight_ptr: int, bias_ptr: int, out_ptr: int, repeat: int = 0) -> float:
    ms = C.c_float(0)
    check(lib().fpb200_debug_stem_conv(C.c_void_p(packed_ptr), C.c_int(n_img), C.c_void_p(weight_ptr), C.c_void_p(bias_ptr),
                                       C.c_void_p(out_ptr), C.c_int(repeat), C.byref(ms)))
    return ms.value
