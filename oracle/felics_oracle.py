"""ctypes binding of the CPU parity oracle (oracle/felics_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(felics_b200) never imports this module.

The oracle is a C restatement of the reference's channel codec
(/root/reference/src/compression.rs:76-248 and src/coding/*.rs); see the header
of felics_oracle.c for what pins it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libfelics_oracle.so"

OK = 0
IO_ERROR = -1
INVALID_VALUE = -2
VALUE_OVERFLOW = -3
INVALID_DIMENSIONS = -4
INVALID_COLOR_TYPE = -5
INVALID_PIXEL_DEPTH = -6
INVALID_SIGNATURE = -7
BUFFER_TOO_SMALL = -8
PANIC = -11


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (make -C oracle)."""
    src = _HERE / "felics_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u8p, u32p, u64p, i32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32))
        L.felics_oracle_read_header.argtypes = [C.c_void_p, C.c_size_t, u8p, u8p, u32p, u32p]
        L.felics_oracle_compress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.felics_oracle_compress_opts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                                                  C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), u64p]
        L.felics_oracle_compress_bound.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L.felics_oracle_compress_bound.restype = C.c_size_t
        L.felics_oracle_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, u8p, u8p, u32p, u32p]
        L.felics_oracle_rice_bits.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_size_t]
        L.felics_oracle_rice_code_length.argtypes = [C.c_uint32, C.c_uint32]
        L.felics_oracle_rice_code_length.restype = C.c_uint32
        L.felics_oracle_phase_in_params.argtypes = [C.c_uint32, u32p, u32p, u32p]
        L.felics_oracle_phase_in_bits.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_char_p, C.c_size_t]
        L.felics_oracle_codes_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.felics_oracle_kest_new.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32]
        L.felics_oracle_kest_new.restype = C.c_void_p
        L.felics_oracle_kest_update.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.felics_oracle_kest_get_k.argtypes = [C.c_void_p, C.c_uint32]
        L.felics_oracle_kest_entry.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.felics_oracle_kest_entry.restype = C.c_uint32
        L.felics_oracle_kest_free.argtypes = [C.c_void_p]
        L.felics_oracle_kest_free.restype = None
        L.felics_oracle_nearest_neighbours.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p]
        L.felics_oracle_rgb_to_ycocg.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p]
        L.felics_oracle_rgb_to_ycocg.restype = None
        L.felics_oracle_ycocg_to_rgb.argtypes = [C.c_int32, C.c_int32, C.c_int32, i32p]
        L.felics_oracle_ycocg_to_rgb.restype = None
        L.felics_oracle_color_transform8_exhaustive.argtypes = [i32p]
        L.felics_oracle_trace_channel.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64p]
        L.felics_oracle_compress_many.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t, u64p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int):
        super().__init__(f"felics oracle error {code}")
        self.code = code


def _image_params(image: np.ndarray):
    """(color, depth, width, height) of an HxW (gray) or HxWx3 (rgb) u8/u16 array."""
    if image.dtype == np.uint8:
        depth = 0
    elif image.dtype == np.uint16:
        depth = 1
    else:
        raise TypeError("image must be uint8 or uint16")
    if image.ndim == 2:
        color = 0
    elif image.ndim == 3 and image.shape[2] == 3:
        color = 1
    else:
        raise ValueError("image must be HxW or HxWx3")
    return color, depth, image.shape[1], image.shape[0]


def compress(image: np.ndarray, *, k_values=None, scaling=1024, use_transform=True, return_bits=False):
    """compress_image (compression.rs:412): returns the whole .fel file as bytes."""
    image = np.ascontiguousarray(image)
    color, depth, w, h = _image_params(image)
    L = lib()
    cap = 14 + 64 + image.size * image.itemsize * 2 + 1024
    while True:
        out = np.empty(cap, dtype=np.uint8)
        out_len = C.c_size_t(0)
        bits = C.c_uint64(0)
        if k_values is None:
            kv, nk, has = None, 0, 0
        else:
            kv_arr = np.asarray(k_values, dtype=np.uint8)
            kv, nk, has = kv_arr.ctypes.data, len(kv_arr), 0 if scaling is None else 1
        rc = L.felics_oracle_compress_opts(image.ctypes.data, color, depth, w, h, kv, nk, has, scaling or 0,
                                           1 if use_transform else 0, out.ctypes.data, cap, C.byref(out_len), C.byref(bits))
        if rc == BUFFER_TOO_SMALL:
            cap = out_len.value + 16
            continue
        if rc != OK:
            raise OracleError(rc)
        data = out[: out_len.value].tobytes()
        return (data, bits.value) if return_bits else data


def read_header(fel: bytes):
    """read_header (format.rs:63-84) -> (color, depth, width, height)."""
    buf = np.frombuffer(fel, dtype=np.uint8)
    c, d, w, h = C.c_uint8(), C.c_uint8(), C.c_uint32(), C.c_uint32()
    rc = lib().felics_oracle_read_header(buf.ctypes.data if len(buf) else None, len(buf), C.byref(c), C.byref(d), C.byref(w), C.byref(h))
    if rc != OK:
        raise OracleError(rc)
    return c.value, d.value, w.value, h.value


def decompress(fel: bytes) -> np.ndarray:
    """decompress_image (compression.rs:420-441) -> HxW or HxWx3 array."""
    color, depth, w, h = read_header(fel)
    nch = 3 if color else 1
    dtype = np.uint8 if depth == 0 else np.uint16
    if w * h > (1 << 34):
        raise OracleError(INVALID_DIMENSIONS)
    out = np.zeros(max(1, w * h * nch), dtype=dtype)
    buf = np.frombuffer(fel, dtype=np.uint8)
    rc = lib().felics_oracle_decompress(buf.ctypes.data, len(buf), out.ctypes.data, out.nbytes, None, None, None, None)
    if rc != OK:
        raise OracleError(rc)
    out = out[: w * h * nch]
    return out.reshape((h, w, 3)) if color else out.reshape((h, w))


def rice_bits(k: int, number: int, mock: bool = False) -> str:
    buf = C.create_string_buffer(number + 64 + 40)
    rc = lib().felics_oracle_rice_bits(k, number, int(mock), buf, len(buf))
    if rc != OK:
        raise OracleError(rc)
    return buf.value.decode()


def rice_code_length(k: int, number: int) -> int:
    return lib().felics_oracle_rice_code_length(k, number)


def phase_in_params(n: int):
    m, lp, rp = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = lib().felics_oracle_phase_in_params(n, C.byref(m), C.byref(lp), C.byref(rp))
    if rc != OK:
        raise OracleError(rc)
    return m.value, lp.value, rp.value


def phase_in_bits(n: int, number: int, mock: bool = False) -> str:
    buf = C.create_string_buffer(80)
    rc = lib().felics_oracle_phase_in_bits(n, number, int(mock), buf, len(buf))
    if rc != OK:
        raise OracleError(rc)
    return buf.value.decode()


def codes_roundtrip(kinds, a, b):
    """Encode Rice (kind 0: k, value) / phased-in (kind 1: n, value) codes with the
    MSB-first writer, align, decode.  Returns (decoded values, bytes)."""
    kinds = np.ascontiguousarray(kinds, dtype=np.uint32)
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    n = len(kinds)
    dec = np.zeros(n, dtype=np.uint32)
    cap = int(b.sum()) // 8 + 8 * n + 64
    out = np.zeros(cap, dtype=np.uint8)
    nb = C.c_size_t(0)
    rc = lib().felics_oracle_codes_roundtrip(kinds.ctypes.data, a.ctypes.data, b.ctypes.data, n, dec.ctypes.data,
                                             out.ctypes.data, cap, C.byref(nb))
    if rc != OK:
        raise OracleError(rc)
    return dec, out[: nb.value].tobytes()


class KEstimator:
    """KEstimator (parameter_selection.rs:5-86)."""

    def __init__(self, max_context: int, k_values, halve_at=None):
        kv = np.asarray(k_values, dtype=np.uint8)
        if len(kv) == 0:
            raise OracleError(PANIC)  # parameter_selection.rs:25-27
        self._kv = kv
        self._h = lib().felics_oracle_kest_new(max_context, kv.ctypes.data, len(kv), 0 if halve_at is None else 1, halve_at or 0)
        if not self._h:
            raise OracleError(PANIC)

    def update(self, context: int, encoded: int):
        rc = lib().felics_oracle_kest_update(self._h, context, encoded)
        if rc != OK:
            raise OracleError(rc)

    def get_k(self, context: int) -> int:
        k = lib().felics_oracle_kest_get_k(self._h, context)
        if k < 0:
            raise OracleError(PANIC)
        return k

    def entry(self, context: int, ki: int) -> int:
        return lib().felics_oracle_kest_entry(self._h, context, ki)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().felics_oracle_kest_free(self._h)
            self._h = None


def nearest_neighbours(i: int, width: int):
    a, b = C.c_uint64(), C.c_uint64()
    ok = lib().felics_oracle_nearest_neighbours(i, width, C.byref(a), C.byref(b))
    return (a.value, b.value) if ok else None


def rgb_to_ycocg(r, g, b):
    out = (C.c_int32 * 3)()
    lib().felics_oracle_rgb_to_ycocg(r, g, b, out)
    return tuple(out)


def ycocg_to_rgb(y, co, cg):
    out = (C.c_int32 * 3)()
    lib().felics_oracle_ycocg_to_rgb(y, co, cg, out)
    return tuple(out)


def color_transform8_exhaustive():
    r = (C.c_int32 * 6)()
    rc = lib().felics_oracle_color_transform8_exhaustive(r)
    return rc, list(r)


def planes_of(image: np.ndarray):
    """The i32 planes the reference feeds compress_channel (compression.rs:276, :346-356)."""
    image = np.ascontiguousarray(image)
    if image.ndim == 2:
        return [image.astype(np.int32)]
    r, g, b = (image[..., i].astype(np.int32) for i in range(3))
    co = r - b
    t = b + np.trunc(co / 2).astype(np.int32)
    cg = g - t
    y = t + np.trunc(cg / 2).astype(np.int32)
    return [y, co, cg]


def trace_channel(channel: np.ndarray, depth: int = 0):
    """Per-pixel (cls, k, ctx, len) of one i32 plane, plus the plane's bit count."""
    ch = np.ascontiguousarray(channel, dtype=np.int32)
    h, w = ch.shape
    n = max(1, w * h)
    cls = np.zeros(n, dtype=np.uint8)
    k = np.zeros(n, dtype=np.uint8)
    ctx = np.zeros(n, dtype=np.uint32)
    ln = np.zeros(n, dtype=np.uint32)
    bits = C.c_uint64(0)
    rc = lib().felics_oracle_trace_channel(ch.ctypes.data, w, h, depth, cls.ctypes.data, k.ctypes.data, ctx.ctypes.data, ln.ctypes.data, C.byref(bits))
    if rc != OK:
        raise OracleError(rc)
    return cls[: w * h], k[: w * h], ctx[: w * h], ln[: w * h], bits.value


def compress_many(pixels: np.ndarray, color: int, depth: int, width: int, height: int) -> int:
    """Encode pixels[i] for every i on the calling thread; returns total .fel bytes
    (timing harness for the CPU baseline)."""
    pixels = np.ascontiguousarray(pixels)
    n = pixels.shape[0]
    stride = pixels[0].nbytes
    cap = lib().felics_oracle_compress_bound(color, depth, width, height)
    cap = min(cap, 14 + 64 + stride * 3 + 4096)
    scratch = np.empty(cap, dtype=np.uint8)
    tot = C.c_uint64(0)
    rc = lib().felics_oracle_compress_many(pixels.ctypes.data, stride, color, depth, width, height, n, scratch.ctypes.data, cap, C.byref(tot))
    if rc != OK:
        raise OracleError(rc)
    return tot.value
