"""felics_b200 -- host-side mirror of the visanalexandru/felics public API over the
B200 C ABI (include/felics_b200.h).

The reference's API (paths under /root/reference/src):
  compress_image(to, image)            compression.rs:412-418
  decompress_image(from)               compression.rs:420-441
  CompressDecompress::{compress, decompress, decompress_with_header}   compression/traits.rs:47-65
  read_header / write_header / Header / ColorType / PixelDepth        compression/format.rs:8-84
  DecompressionError                   compression/error.rs:5-19

Images are numpy arrays: HxW (Luma) or HxWx3 (Rgb), dtype uint8 / uint16 -- the layout of
image::ImageBuffer::as_raw().  All compute happens in hand-written sm_100a CUDA kernels
behind the C ABI; there is no CPU fallback and importing the oracle from here is forbidden.
"""
from __future__ import annotations

import ctypes as C
import enum
import io
from dataclasses import dataclass
from pathlib import Path
from typing import BinaryIO, Optional, Sequence

import numpy as np

__all__ = [
    "ColorType", "PixelDepth", "Header", "DecompressionError", "FelicsError", "Codec",
    "read_header", "write_header", "compress_image", "decompress_image", "load_library",
]

_LIB_PATH = Path(__file__).resolve().parent / "libfelics_b200.so"
_lib: Optional[C.CDLL] = None

HEADER_BYTES = 14


class ColorType(enum.IntEnum):  # format.rs:8-11
    Gray = 0
    Rgb = 1


class PixelDepth(enum.IntEnum):  # format.rs:27-30
    Eight = 0
    Sixteen = 1


@dataclass(frozen=True)
class Header:  # format.rs:44-49
    color_type: ColorType
    pixel_depth: PixelDepth
    width: int
    height: int


class FelicsError(RuntimeError):
    """Failure that is not one of the reference's DecompressionError variants."""

    def __init__(self, code: int, message: str = ""):
        super().__init__(f"felics_b200 error {code}: {message}" if message else f"felics_b200 error {code}")
        self.code = code


class DecompressionError(Exception):
    """compression/error.rs:5-19.  `.kind` is the variant name, `.code` the C ABI code."""

    KINDS = {-1: "IoError", -2: "InvalidValue", -3: "ValueOverflow", -4: "InvalidDimensions",
             -5: "InvalidColorType", -6: "InvalidPixelDepth", -7: "InvalidSignature"}

    def __init__(self, code: int):
        self.code = code
        self.kind = self.KINDS.get(code, f"code {code}")
        super().__init__(self.kind)


class _CHeader(C.Structure):
    _fields_ = [("color_type", C.c_uint8), ("pixel_depth", C.c_uint8), ("width", C.c_uint32), ("height", C.c_uint32)]


def load_library() -> C.CDLL:
    """Load the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(f"{_LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                          "(there is no CPU fallback for the FELICS hot path)")
    L = C.CDLL(str(_LIB_PATH))
    vp, sz, u64p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)
    hp = C.POINTER(_CHeader)
    L.felics_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.felics_ctx_destroy.argtypes = [vp]
    L.felics_ctx_destroy.restype = None
    L.felics_ctx_set_stream.argtypes = [vp, vp]
    L.felics_last_error.restype = C.c_char_p
    L.felics_version.restype = C.c_char_p
    L.felics_read_header.argtypes = [vp, sz, hp]
    L.felics_write_header.argtypes = [hp, vp]
    L.felics_pixel_bytes.argtypes = [hp]
    L.felics_pixel_bytes.restype = sz
    L.felics_compress_bound.argtypes = [hp]
    L.felics_compress_bound.restype = sz
    L.felics_compress.argtypes = [vp, vp, hp, vp, sz, C.POINTER(sz)]
    L.felics_decompress.argtypes = [vp, vp, sz, vp, sz, hp]
    L.felics_compress_device.argtypes = [vp, vp, hp, vp, sz, C.POINTER(sz)]
    L.felics_decompress_device.argtypes = [vp, vp, sz, vp, sz, hp]
    L.felics_compress_batch.argtypes = [vp, sz, vp, hp, vp, sz, u64p]
    L.felics_compress_batch_device.argtypes = [vp, sz, vp, hp, vp, sz, u64p]
    L.felics_decompress_batch.argtypes = [vp, sz, vp, u64p, hp, vp, C.POINTER(C.c_int)]
    L.felics_decompress_batch_device.argtypes = [vp, sz, vp, u64p, hp, vp, C.POINTER(C.c_int)]
    L.felics_compress_batch_v.argtypes = [vp, sz, C.POINTER(vp), hp, vp, sz, u64p]
    L.felics_decompress_batch_v.argtypes = [vp, sz, vp, u64p, C.POINTER(vp), C.POINTER(sz), hp, C.POINTER(C.c_int)]
    L.felics_sidecar_build.argtypes = [vp, C.c_uint32, vp, sz, C.POINTER(sz)]
    L.felics_decompress_sidecar.argtypes = [vp, vp, sz, vp, sz, vp, sz, hp]
    L.felics_profile_enable.argtypes = [vp, C.c_int]
    L.felics_profile_reset.argtypes = [vp]
    L.felics_profile_stage_name.argtypes = [C.c_int]
    L.felics_profile_stage_name.restype = C.c_char_p
    L.felics_profile_stage_ms.argtypes = [vp, C.c_int]
    L.felics_profile_stage_ms.restype = C.c_double
    L.felics_profile_stage_launches.argtypes = [vp, C.c_int]
    L.felics_profile_stage_launches.restype = C.c_uint64
    L.felics_profile_total_launches.argtypes = [vp]
    L.felics_profile_total_launches.restype = C.c_uint64
    # test / bench aids (include/felics_b200_debug.h)
    L.felics_debug_last_records.argtypes = [vp, vp, sz]
    L.felics_debug_counters.argtypes = [vp, vp]
    L.felics_debug_stream_redone.argtypes = [vp]
    L.felics_debug_stream_redone.restype = C.c_uint64
    L.felics_debug_generate_tiles.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64]
    _lib = L
    return L


def _raise(code: int):
    if -7 <= code <= -1:
        raise DecompressionError(code)
    msg = load_library().felics_last_error().decode(errors="replace")
    raise FelicsError(code, msg)


def _header_of(image: np.ndarray) -> Header:
    if image.dtype == np.uint8:
        depth = PixelDepth.Eight
    elif image.dtype == np.uint16:
        depth = PixelDepth.Sixteen
    else:
        raise TypeError("image dtype must be uint8 or uint16")
    if image.ndim == 2:
        color = ColorType.Gray
    elif image.ndim == 3 and image.shape[2] == 3:
        color = ColorType.Rgb
    else:
        raise ValueError("image must be HxW (Luma) or HxWx3 (Rgb)")
    return Header(color, depth, int(image.shape[1]), int(image.shape[0]))


def _c_header(h: Header) -> _CHeader:
    return _CHeader(int(h.color_type), int(h.pixel_depth), int(h.width), int(h.height))


def _shape_of(h: Header):
    return (h.height, h.width, 3) if h.color_type == ColorType.Rgb else (h.height, h.width)


def _dtype_of(h: Header):
    return np.uint16 if h.pixel_depth == PixelDepth.Sixteen else np.uint8


def read_header(source) -> Header:
    """format.rs:63-84.  `source` is bytes or a binary file object (14 bytes are consumed)."""
    data = source if isinstance(source, (bytes, bytearray, memoryview)) else source.read(HEADER_BYTES)
    buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(bytes(data) or b"\0")
    ch = _CHeader()
    rc = load_library().felics_read_header(buf, len(data), C.byref(ch))
    if rc:
        _raise(rc)
    return Header(ColorType(ch.color_type), PixelDepth(ch.pixel_depth), ch.width, ch.height)


def write_header(header: Header, to: BinaryIO) -> None:
    """format.rs:51-61."""
    out = (C.c_uint8 * HEADER_BYTES)()
    ch = _c_header(header)
    rc = load_library().felics_write_header(C.byref(ch), out)
    if rc:
        _raise(rc)
    to.write(bytes(out))


class Codec:
    """One engine context (device + stream + scratch): the object behind compress_image /
    decompress_image.  Use as a context manager or call close()."""

    def __init__(self, device: int = -1):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.felics_ctx_create(device, C.byref(h))
        if rc:
            _raise(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.felics_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- CompressDecompress::compress (traits.rs:48-50) -------------------------------------
    def compress(self, image: np.ndarray) -> bytes:
        image = np.ascontiguousarray(image)
        hdr = _header_of(image)
        ch = _c_header(hdr)
        cap = HEADER_BYTES + 64 + image.nbytes + image.nbytes // 2 + 4096
        for _ in range(2):
            out = np.empty(cap, dtype=np.uint8)
            n = C.c_size_t(0)
            rc = self._lib.felics_compress(self._h, image.ctypes.data, C.byref(ch), out.ctypes.data, cap, C.byref(n))
            if rc == -8:  # FELICS_ERR_BUFFER_TOO_SMALL: n holds the exact size
                cap = n.value + 8
                continue
            if rc:
                _raise(rc)
            return out[: n.value].tobytes()
        _raise(rc)

    # ---- CompressDecompress::decompress (traits.rs:57-64) ------------------------------------
    def decompress(self, fel: bytes, expect: Optional[Header] = None) -> np.ndarray:
        """decompress_image; with `expect` behaves like decompress_with_header for that image
        type (InvalidColorType / InvalidPixelDepth on mismatch, compression.rs:289-294)."""
        hdr = read_header(fel)
        if expect is not None:
            if hdr.color_type != expect.color_type:
                raise DecompressionError(-5)
            if hdr.pixel_depth != expect.pixel_depth:
                raise DecompressionError(-6)
        npix = hdr.width * hdr.height
        if npix > 0xFFFFFFFF:
            raise DecompressionError(-4)
        out = np.zeros(_shape_of(hdr), dtype=_dtype_of(hdr))
        src = np.frombuffer(fel, dtype=np.uint8)
        ch = _CHeader()
        rc = self._lib.felics_decompress(self._h, src.ctypes.data, len(src), out.ctypes.data, max(out.nbytes, 1), C.byref(ch))
        if rc:
            _raise(rc)
        return out

    # ---- band sidecar: opt-in, NOT part of the reference format (include/felics_b200.h) ----------
    def compress_with_sidecar(self, image: np.ndarray, band_rows: int = 0):
        """(fel, sidecar): the reference-format bytes of `compress` plus a side file that lets the bands of the image
        decode in parallel (`decompress_with_sidecar`).  8-bit images of more than two pixels only."""
        fel = self.compress(image)
        n = C.c_size_t(0)
        rc = self._lib.felics_sidecar_build(self._h, band_rows, None, 0, C.byref(n))
        if rc != -8:
            _raise(rc if rc else -12)
        out = np.empty(n.value, dtype=np.uint8)
        rc = self._lib.felics_sidecar_build(self._h, band_rows, out.ctypes.data, out.size, C.byref(n))
        if rc:
            _raise(rc)
        return fel, out.tobytes()

    def decompress_with_sidecar(self, fel: bytes, sidecar: bytes) -> np.ndarray:
        hdr = read_header(fel)
        out = np.zeros(_shape_of(hdr), dtype=_dtype_of(hdr))
        src = np.frombuffer(fel, dtype=np.uint8)
        side = np.frombuffer(sidecar, dtype=np.uint8)
        ch = _CHeader()
        rc = self._lib.felics_decompress_sidecar(self._h, src.ctypes.data, len(src), side.ctypes.data, len(side), out.ctypes.data,
                                                 max(out.nbytes, 1), C.byref(ch))
        if rc:
            _raise(rc)
        return out

    # ---- batches of equally shaped images -------------------------------------------------
    def compress_batch(self, images: np.ndarray):
        """images: (n, H, W) or (n, H, W, 3).  Returns (arena bytes as np.uint8, offsets[n+1])."""
        images = np.ascontiguousarray(images)
        n = images.shape[0]
        hdr = _header_of(images[0]) if n else Header(ColorType.Gray, PixelDepth.Eight, 0, 0)
        ch = _c_header(hdr)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        cap = n * (HEADER_BYTES + 64) + images.nbytes + images.nbytes // 2 + 4096
        for _ in range(2):
            arena = np.empty(cap, dtype=np.uint8)
            rc = self._lib.felics_compress_batch(self._h, n, images.ctypes.data, C.byref(ch), arena.ctypes.data, cap,
                                                 offsets.ctypes.data_as(C.POINTER(C.c_uint64)))
            if rc == -8:
                cap = int(offsets[n]) * 2 + 4096
                continue
            if rc:
                _raise(rc)
            return arena[: int(offsets[n])], offsets
        _raise(rc)

    def decompress_batch(self, arena: np.ndarray, offsets: Sequence[int], header: Header):
        """Returns (images, status[n]); raises only for failures that are not per-image."""
        arena = np.ascontiguousarray(arena, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros((n,) + _shape_of(header), dtype=_dtype_of(header))
        status = np.zeros(max(n, 1), dtype=np.int32)
        ch = _c_header(header)
        rc = self._lib.felics_decompress_batch(self._h, n, arena.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               C.byref(ch), out.ctypes.data, status.ctypes.data_as(C.POINTER(C.c_int)))
        if rc and (rc in (-9, -10, -12) or rc not in status[:n]):   # a failure of the whole call, not of one image
            _raise(rc)
        return out, status[:n]

    # ---- batches of images of different shapes / pixel types (felics_compress_batch_v) ---------
    def compress_many(self, images: Sequence[np.ndarray]):
        """images: any mix of HxW / HxWx3, uint8 / uint16.  Returns (arena bytes as np.uint8, offsets[n+1])."""
        images = [np.ascontiguousarray(im) for im in images]
        n = len(images)
        hdrs = (_CHeader * max(n, 1))(*[_c_header(_header_of(im)) for im in images])
        ptrs = (C.c_void_p * max(n, 1))(*[im.ctypes.data for im in images])
        offsets = np.zeros(n + 1, dtype=np.uint64)
        cap = sum(im.nbytes for im in images) * 3 // 4 + 128 * n + 4096
        for _ in range(2):
            arena = np.empty(cap, dtype=np.uint8)
            rc = self._lib.felics_compress_batch_v(self._h, n, ptrs, hdrs, arena.ctypes.data, cap, offsets.ctypes.data_as(C.POINTER(C.c_uint64)))
            if rc == -8:
                cap = int(offsets[n]) + 64
                continue
            if rc:
                _raise(rc)
            return arena[: int(offsets[n])], offsets
        _raise(rc)

    def decompress_many(self, arena: np.ndarray, offsets: Sequence[int]):
        """Returns (list of images or None, status[n]); shapes and types come from the files' own headers."""
        arena = np.ascontiguousarray(arena, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        outs, caps = [], []
        for i in range(n):
            try:
                h = read_header(arena[int(offsets[i]):int(offsets[i + 1])].tobytes())
                outs.append(np.zeros(_shape_of(h), dtype=_dtype_of(h)))
            except DecompressionError:
                outs.append(np.zeros(1, np.uint8))
            caps.append(outs[-1].nbytes)
        ptrs = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
        capv = (C.c_size_t * max(n, 1))(*caps)
        status = np.zeros(max(n, 1), dtype=np.int32)
        rc = self._lib.felics_decompress_batch_v(self._h, n, arena.ctypes.data, offsets.ctypes.data_as(C.POINTER(C.c_uint64)), ptrs, capv, None,
                                                 status.ctypes.data_as(C.POINTER(C.c_int)))
        if rc and (rc in (-9, -10, -12) or rc not in status[:n]):
            _raise(rc)
        return [o if status[i] == 0 else None for i, o in enumerate(outs)], status[:n]

    # ---- device-resident entry points (raw device pointers, e.g. torch tensors' data_ptr()) ----
    def set_stream(self, cuda_stream: int):
        rc = self._lib.felics_ctx_set_stream(self._h, C.c_void_p(cuda_stream))
        if rc:
            _raise(rc)

    def compress_batch_device(self, n: int, d_pixels: int, header: Header, d_arena: int, arena_cap: int) -> np.ndarray:
        offsets = np.zeros(n + 1, dtype=np.uint64)
        ch = _c_header(header)
        rc = self._lib.felics_compress_batch_device(self._h, n, C.c_void_p(d_pixels), C.byref(ch), C.c_void_p(d_arena), arena_cap,
                                                    offsets.ctypes.data_as(C.POINTER(C.c_uint64)))
        if rc:
            _raise(rc)
        return offsets

    def decompress_batch_device(self, n: int, d_arena: int, offsets: np.ndarray, header: Header, d_pixels_out: int) -> np.ndarray:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        ch = _c_header(header)
        rc = self._lib.felics_decompress_batch_device(self._h, n, C.c_void_p(d_arena), offsets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                      C.byref(ch), C.c_void_p(d_pixels_out), status.ctypes.data_as(C.POINTER(C.c_int)))
        if rc and (rc in (-9, -10, -12) or rc not in status[:n]):   # a failure of the whole call, not of one image
            _raise(rc)
        return status[:n]

    # ---- instrumentation ---------------------------------------------------------------
    def profile(self, on: bool = True):
        self._lib.felics_profile_enable(self._h, int(on))
        self._lib.felics_profile_reset(self._h)

    def profile_reset(self):
        self._lib.felics_profile_reset(self._h)

    def stage_times(self):
        """{stage: (milliseconds, launches)} accumulated since the last reset."""
        out = {}
        for i in range(self._lib.felics_profile_stage_count()):
            name = self._lib.felics_profile_stage_name(i).decode()
            out[name] = (self._lib.felics_profile_stage_ms(self._h, i), int(self._lib.felics_profile_stage_launches(self._h, i)))
        return out

    def total_launches(self) -> int:
        return int(self._lib.felics_profile_total_launches(self._h))

    def debug_counters(self) -> np.ndarray:
        """Device counters of the last encode: [0] live chains, [2] error flags, [3] chains tried by the
        speculative walk, [4] chains it resolved, [5] segments crossed by a hop, [6] hops refused."""
        out = np.zeros(8, dtype=np.uint32)
        self._lib.felics_debug_counters(self._h, out.ctypes.data)
        return out

    def stream_redone(self) -> int:
        """Images the streaming band encoder handed to the general pipeline (streams above 10 bits per pixel)."""
        return int(self._lib.felics_debug_stream_redone(self._h))

    def generate_tiles(self, d_out: int, first_tile: int, n_tiles: int, seed: int = 1):
        """configs[3] tiles generated on the device (felics_b200_debug.h); numpy twin: felics_b200.synth.tile_batch."""
        rc = self._lib.felics_debug_generate_tiles(self._h, C.c_void_p(d_out), first_tile, n_tiles, seed)
        if rc:
            _raise(rc)

    def debug_last_records(self, count: int) -> np.ndarray:
        out = np.zeros(count, dtype=np.uint32)
        rc = self._lib.felics_debug_last_records(self._h, out.ctypes.data, count)
        if rc:
            _raise(rc)
        return out


_default: Optional[Codec] = None


def _default_codec() -> Codec:
    global _default
    if _default is None:
        _default = Codec()
    return _default


def compress_image(to: BinaryIO, image: np.ndarray) -> None:
    """compress_image(to, image) (compression.rs:412-418): writes the .fel bytes to `to`."""
    to.write(_default_codec().compress(image))


def decompress_image(source) -> np.ndarray:
    """decompress_image(from) (compression.rs:420-441): the DynamicImage is the returned
    array's shape/dtype (HxW u8 = ImageLuma8, HxWx3 u8 = ImageRgb8, u16 likewise)."""
    data = source if isinstance(source, (bytes, bytearray, memoryview)) else source.read()
    return _default_codec().decompress(bytes(data))
