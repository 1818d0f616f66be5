"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle -- bit-exact
.fel bytes and lossless decode.  Run on the B200 box: pytest -m gpu."""
import hashlib

import numpy as np
import pytest

import felics_b200
from conftest import gnat_image, gnat_rgb
from oracle import felics_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    with felics_b200.Codec(device=0) as c:
        yield c


def explain_mismatch(codec, img, got, want):
    """Localise a byte mismatch: compare per-pixel code lengths with the oracle's trace."""
    msg = [f"len got {len(got)} want {len(want)}"]
    n = min(len(got), len(want))
    diff = next((i for i in range(n) if got[i] != want[i]), n)
    msg.append(f"first differing byte {diff}: got {got[diff:diff+8].hex()} want {want[diff:diff+8].hex()}")
    try:
        planes = fo.planes_of(img)
        npix = planes[0].size
        recs = codec.debug_last_records(npix * len(planes))
        for pi, pl in enumerate(planes):
            cls, k, ctx, ln, bits = fo.trace_channel(pl)
            glen = (recs[pi * npix:(pi + 1) * npix] >> 22).astype(np.int64)
            bad = np.nonzero(glen != ln.astype(np.int64))[0]
            bad = bad[bad >= 2]
            if len(bad):
                i = int(bad[0])
                msg.append(f"plane {pi}: {len(bad)} pixels with wrong length; first at i={i} (x={i % pl.shape[1]}, y={i // pl.shape[1]}): "
                           f"gpu len {glen[i]} oracle len {ln[i]} cls {cls[i]} k {k[i]} ctx {ctx[i]}")
            else:
                msg.append(f"plane {pi}: all code lengths agree ({bits} bits)")
    except Exception as exc:  # debugging aid only
        msg.append(f"(trace unavailable: {exc})")
    return "; ".join(msg)


def check(codec, img):
    want = fo.compress(img)
    got = codec.compress(img)
    assert got == want, explain_mismatch(codec, img, got, want)
    out = codec.decompress(got)
    assert out.dtype == img.dtype and out.shape == img.shape and np.array_equal(out, img)
    return got


# ---- known-answer bytes --------------------------------------------------
def test_known_answers(codec):
    hdr = bytes.fromhex("464c4353" "00" "00" "00000003" "00000001")
    raw = bytes.fromhex("0000000a" "00000014")
    assert codec.compress(np.array([[10, 20, 15]], np.uint8)) == hdr + raw + bytes([0xA0])
    assert codec.compress(np.array([[10, 20, 25]], np.uint8)) == hdr + raw + bytes([0x44])
    assert codec.compress(np.array([[10, 20, 3]], np.uint8)) == hdr + raw + bytes([0x06])
    assert codec.compress(np.array([[[231, 27, 30]]], np.uint8)) == bytes.fromhex(
        "464c4353" "01" "00" "00000001" "00000001" "0000004f" "00000000" "000000c9" "00000000" "ffffff99" "00000000")


def test_zero_width(codec):  # compression.rs:457-463
    img = np.zeros((3, 0), np.uint8)
    fel = codec.compress(img)
    assert fel == bytes.fromhex("464c4353" "00" "00" "00000000" "00000003") + bytes(8)
    assert codec.decompress(fel).shape == (3, 0)
    img = np.zeros((0, 5, 3), np.uint8)
    assert codec.decompress(check(codec, img)).shape == (0, 5, 3)


DIMENSIONS = [(2, 1), (1, 2), (1, 1), (4, 7), (100, 40), (124, 274), (1447, 8), (44, 1), (1, 100), (680, 480)]


@pytest.mark.parametrize("width,height", DIMENSIONS)
def test_compression_decompression_grayscale(codec, width, height):  # compression.rs:500-530 (u8 half)
    rng = np.random.default_rng(width * 1000 + height)
    check(codec, rng.integers(0, 256, (height, width), dtype=np.uint8))


def test_compression_decompression_intensive(codec):  # compression.rs:544-558, u8 gray + rgb, all sizes below 12
    rng = np.random.default_rng(7)
    for width in range(0, 12):
        for height in range(0, 12):
            check(codec, rng.integers(0, 256, (height, width), dtype=np.uint8))
            check(codec, rng.integers(0, 256, (height, width, 3), dtype=np.uint8))


@pytest.mark.parametrize("width,height", [(4, 7), (100, 40), (333, 65), (1, 9), (2, 2), (640, 3)])
def test_rgb_random(codec, width, height):
    rng = np.random.default_rng(width + 7 * height)
    check(codec, rng.integers(0, 256, (height, width, 3), dtype=np.uint8))


# ---- committed real images (tests/golden) ---------------------------------------------------
@pytest.mark.parametrize("name,file", [
    ("gray8_5.1.09", "image-suite/grayscale/8bit/5.1.09.tiff"), ("gray8_boat.512", "image-suite/grayscale/8bit/boat.512.tiff"),
    ("rgb8_lena_color_256", "image-suite/rgb/8bit/lena_color_256.tif"), ("rgb8_pluto", "bench/tiff_files/pluto.tiff")])
def test_golden_images(codec, golden_images, corpus_manifest, name, file):
    entry = next(e for e in corpus_manifest if e["file"] == file)
    fel = check(codec, golden_images[name])
    assert len(fel) == entry["fel_bytes"] and hashlib.sha256(fel).hexdigest() == entry["fel_sha256"]


# ---- synthetic structure -----------------------------------------------------------------
def test_gnat_1024(codec):
    check(codec, gnat_image(1024, 1024))


def test_uniform_noise_and_constant(codec):
    check(codec, np.random.default_rng(0).integers(0, 256, (512, 512), dtype=np.uint8))
    check(codec, np.full((300, 500), 77, np.uint8))


def test_long_unary_codes(codec):
    # checkerboards / stripes force e up to 254 with small k: code words far longer than 32 bits
    yy, xx = np.mgrid[0:200, 0:300]
    check(codec, (((xx + yy) & 1) * 255).astype(np.uint8))
    check(codec, ((xx % 3 == 0) * 255).astype(np.uint8))
    img = np.zeros((64, 64), np.uint8)
    img[::7, ::5] = 255
    check(codec, img)
    rgb = np.zeros((40, 60, 3), np.uint8)
    rgb[::2, ::3, 0] = 255
    rgb[1::2, ::2, 2] = 255
    check(codec, rgb)


def test_many_halvings_single_context(codec):
    # smooth ramp + small noise keeps nearly all out-of-range pixels in a few contexts -> long chains
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:700, 0:900]
    check(codec, np.clip(100 + xx // 9 + rng.integers(-2, 3, xx.shape), 0, 255).astype(np.uint8))


def test_gnat_rgb(codec):
    check(codec, gnat_rgb(500, 300))


def test_tile_boundaries(codec):
    rng = np.random.default_rng(11)
    for npx in (4095, 4096, 4097, 8192, 8193, 3 * 4096 + 5):
        check(codec, rng.integers(0, 256, (1, npx), dtype=np.uint8))
        check(codec, rng.integers(100, 110, (npx, 1), dtype=np.uint8))


# ---- batches -----------------------------------------------------------------------------
def test_batch_matches_single(codec):
    rng = np.random.default_rng(3)
    imgs = np.stack([gnat_image(128, 96, seed=s) for s in range(5)] + [rng.integers(0, 256, (96, 128), dtype=np.uint8)])
    arena, offsets = codec.compress_batch(imgs)
    assert offsets[0] == 0 and len(offsets) == len(imgs) + 1
    for i, img in enumerate(imgs):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(img), f"image {i}"
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, 128, 96)
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)


def test_batch_rgb(codec):
    imgs = np.stack([gnat_rgb(64, 48), gnat_rgb(64, 48)[::-1].copy(), np.zeros((48, 64, 3), np.uint8)])
    arena, offsets = codec.compress_batch(imgs)
    for i, img in enumerate(imgs):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(img), f"image {i}"
    hdr = felics_b200.Header(felics_b200.ColorType.Rgb, felics_b200.PixelDepth.Eight, 64, 48)
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)


# ---- decode errors -------------------------------------------------------------------------
def test_truncated_stream_is_io_error(codec):
    img = np.random.default_rng(0).integers(0, 256, (20, 30), dtype=np.uint8)
    fel = codec.compress(img)
    with pytest.raises(felics_b200.DecompressionError) as e:
        codec.decompress(fel[: len(fel) // 2])
    assert e.value.kind == "IoError"
    with pytest.raises(felics_b200.DecompressionError) as e:
        codec.decompress(b"FLCX" + fel[4:])
    assert e.value.kind == "InvalidSignature"


def test_decompress_with_header_type_checks(codec):  # compression.rs:289-294, :378-383
    gray = codec.compress(np.zeros((4, 4), np.uint8))
    with pytest.raises(felics_b200.DecompressionError) as e:
        codec.decompress(gray, expect=felics_b200.Header(felics_b200.ColorType.Rgb, felics_b200.PixelDepth.Eight, 4, 4))
    assert e.value.kind == "InvalidColorType"


def test_module_level_api(golden_images):  # compress_image / decompress_image (compression.rs:412-441)
    import io
    img = golden_images["gray8_5.1.09"]
    sink = io.BytesIO()
    felics_b200.compress_image(sink, img)
    assert sink.getvalue() == fo.compress(img)
    assert np.array_equal(felics_b200.decompress_image(io.BytesIO(sink.getvalue())), img)


def test_long_chains(codec):
    # chains longer than 65,536 elements take the table-driven walker (k_walk_long)
    check(codec, gnat_image(2048, 1024))                                           # stationary: one binding counter
    rng = np.random.default_rng(21)
    check(codec, rng.integers(0, 256, (600, 700), dtype=np.uint8).repeat(2, axis=1))  # noise, wide cost ranges
    yy, xx = np.mgrid[0:900, 0:1400]
    check(codec, np.clip(90 + ((xx * 3 + yy) // 11) % 60 + rng.integers(-1, 2, xx.shape), 0, 255).astype(np.uint8))
    check(codec, np.clip(128 + rng.normal(0, 1.2, (1100, 1000)), 0, 255).astype(np.uint8))  # near-tied counters


# ---- 16-bit samples (traits.rs:35-43): K = {0..14}, contexts up to 131070 --------------------
def test_gray16_golden_crop(codec, golden_images):
    check(codec, golden_images["gray16_crop"])


@pytest.mark.parametrize("width,height", [(2, 1), (1, 2), (1, 1), (4, 7), (100, 40), (124, 74), (44, 1), (1, 100)])
def test_compression_decompression_16bit(codec, width, height):  # compression.rs:500-530 (u16 half)
    rng = np.random.default_rng(width * 77 + height)
    check(codec, rng.integers(0, 65536, (height, width), dtype=np.uint16))
    check(codec, rng.integers(0, 65536, (height, width, 3), dtype=np.uint16))


def test_16bit_structure_and_extremes(codec):
    rng = np.random.default_rng(9)
    yy, xx = np.mgrid[0:120, 0:200]
    smooth = np.clip(30000 + 9000 * np.sin(xx / 23.0) * np.cos(yy / 17.0) + rng.normal(0, 40, xx.shape), 0, 65535).astype(np.uint16)
    check(codec, smooth)
    check(codec, np.stack([smooth, np.roll(smooth, 3, 1), smooth[::-1]], axis=-1).copy())
    check(codec, (((xx + yy) & 1) * 65535).astype(np.uint16))          # residuals of 65534 with k = 14 .. 0: very long unary runs
    check(codec, np.full((30, 50), 65535, np.uint16))
    check(codec, np.zeros((3, 0), np.uint16))
    z = np.zeros((20, 30, 3), np.uint16)
    z[::2, ::3, 0] = 65535
    z[1::2, ::2, 2] = 65535
    check(codec, z)                                                     # Co/Cg of +-65535


def test_batch_16bit(codec):
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 65536, (3, 40, 60), dtype=np.uint16)
    imgs[1] >>= 6
    arena, offsets = codec.compress_batch(imgs)
    for i, img in enumerate(imgs):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(img), f"image {i}"
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Sixteen, 60, 40)
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)


def gnat16(width, height, sigma, seed=2, amp=9000.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width]
    v = 30000 + amp * np.sin(xx / 97.0) * np.cos(yy / 131.0) + 0.4 * amp * np.sin((xx + yy) / 37.0) + rng.normal(0, sigma, xx.shape)
    return np.clip(v, 0, 65535).astype(np.uint16)


def test_16bit_parallel_pipeline_larger_images(codec):
    # several tiles per plane and several histogram chunks; every bucket in use; halvings in the popular rows
    check(codec, gnat16(1500, 1100, 300))                 # contexts spread over hundreds of rows' worth of buckets
    check(codec, gnat16(1200, 700, 3, amp=200.0))         # 8-bit-like content: few buckets, thousands of halvings per row
    rng = np.random.default_rng(31)
    check(codec, rng.integers(0, 65536, (500, 600), dtype=np.uint16))       # contexts up to 65535: rows 0..127 of every bucket
    rgb = np.stack([gnat16(700, 500, 120, seed=s) for s in (3, 4, 5)], axis=-1)
    check(codec, np.ascontiguousarray(rgb))
    check(codec, rng.integers(0, 65536, (300, 257, 3), dtype=np.uint16))    # Co/Cg contexts up to 131070: rows up to 255
    # width 1 and 2: first-column / first-row neighbour rules only
    check(codec, gnat16(2, 9000, 50))
    check(codec, gnat16(9000, 1, 50))


def test_16bit_batch_parallel(codec):
    imgs = np.stack([gnat16(333, 129, 20 * (s + 1), seed=s) for s in range(9)])
    arena, offsets = codec.compress_batch(imgs)
    for i, img in enumerate(imgs):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(img), f"image {i}"
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Sixteen, 333, 129)
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)
    rgb = np.stack([np.stack([imgs[i], imgs[(i + 1) % 9], imgs[(i + 2) % 9]], axis=-1) for i in range(4)])
    arena, offsets = codec.compress_batch(rgb)
    for i, img in enumerate(rgb):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(img), f"rgb image {i}"


def test_16bit_random_structured_images(codec):
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=30, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 260), st.integers(1, 100), st.booleans(), st.integers(0, 2**31 - 1), st.sampled_from([0, 1, 7, 300, 20000]),
           st.sampled_from([1, 9, 64]), st.sampled_from(["clip", "wrap", "binary"]))
    def run(width, height, rgb, seed, noise, period, mode):
        rng = np.random.default_rng(seed)
        shape = (height, width, 3) if rgb else (height, width)
        yy, xx = np.mgrid[0:height, 0:width]
        base = 32768 + 25000 * np.sin(xx / period) * np.cos(yy / (period + 2.5))
        if rgb:
            base = np.stack([base, np.roll(base, 2, 1) * 0.7 + 3000, 65535 - base], axis=-1)
        v = base + (rng.integers(-noise, noise + 1, shape) if noise else 0)
        if mode == "clip":
            img = np.clip(v, 0, 65535)
        elif mode == "wrap":
            img = np.mod(v * 3, 65536)
        else:
            img = (v > 32768) * 65535
        check(codec, np.ascontiguousarray(img.astype(np.uint16)))

    run()


def test_16bit_truncated(codec):
    img = np.random.default_rng(1).integers(0, 65536, (20, 30), dtype=np.uint16)
    fel = codec.compress(img)
    with pytest.raises(felics_b200.DecompressionError) as e:
        codec.decompress(fel[: len(fel) // 2])
    assert e.value.kind == "IoError"


def test_speculation_and_segment_hops_are_exercised(codec):
    # one big stationary image: some chains resolve speculatively, rejected ones cross segments by hops, the rest is walked
    img = gnat_image(2048, 2048)
    check(codec, img)
    cnt = codec.debug_counters()
    assert cnt[3] >= 4 and 1 <= cnt[4] < cnt[3], cnt          # chains tried / resolved by the speculative walk
    assert cnt[5] >= 1 and cnt[6] >= 1, cnt                   # hops taken and hops refused (flip inside the segment)


# ---- natural statistics at larger sizes (configs[4]: "synthetic-resized" = mirror-tiled corpus images) ----
def mirror_tile(img, ny, nx):
    rows = [np.concatenate([img[:, ::-1] if (i + j) % 2 else img for j in range(nx)], axis=1) for i in range(ny)]
    return np.ascontiguousarray(np.concatenate([r[::-1] if i % 2 else r for i, r in enumerate(rows)], axis=0))


@pytest.mark.parametrize("name,file", [("gray8_5.3.01", "image-suite/grayscale/8bit/5.3.01.tiff"), ("rgb8_mandril", "image-suite/rgb/8bit/mandril_color.tif")])
def test_golden_natural_images(codec, golden_images, corpus_manifest, name, file):
    entry = next(e for e in corpus_manifest if e["file"] == file)
    fel = check(codec, golden_images[name])
    assert len(fel) == entry["fel_bytes"] and hashlib.sha256(fel).hexdigest() == entry["fel_sha256"]


def test_natural_image_mirror_tiled(codec, golden_images):
    # long chains with natural (non-stationary) statistics: speculation mostly rejected, hops and serial walk mixed
    check(codec, mirror_tile(golden_images["gray8_5.3.01"], 2, 2))      # 2048 x 2048
    check(codec, mirror_tile(golden_images["rgb8_mandril"], 2, 3))      # 1536 x 1024 RGB
    big = mirror_tile(golden_images["gray8_5.3.01"], 3, 4)              # 4096 x 3072
    got = codec.compress(big)
    assert got == fo.compress(big)
    cnt = codec.debug_counters()
    assert cnt[2] == 0


def test_batch_streams_through_sub_batches(codec):
    # 150 images from host memory: five sub-batches, copy-in / encode / copy-out double buffered on three streams
    rng = np.random.default_rng(17)
    imgs = np.stack([np.clip(gnat_image(48, 40, seed=s) + rng.integers(-3, 4, (40, 48)), 0, 255).astype(np.uint8) for s in range(150)])
    arena, offsets = codec.compress_batch(imgs)
    assert len(offsets) == 151 and int(offsets[150]) == len(arena)
    for i in (0, 1, 31, 32, 33, 63, 64, 95, 96, 127, 128, 149):
        assert arena[int(offsets[i]):int(offsets[i + 1])].tobytes() == fo.compress(imgs[i]), f"image {i}"
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, 48, 40)
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)
    # too small an arena reports the size it needs and leaves no copy in flight
    lib = felics_b200.load_library()
    import ctypes as C
    small = np.empty(1000, np.uint8)
    offs = np.zeros(151, np.uint64)
    chdr = felics_b200._c_header(hdr)
    rc = lib.felics_compress_batch(codec._h, 150, imgs.ctypes.data, C.byref(chdr), small.ctypes.data, small.size, offs.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == -8 and int(offs[150]) >= 1000


# ---- randomised structure (hypothesis): shapes, dynamic range, smoothness, channel count ----
def test_random_structured_images(codec):
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 300), st.integers(1, 120), st.booleans(), st.integers(0, 2**31 - 1), st.sampled_from([0, 1, 2, 5, 20, 128]),
           st.sampled_from([1, 3, 17, 64]), st.sampled_from(["clip", "wrap", "binary"]))
    def run(width, height, rgb, seed, noise, period, mode):
        rng = np.random.default_rng(seed)
        shape = (height, width, 3) if rgb else (height, width)
        yy, xx = np.mgrid[0:height, 0:width]
        base = 128 + 100 * np.sin(xx / period) * np.cos(yy / (period + 2.5))
        if rgb:
            base = np.stack([base, np.roll(base, 2, 1) * 0.7 + 30, 255 - base], axis=-1)
        v = base + (rng.integers(-noise, noise + 1, shape) if noise else 0)
        if mode == "clip":
            img = np.clip(v, 0, 255)
        elif mode == "wrap":
            img = np.mod(v * 3, 256)
        else:
            img = (v > 128) * 255
        check(codec, np.ascontiguousarray(img.astype(np.uint8)))

    run()


# ---- band sidecar (opt-in, not the reference format): parallel decode of one big image ----
def test_sidecar_round_trip(codec):
    rng = np.random.default_rng(5)
    cases = [
        (gnat_image(1024, 700), 0),                                             # automatic band height
        (gnat_image(1024, 700), 4),                                             # many bands: band_rows = one tile
        (gnat_image(512, 300), 8),
        (rng.integers(0, 256, (260, 4096), dtype=np.uint8), 2),                 # noise, one row pair per band
        (np.stack([gnat_image(768, 400, seed=s) for s in (3, 4, 5)], axis=-1), 16),   # RGB: bands of three planes
        (gnat_image(333, 200), 0),                                              # odd width: one band per plane only
        (np.stack([gnat_image(333, 200, seed=s) for s in (6, 7, 8)], axis=-1), 0),    # RGB of odd width: three entries whose rows are not a multiple of eight bytes
        (gnat_image(334, 4100), 2048),                                          # width = 2 mod 4 with two bands per plane
        (np.clip(128 + rng.normal(0, 1.2, (2048, 1024)), 0, 255).astype(np.uint8), 32),   # long chains with many halvings before each band
        (gnat_image(2048, 2048), 64),                                           # speculation, hops and serial walk all feed the snapshots
    ]
    for img, rows in cases:
        img = np.ascontiguousarray(img)
        fel, side = codec.compress_with_sidecar(img, rows)
        assert fel == fo.compress(img)                                           # the .fel bytes are the reference format, untouched
        assert np.array_equal(codec.decompress_with_sidecar(fel, side), img), (img.shape, rows)
        assert np.array_equal(codec.decompress(fel), img)


def test_sidecar_rejects_foreign_or_damaged_side_files(codec):
    a, b = gnat_image(1024, 256, seed=1), gnat_image(1024, 256, seed=2)
    fel_a, side_a = codec.compress_with_sidecar(a, 4)
    fel_b, side_b = codec.compress_with_sidecar(b, 4)
    failure = (felics_b200.DecompressionError, felics_b200.FelicsError)
    with pytest.raises(failure):                                                 # bands do not end where the next one starts
        codec.decompress_with_sidecar(fel_a, side_b)
    with pytest.raises(failure):
        codec.decompress_with_sidecar(fel_a, side_a[:-4])
    bad = bytearray(side_a)
    bad[0] ^= 1
    with pytest.raises(failure):
        codec.decompress_with_sidecar(fel_a, bytes(bad))
    assert np.array_equal(codec.decompress_with_sidecar(fel_a, side_a), a)      # the context survives the failures
    with pytest.raises(felics_b200.FelicsError):                                 # band_rows must keep band starts on tile boundaries
        codec.compress_with_sidecar(gnat_image(333, 200), 3)
    with pytest.raises(felics_b200.FelicsError):                                 # 16-bit: no sidecar
        codec.compress_with_sidecar(np.zeros((64, 64), np.uint16))


def test_16bit_serial_encoder_cross_check(monkeypatch):
    # FELICS_B200_SERIAL16=1 selects the one-warp-per-image encoder (the reference loops as written): both device paths
    # and the oracle must agree byte for byte
    img = gnat16(257, 131, 90)
    rgb = np.ascontiguousarray(np.stack([img, np.roll(img, 5, 1), img[::-1]], axis=-1))
    monkeypatch.setenv("FELICS_B200_SERIAL16", "1")
    with felics_b200.Codec(device=0) as serial:
        monkeypatch.delenv("FELICS_B200_SERIAL16")
        with felics_b200.Codec(device=0) as parallel:
            for im in (img, rgb):
                a, b = serial.compress(im), parallel.compress(im)
                assert a == b == fo.compress(im)


def test_sidecar_random_shapes(codec):
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=12, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.sampled_from([8, 64, 256, 512, 1024, 2048, 4096]), st.integers(3, 700), st.booleans(), st.integers(0, 2**31 - 1),
           st.sampled_from([0, 2, 30]), st.integers(1, 6))
    def run(width, height, rgb, seed, noise, mult):
        rng = np.random.default_rng(seed)
        height = max(height, 8192 // width + 3)                      # more than two tiles, so that several bands exist
        shape = (height, width, 3) if rgb else (height, width)
        yy, xx = np.mgrid[0:height, 0:width]
        base = 128 + 90 * np.sin(xx / 23.0) * np.cos(yy / 31.0)
        if rgb:
            base = np.stack([base, 255 - base, np.roll(base, 7, 0)], axis=-1)
        img = np.clip(base + (rng.integers(-noise, noise + 1, shape) if noise else 0), 0, 255).astype(np.uint8)
        unit = 4096 // np.gcd(width, 4096)
        rows = int(unit * mult) if unit * mult >= 2 else 2
        fel, side = codec.compress_with_sidecar(img, rows)
        assert fel == fo.compress(img)
        assert np.array_equal(codec.decompress_with_sidecar(fel, side), img), (shape, rows)

    run()
