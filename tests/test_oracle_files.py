"""Whole-file pins of the CPU oracle: hand-derived byte vectors, the reference's
published size tables (DOC.md), the committed corpus manifest and the
round-trip cases of the reference's image-level tests."""
import glob
import hashlib

import numpy as np
import pytest

from conftest import REFERENCE
from oracle import felics_oracle as fo

HDR_3x1 = bytes.fromhex("464c4353" "00" "00" "00000003" "00000001")
RAW_10_20 = bytes.fromhex("0000000a" "00000014")


# ---- known-answer bytes (derived by hand from the reference code) ----
def test_known_answer_in_range():
    # P=15 in [10,20]: '1' then phased-in(n=11): m=3, right_p=5, left_p=3, x=(5+11-3)%11=2 -> '010'
    assert fo.compress(np.array([[10, 20, 15]], np.uint8)) == HDR_3x1 + RAW_10_20 + bytes([0xA0])


def test_known_answer_above():
    # P=25 > H=20: '01', k=5 (untouched context), e=4 -> unary '0' + '00100'
    assert fo.compress(np.array([[10, 20, 25]], np.uint8)) == HDR_3x1 + RAW_10_20 + bytes([0x44])


def test_known_answer_below():
    # P=3 < L=10: '00', e=6 -> '0' + '00110'
    assert fo.compress(np.array([[10, 20, 3]], np.uint8)) == HDR_3x1 + RAW_10_20 + bytes([0x06])


def test_known_answer_rgb_1x1():  # DOC.md:465: (231,27,30) -> (79, 201, -103); compression.rs:99-103
    fel = fo.compress(np.array([[[231, 27, 30]]], np.uint8))
    assert fel == bytes.fromhex("464c4353" "01" "00" "00000001" "00000001"
                                "0000004f" "00000000" "000000c9" "00000000" "ffffff99" "00000000")


def test_known_answer_zero_width():  # compression.rs:94-98, test :457-463
    img = np.zeros((3, 0), np.uint8)
    fel = fo.compress(img)
    assert fel == bytes.fromhex("464c4353" "00" "00" "00000000" "00000003") + bytes(8)
    assert fo.decompress(fel).shape == (3, 0)


def test_header_errors():  # format.rs:63-84
    good = fo.compress(np.array([[1, 2, 3]], np.uint8))
    assert fo.read_header(good) == (0, 0, 3, 1)
    for bad, code in ((b"FLCX" + good[4:], fo.INVALID_SIGNATURE), (good[:4] + b"\x02" + good[5:], fo.INVALID_COLOR_TYPE),
                      (good[:5] + b"\x07" + good[6:], fo.INVALID_PIXEL_DEPTH), (good[:9], fo.IO_ERROR), (b"", fo.IO_ERROR)):
        with pytest.raises(fo.OracleError) as e:
            fo.read_header(bad)
        assert e.value.code == code


def test_truncated_stream_is_io_error():  # BitRead errors map to DecompressionError::IoError (error.rs:21-25)
    img = np.random.default_rng(0).integers(0, 256, (20, 30), dtype=np.uint8)
    fel = fo.compress(img)
    with pytest.raises(fo.OracleError) as e:
        fo.decompress(fel[: len(fel) // 2])
    assert e.value.code == fo.IO_ERROR


# ---- reference image-level tests (compression.rs:457-558) ---------------------------
DIMENSIONS = [(2, 1), (1, 2), (1, 1), (4, 7), (100, 40), (124, 274), (1447, 8), (44, 1), (1, 100), (680, 480)]


@pytest.mark.parametrize("width,height", DIMENSIONS)
def test_compression_decompression_grayscale(width, height):  # compression.rs:500-530
    rng = np.random.default_rng(width * 1000 + height)
    for dtype in (np.uint8, np.uint16):
        img = rng.integers(0, np.iinfo(dtype).max + 1, (height, width), dtype=dtype)
        out = fo.decompress(fo.compress(img))
        assert out.dtype == img.dtype and np.array_equal(out, img)


def test_compression_decompression_intensive():  # compression.rs:544-558 (#[ignore] in the reference)
    rng = np.random.default_rng(7)
    for width in range(0, 20):
        for height in range(0, 20):
            for dtype in (np.uint8, np.uint16):
                g = rng.integers(0, np.iinfo(dtype).max + 1, (height, width), dtype=dtype)
                assert np.array_equal(fo.decompress(fo.compress(g)), g)
                c = rng.integers(0, np.iinfo(dtype).max + 1, (height, width, 3), dtype=dtype)
                assert np.array_equal(fo.decompress(fo.compress(c)), c)


# ---- committed golden images vs. the manifest ------------------------------------------
GOLDEN_FILES = {
    "gray8_5.1.09": "image-suite/grayscale/8bit/5.1.09.tiff",
    "gray8_boat.512": "image-suite/grayscale/8bit/boat.512.tiff",
    "rgb8_lena_color_256": "image-suite/rgb/8bit/lena_color_256.tif",
    "rgb8_pluto": "bench/tiff_files/pluto.tiff",
}


@pytest.mark.parametrize("name", sorted(GOLDEN_FILES))
def test_golden_images_match_manifest(name, golden_images, corpus_manifest):
    entry = next(e for e in corpus_manifest if e["file"] == GOLDEN_FILES[name])
    img = golden_images[name]
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == entry["pixels_sha256"]
    fel = fo.compress(img)
    assert len(fel) == entry["fel_bytes"] and hashlib.sha256(fel).hexdigest() == entry["fel_sha256"]
    assert np.array_equal(fo.decompress(fel), img)


def test_manifest_sums_to_published_totals(corpus_manifest, published_sizes):
    """The per-file oracle sizes in the committed manifest add up to the totals the
    reference publishes (DOC.md:385-396) -- checkable without /root/reference."""
    g8 = sum(e["fel_bytes"] for e in corpus_manifest if e["folder"] == "image-suite/grayscale/8bit")
    g16 = sum(e["fel_bytes"] for e in corpus_manifest if e["folder"] == "image-suite/grayscale/16bit")
    assert g8 == published_sizes["gray8_total_by_kset"]["0-5"] == 8529509
    assert g16 == published_sizes["gray16_total_by_kset"]["0-14"] == 7543288
    by_name = {e["file"].split("/")[-1]: e["fel_bytes"] for e in corpus_manifest if e["folder"] == "image-suite/rgb/8bit"}
    for fname, size in published_sizes["rgb8_with_transform"].items():
        assert by_name[fname] == size
    assert by_name["lena_color_256.tif"] == 110707


def test_boat_and_5109_sizes(golden_images):  # BASELINE.md config 1 anchors
    assert len(fo.compress(golden_images["gray8_boat.512"])) == 168988
    assert len(fo.compress(golden_images["gray8_5.1.09"])) == 42639


# ---- the published tables recomputed from the reference's own TIFFs (only where mounted) ----
needs_reference = pytest.mark.skipif(not REFERENCE.exists(), reason="reference mount absent (GPU box)")


def _load(path):
    from PIL import Image
    return np.ascontiguousarray(np.array(Image.open(path)))


@needs_reference
@pytest.mark.reference
def test_published_gray_tables_from_corpus(published_sizes):  # DOC.md:385-396
    imgs8 = [_load(p) for p in sorted(glob.glob(str(REFERENCE / "image-suite/grayscale/8bit/*")))]
    for kset, total in published_sizes["gray8_total_by_kset"].items():
        lo, hi = map(int, kset.split("-"))
        assert sum(len(fo.compress(a, k_values=list(range(lo, hi + 1)))) for a in imgs8) == total
    imgs16 = [_load(p) for p in sorted(glob.glob(str(REFERENCE / "image-suite/grayscale/16bit/*")))]
    for kset, total in published_sizes["gray16_total_by_kset"].items():
        lo, hi = map(int, kset.split("-"))
        assert sum(len(fo.compress(a, k_values=list(range(lo, hi + 1)))) for a in imgs16) == total


@needs_reference
@pytest.mark.reference
def test_published_rgb_table_from_corpus(published_sizes):  # DOC.md:469-477
    for fname, size in published_sizes["rgb8_with_transform"].items():
        a = _load(REFERENCE / "image-suite/rgb/8bit" / fname)
        assert len(fo.compress(a)) == size
        assert len(fo.compress(a, use_transform=False, k_values=[0, 1, 2, 3, 4, 5])) == published_sizes["rgb8_without_transform"][fname]


@needs_reference
@pytest.mark.reference
def test_compress_suite_roundtrip(corpus_manifest):  # tests/compress.rs:74-103
    for e in corpus_manifest:
        a = _load(REFERENCE / e["file"])
        if a.ndim == 3 and a.shape[2] == 4:
            a = np.ascontiguousarray(a[..., :3])
        fel = fo.compress(a)
        assert len(fel) == e["fel_bytes"] and hashlib.sha256(fel).hexdigest() == e["fel_sha256"], e["file"]
        assert np.array_equal(fo.decompress(fel), a), e["file"]
