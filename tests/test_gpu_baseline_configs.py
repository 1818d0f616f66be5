"""GPU parity at the shapes BASELINE.json names (configs[1..4]) and through the mixed-shape batch entry points
(felics_compress_batch_v / felics_decompress_batch_v; the reference's own loop over differently sized files is
tests/compress.rs:74-103).  The oracle is the checker; everything measured goes through the C ABI."""
import json

import numpy as np
import pytest

import felics_b200
from conftest import GOLDEN, gnat_image, gnat_rgb
from felics_b200 import synth
from oracle import felics_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    with felics_b200.Codec(device=0) as c:
        yield c


@pytest.fixture(scope="module")
def bench_corpus():
    return dict(np.load(GOLDEN / "bench_corpus.npz")), json.loads((GOLDEN / "bench_corpus.json").read_text())


def roundtrip(codec, img):
    want = fo.compress(img)
    got = codec.compress(img)
    assert got == want, f"stream differs: {len(got)} bytes against the oracle's {len(want)}"
    out = codec.decompress(got)
    assert out.dtype == img.dtype and out.shape == img.shape and np.array_equal(out, img)
    return got


# ---- configs[1]: one 8192 x 8192 gray8 image, both distributions of SURVEY.md 8(d)2 -------------------------------------
def test_config1_gnat_8192(codec):
    roundtrip(codec, gnat_image(8192, 8192))


def test_config1_uniform_noise_8192(codec):   # the worst case: 8.95 bits per pixel (compression.rs:466-483 at scale)
    fel = roundtrip(codec, np.random.default_rng(0).integers(0, 256, (8192, 8192), dtype=np.uint8))
    assert 8.9 < 8 * len(fel) / 8192 ** 2 < 9.0


# ---- configs[2]: one 7680 x 4320 RGB8 frame ----------------------------------------------------------------------------
def test_config2_rgb_8k_frame(codec):
    roundtrip(codec, gnat_rgb(7680, 4320))


# ---- configs[3]: 512 x 512 tiles of the integer generator, 256 of them, as one batch --------------------------------------
def test_config3_256_generator_tiles(codec):
    first = 65536 - 256   # the last tiles of the full batch: the generator's tile index reaches its largest values
    tiles = synth.tile_batch(256, first=first)
    arena, offsets = codec.compress_batch(tiles)
    for i in range(256):
        want = fo.compress(tiles[i])
        assert bytes(arena[int(offsets[i]):int(offsets[i + 1])]) == want, f"tile {first + i}"
    out, status = codec.decompress_batch(arena, offsets, felics_b200._header_of(tiles[0]))
    assert not status.any() and np.array_equal(out.reshape(tiles.shape), tiles)


# ---- configs[4]: the bench corpus at its native sizes, and mirror-tiled, through the mixed-shape batch ---------------------
def test_config4_corpus_native_sizes(codec, bench_corpus):
    images, meta = bench_corpus
    names = sorted(images)
    arena, offsets = codec.compress_many([images[n] for n in names])
    for i, n in enumerate(names):
        fel = bytes(arena[int(offsets[i]):int(offsets[i + 1])])
        assert len(fel) == meta["images"][n]["fel_bytes"], n
        assert fel == fo.compress(images[n]), n
    assert int(offsets[-1]) == meta["total_fel_bytes"] == 4984136   # DOC.md's table of the seven RGB files
    outs, status = codec.decompress_many(arena, offsets)
    assert not status.any()
    for n, o in zip(names, outs):
        assert o.dtype == images[n].dtype and np.array_equal(o, images[n]), n


def test_config4_mirror_tiled(codec, bench_corpus):
    images, _ = bench_corpus
    tiled = [synth.mirror_tile(images[n], 2048, 2048) for n in sorted(images)[:3]]
    arena, offsets = codec.compress_many(tiled)
    for i, img in enumerate(tiled):
        assert bytes(arena[int(offsets[i]):int(offsets[i + 1])]) == fo.compress(img)
    outs, status = codec.decompress_many(arena, offsets)
    assert not status.any() and all(np.array_equal(o, img) for o, img in zip(outs, tiled))


# ---- mixed-shape batches -----------------------------------------------------------------------------------------------------
def test_mixed_shapes_and_pixel_types(codec):
    rng = np.random.default_rng(21)
    images = []
    for i in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        kind = i % 4
        if kind == 0:
            images.append(rng.integers(90, 130, (h, w), dtype=np.uint8))
        elif kind == 1:
            images.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        elif kind == 2:
            images.append(rng.integers(1000, 1400, (h, w), dtype=np.uint16))
        else:
            images.append(rng.integers(0, 65536, (h, w, 3), dtype=np.uint16))
    images += [images[0].copy(), images[4].copy(), np.zeros((0, 5), np.uint8), np.zeros((3, 0, 3), np.uint16)]   # repeated shapes, empty images
    arena, offsets = codec.compress_many(images)
    assert len(offsets) == len(images) + 1 and offsets[0] == 0
    for i, img in enumerate(images):
        assert bytes(arena[int(offsets[i]):int(offsets[i + 1])]) == fo.compress(img), i
    outs, status = codec.decompress_many(arena, offsets)
    assert not status.any()
    for o, img in zip(outs, images):
        assert o.dtype == img.dtype and o.shape == img.shape and np.array_equal(o, img)


def test_mixed_batch_reports_per_image_errors(codec):
    rng = np.random.default_rng(22)
    images = [rng.integers(0, 256, (30, 20), dtype=np.uint8), rng.integers(0, 256, (9, 40, 3), dtype=np.uint8), rng.integers(0, 256, (30, 20), dtype=np.uint8)]
    arena, offsets = codec.compress_many(images)
    arena = arena.copy()
    arena[int(offsets[1])] ^= 0xff                       # image 1: signature damaged (format.rs:66-69)
    offsets = offsets.copy()
    cut = 25
    arena = arena[: int(offsets[3]) - cut].copy()       # image 2: truncated -> IoError
    offsets[3] -= cut
    outs, status = codec.decompress_many(arena, offsets)
    assert list(status) == [0, -7, -1]
    assert np.array_equal(outs[0], images[0]) and outs[1] is None and outs[2] is None


def test_empty_mixed_batch(codec):
    arena, offsets = codec.compress_many([])
    assert len(arena) == 0 and list(offsets) == [0]


def test_mixed_batch_larger_than_one_chunk(codec):
    # felics_*_batch_v move the images in chunks of about 1 GiB of pixels: eight big images (1.15 GB) of two shapes cross the
    # boundary, the groups of the two chunks interleave; two distinct images keep the oracle's share short (the decoder's side of
    # the chunking runs at small scale in the tests above: a big image decodes for tens of seconds)
    rgb = gnat_rgb(8192, 8192)
    gray = gnat_image(8192, 8192, seed=9)
    order = [rgb, gray, rgb, gray, rgb, rgb, gray, rgb]
    want = {id(rgb): fo.compress(rgb), id(gray): fo.compress(gray)}
    arena, offsets = codec.compress_many(order)
    for i, img in enumerate(order):
        assert bytes(arena[int(offsets[i]):int(offsets[i + 1])]) == want[id(img)], i
    assert int(offsets[-1]) == 5 * len(want[id(rgb)]) + 3 * len(want[id(gray)])


def test_mixed_batch_in_many_small_chunks(monkeypatch):
    # the same chunking with a 64 KB chunk (test switch): 60 small images of four pixel types fall into about twenty chunks,
    # both ways
    monkeypatch.setenv("FELICS_B200_V_CHUNK_KB", "64")
    rng = np.random.default_rng(23)
    images = []
    for i in range(60):
        h, w = int(rng.integers(8, 120)), int(rng.integers(8, 120))
        if i % 4 == 0:
            images.append(rng.integers(100, 140, (h, w), dtype=np.uint8))
        elif i % 4 == 1:
            images.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        elif i % 4 == 2:
            images.append(rng.integers(2000, 2300, (h, w), dtype=np.uint16))
        else:
            images.append(images[i - 3].copy())          # shapes that come back in a later chunk
    with felics_b200.Codec(device=0) as c:
        arena, offsets = c.compress_many(images)
        for i, img in enumerate(images):
            assert bytes(arena[int(offsets[i]):int(offsets[i + 1])]) == fo.compress(img), i
        outs, status = c.decompress_many(arena, offsets)
    assert not status.any()
    for o, img in zip(outs, images):
        assert o.dtype == img.dtype and np.array_equal(o, img)
