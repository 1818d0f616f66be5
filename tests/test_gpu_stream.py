"""GPU parity of the streaming band encoder (felics_b200/csrc/stream.cu): batches of gray8 images, one thread
block per image, against the CPU oracle -- bit-exact .fel bytes, lossless decode.  Run on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import felics_b200
from conftest import gnat_image
from felics_b200 import synth
from oracle import felics_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    with felics_b200.Codec(device=0) as c:
        yield c


def first_diff(got, want):
    n = min(len(got), len(want))
    d = next((i for i in range(n) if got[i] != want[i]), n)
    return f"len got {len(got)} want {len(want)}, first differing byte {d}: got {got[d:d + 8].hex()} want {want[d:d + 8].hex()}"


def check_batch(codec, imgs, sample=None, decode=True):
    imgs = np.ascontiguousarray(imgs)
    n = len(imgs)
    arena, offsets = codec.compress_batch(imgs)
    assert offsets[0] == 0 and len(offsets) == n + 1 and int(offsets[n]) == len(arena)
    for i in (range(n) if sample is None else sample):
        got = arena[int(offsets[i]):int(offsets[i + 1])].tobytes()
        want = fo.compress(imgs[i])
        assert got == want, f"image {i}: {first_diff(got, want)}"
    if decode:
        h, w = imgs.shape[1:3]
        hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, w, h)
        out, status = codec.decompress_batch(arena, offsets, hdr)
        assert not status.any() and np.array_equal(out, imgs)
    return arena, offsets


def test_generator_tiles(codec):
    # BASELINE.json configs[3]: 512x512 tiles of the integer generator (64 bands per tile)
    imgs = synth.tile_batch(128)
    before = codec.stream_redone()
    check_batch(codec, imgs)
    assert codec.stream_redone() == before


def test_small_images_single_band(codec):
    rng = np.random.default_rng(5)
    imgs = np.stack([np.clip(gnat_image(64, 48, seed=s).astype(np.int64) + rng.integers(-3, 4, (48, 64)), 0, 255).astype(np.uint8) for s in range(100)])
    check_batch(codec, imgs)


@pytest.mark.parametrize("width,height", [(8, 1), (8, 3), (12, 7), (20, 33), (36, 129), (516, 9), (128, 96), (4100, 3), (8192, 2), (2048, 5)])
def test_shapes(codec, width, height):
    # widths that are multiples of 4 but not of 16 (4-byte copies), rows longer than a band, partial last bands
    rng = np.random.default_rng(width * 1000 + height)
    n = 97
    base = gnat_image(width, height, seed=7).astype(np.int64)
    imgs = np.stack([np.clip(base + rng.integers(-6, 7, base.shape) * (1 + s % 5), 0, 255).astype(np.uint8) for s in range(n)])
    check_batch(codec, imgs)


def test_content_extremes(codec):
    # constant (no out-of-range pixel), uniform noise (all contexts, many short chains), a ramp, one context only,
    # and alternating 0/255 columns: unary runs of 250+ bits, far above the 10 bits per pixel of a slot, so those images
    # come back through the general pipeline
    rng = np.random.default_rng(9)
    h, w = 96, 128
    imgs = []
    for s in range(96):
        m = s % 6
        if m == 0:
            img = np.full((h, w), s, np.uint8)
        elif m == 1:
            img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif m == 2:
            img = ((np.arange(w)[None, :] * 2 + np.arange(h)[:, None]) % 256).astype(np.uint8)
        elif m == 3:
            img = (100 + rng.integers(0, 2, (h, w)) * (1 + s % 3)).astype(np.uint8)
        elif m == 4:
            img = np.where((np.arange(w)[None, :] + np.arange(h)[:, None]) % 2 == 0, 0, 255).astype(np.uint8)
            img[rng.integers(0, h, 40), rng.integers(0, w, 40)] = 128
        else:
            img = np.clip(gnat_image(w, h, sigma=1.0 + s, seed=s), 0, 255).astype(np.uint8)
        imgs.append(img)
    before = codec.stream_redone()
    check_batch(codec, np.stack(imgs))
    assert codec.stream_redone() > before     # the 0/255 images overflowed their slots


def test_more_than_16_bits_per_pixel_but_inside_the_slot_path(codec):
    # bands whose codes exceed the 64 Kbit window while the image as a whole stays below 10 bits per pixel:
    # a few rows of 0/255 stripes inside an otherwise flat image
    h, w = 192, 256
    imgs = []
    for s in range(96):
        img = np.full((h, w), 90 + s % 7, np.uint8)
        r0 = 16 + (s % 5) * 16
        img[r0:r0 + 6] = np.where(np.arange(w)[None, :] % 2 == 0, 0, 255)
        imgs.append(img)
    before = codec.stream_redone()
    check_batch(codec, np.stack(imgs))
    assert codec.stream_redone() == before


def test_natural_crops(codec, golden_images):
    # real photographs: 180 contexts in use, chains of very different lengths in every band
    crops = []
    for name in ("gray8_boat.512", "gray8_5.3.01", "gray8_5.1.09"):
        img = golden_images[name]
        for y0 in range(0, img.shape[0] - 255, 256):
            for x0 in range(0, img.shape[1] - 255, 256):
                crops.append(img[y0:y0 + 256, x0:x0 + 256])
    crops = np.stack(crops)
    reps = (96 + len(crops) - 1) // len(crops)
    imgs = np.concatenate([crops] + [np.ascontiguousarray(crops[:, ::-1]) if r % 2 else np.ascontiguousarray(crops[:, :, ::-1]) for r in range(reps)])
    check_batch(codec, imgs)


def test_long_chains_and_many_halvings(codec):
    # one context takes nearly every pixel: a single chain of ~250k elements per image, thousands of halvings
    rng = np.random.default_rng(21)
    imgs = np.stack([(100 + (rng.integers(0, 100, (512, 512)) < 40 + s % 30) * (1 + s % 2)).astype(np.uint8) for s in range(96)])
    check_batch(codec, imgs, sample=range(0, 96, 5))


def test_stream_and_general_pipeline_agree(codec, monkeypatch):
    imgs = synth.tile_batch(100, first=1000)
    arena, offsets = codec.compress_batch(imgs)
    monkeypatch.setenv("FELICS_B200_NO_STREAM", "1")
    with felics_b200.Codec(device=0) as other:
        arena2, offsets2 = other.compress_batch(imgs)
    assert np.array_equal(offsets, offsets2) and np.array_equal(arena, arena2)


def test_device_entry_point_and_device_generator(codec):
    torch = pytest.importorskip("torch")
    n, first = 200, 4321
    dev = torch.device("cuda", 0)
    d_in = torch.empty(n * 512 * 512, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    codec.generate_tiles(d_in.data_ptr(), first, n)
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, 512, 512)
    cap = n * 512 * 512
    d_out = torch.empty(cap + 3, dtype=torch.uint8, device=dev)
    for shift in (0, 3):     # the arena may start at any byte
        offsets = codec.compress_batch_device(n, d_in.data_ptr(), hdr, d_out.data_ptr() + shift, cap)
        torch.cuda.synchronize()
        host_in = d_in.cpu().numpy().reshape(n, 512, 512)
        assert np.array_equal(host_in[:3], synth.tile_batch(3, first=first)) and np.array_equal(host_in[-2:], synth.tile_batch(2, first=first + n - 2))
        out = d_out.cpu().numpy()[shift:]
        for i in (0, 1, 57, 128, 199):
            got = out[int(offsets[i]):int(offsets[i + 1])].tobytes()
            want = fo.compress(host_in[i])
            assert got == want, f"image {i}: {first_diff(got, want)}"
    # too small an arena: the size it needs comes back
    lib = felics_b200.load_library()
    offs = np.zeros(n + 1, np.uint64)
    chdr = felics_b200._c_header(hdr)
    rc = lib.felics_compress_batch_device(codec._h, n, C.c_void_p(d_in.data_ptr()), C.byref(chdr), C.c_void_p(d_out.data_ptr()), 100000,
                                          offs.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == -8 and int(offs[n]) == int(offsets[n])


def test_host_batch_too_small_arena(codec):
    imgs = synth.tile_batch(96, first=7)
    lib = felics_b200.load_library()
    small = np.empty(500000, np.uint8)
    offs = np.zeros(97, np.uint64)
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, 512, 512)
    chdr = felics_b200._c_header(hdr)
    rc = lib.felics_compress_batch(codec._h, 96, imgs.ctypes.data, C.byref(chdr), small.ctypes.data, small.size, offs.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == -8 and int(offs[96]) > small.size
    arena, offsets = codec.compress_batch(imgs)
    assert int(offsets[96]) == int(offs[96])
