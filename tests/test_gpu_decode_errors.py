"""Decode-error parity: damaged .fel streams must come back from the CUDA decoder (felics_decompress, C ABI) with the
status the CPU oracle reports -- IoError where the input runs out (compression.rs:161-162, BitRead EOF), the reference's
panics (`context <= max_context`, parameter_selection.rs:72; checked_mul, rice_coding.rs:50) as FELICS_ERR_CORRUPT,
InvalidValue only at the final try_into (compression.rs:305-310), in the reference's order of precedence -- and with the
same pixels when the damage still decodes.  Run on the B200 box: pytest -m gpu."""
import ctypes as C

import numpy as np
import pytest

import felics_b200
from conftest import gnat_image
from oracle import felics_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codec():
    with felics_b200.Codec(device=0) as c:
        yield c


def gpu_decode(codec, fel: bytes, nbytes: int):
    lib = felics_b200.load_library()
    buf = np.frombuffer(fel, dtype=np.uint8)
    out = np.zeros(max(nbytes, 1), np.uint8)
    hdr = felics_b200._CHeader()
    rc = lib.felics_decompress(codec._h, buf.ctypes.data if len(buf) else None, len(buf), out.ctypes.data, nbytes, C.byref(hdr))
    return rc, out[:nbytes]


def oracle_decode(fel: bytes):
    try:
        return 0, np.ascontiguousarray(fo.decompress(fel)).view(np.uint8).reshape(-1)
    except fo.OracleError as e:
        return e.code, None


def images():
    rng = np.random.default_rng(2024)
    g8 = np.clip(gnat_image(44, 37, sigma=6.0, seed=3).astype(np.int64) + rng.integers(-9, 10, (37, 44)), 0, 255).astype(np.uint8)
    rgb8 = np.stack([g8, np.roll(g8, 2, 1), 255 - g8], axis=-1).copy()
    g16 = (g8.astype(np.uint16) << 7) + rng.integers(0, 128, g8.shape).astype(np.uint16)
    rgb16 = np.stack([g16, np.roll(g16, 3, 0), g16[::-1]], axis=-1).copy()
    flat = np.full((20, 32), 200, np.uint8)       # every code is the one-bit in-range marker: damage turns into long unary runs
    return {"gray8": g8, "rgb8": rgb8, "gray16": g16, "rgb16": rgb16, "flat8": flat}


def mutations(fel: bytes, rng, count):
    n = len(fel)
    for j in range(count):
        b = bytearray(fel)
        kind = j % 7
        if kind == 0:                                   # one flipped bit, early positions more often (the decoder then runs on for long)
            pos = 14 + int((n - 14) * rng.random() ** 2)
            b[min(pos, n - 1)] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:                                 # truncated anywhere, header included
            b = b[: int(rng.integers(0, n))]
        elif kind == 2:                                 # a span of ones: long unary runs, samples far outside the pixel type
            pos = int(rng.integers(14, n))
            ln = int(rng.integers(1, 24))
            b[pos:pos + ln] = b"\xff" * len(b[pos:pos + ln])
        elif kind == 3:                                 # a span of zeros
            pos = int(rng.integers(14, n))
            ln = int(rng.integers(1, 24))
            b[pos:pos + ln] = bytes(len(b[pos:pos + ln]))
        elif kind == 4:                                 # random bytes inserted (everything behind them shifts)
            pos = int(rng.integers(14, n))
            b[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 6)), dtype=np.uint8))
        elif kind == 5:                                 # bytes removed
            pos = int(rng.integers(14, n - 1))
            del b[pos:pos + int(rng.integers(1, 6))]
        else:                                           # signature / colour / depth bytes (format.rs:63-84 check order)
            b[int(rng.integers(0, 6))] = int(rng.integers(0, 256))
        yield kind, bytes(b)


@pytest.mark.parametrize("name", ["gray8", "rgb8", "gray16", "rgb16", "flat8"])
def test_damaged_streams_report_what_the_reference_reports(codec, name):
    img = images()[name]
    fel = fo.compress(img)
    rng = np.random.default_rng(7 + sorted(images()).index(name))
    seen = {}
    bad = []
    for kind, mut in mutations(fel, rng, 126):
        want_rc, want_px = oracle_decode(mut)
        # a damaged colour / depth byte changes the size of the decoded image: size the buffer from the (valid) header
        nbytes = img.nbytes
        if want_rc == 0:
            nbytes = want_px.size
        elif len(mut) >= 6 and mut[:4] == b"FLCS" and mut[4] <= 1 and mut[5] <= 1:
            nbytes = img.shape[0] * img.shape[1] * (3 if mut[4] else 1) * (2 if mut[5] else 1)
        got_rc, got_px = gpu_decode(codec, mut, nbytes)
        seen[want_rc] = seen.get(want_rc, 0) + 1
        if got_rc != want_rc or (want_rc == 0 and not np.array_equal(got_px, want_px)):
            bad.append((kind, got_rc, want_rc, len(mut)))
    assert not bad, f"{len(bad)} of 126 mutations differ (kind, gpu, oracle, length): {bad[:10]}; oracle statuses seen: {seen}"
    assert len(seen) >= 3, seen      # the mutations reach several different outcomes


def test_values_that_leave_the_planes_are_settled_by_the_exact_decoder(codec):
    # 0/255 stripes coded, then the stream is cut and padded with ones: above-range codes with long unary runs push
    # samples past 255 and on past the 16-bit planes of the fast decoder
    img = np.where(np.arange(64)[None, :] % 2 == 0, 0, 255).astype(np.uint8).repeat(24, 0)
    fel = fo.compress(img)
    for cut in (40, 200, len(fel) // 2):
        for fill in (b"\x7f" * 4000, b"\xff" * 3000 + b"\x00" * 50, bytes([0x55]) * 5000):
            mut = fel[:cut] + fill
            want_rc, want_px = oracle_decode(mut)
            got_rc, got_px = gpu_decode(codec, mut, img.nbytes)
            assert got_rc == want_rc, (cut, fill[:2], got_rc, want_rc)
            if want_rc == 0:
                assert np.array_equal(got_px, want_px)


def test_batch_offsets_are_validated(codec):
    imgs = np.stack([gnat_image(32, 16, seed=s) for s in range(4)])
    arena, offsets = codec.compress_batch(imgs)
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, 32, 16)
    bad = offsets.copy()
    bad[2], bad[3] = bad[3], bad[2] - 1 if bad[2] else 0      # decreasing
    bad[2] = offsets[4] + 100                                  # and beyond the arena
    with pytest.raises(felics_b200.FelicsError) as e:
        codec.decompress_batch(arena, bad, hdr)
    assert e.value.code == -12
    out, status = codec.decompress_batch(arena, offsets, hdr)
    assert not status.any() and np.array_equal(out, imgs)


@pytest.mark.parametrize("files_per_warp", ["2", "4", "8"])
def test_damaged_files_inside_a_batch_do_not_disturb_their_neighbours(monkeypatch, files_per_warp):
    # the gray batch decoder runs several files per warp, one per lane (k_decode_g8): every file must come back with the
    # oracle's status and pixels whatever happens to the files on the other lanes
    monkeypatch.setenv("FELICS_B200_G8_FILES", files_per_warp)
    img = images()["gray8"]
    rng = np.random.default_rng(99)
    imgs = [np.clip(img.astype(np.int64) + rng.integers(-3, 4, img.shape), 0, 255).astype(np.uint8) for _ in range(45)]
    fels = [fo.compress(im) for im in imgs]
    muts = []
    for i, fel in enumerate(fels):
        if i % 3 == 0:
            muts.append(fel)                                        # every third file stays intact
        else:
            muts.append(next(m for k, m in mutations(fel, rng, 7) if k == (i % 6)))
    offsets = np.zeros(len(muts) + 1, np.uint64)
    offsets[1:] = np.cumsum([len(m) for m in muts])
    arena = np.frombuffer(b"".join(muts), dtype=np.uint8)
    hdr = felics_b200.Header(felics_b200.ColorType.Gray, felics_b200.PixelDepth.Eight, img.shape[1], img.shape[0])
    with felics_b200.Codec(device=0) as c:
        out, status = c.decompress_batch(arena, offsets, hdr)
    seen = set()
    for i, mut in enumerate(muts):
        want_rc, want_px = oracle_decode(mut)
        if want_rc == 0 and want_px.size != img.size:               # a damaged header byte that still parses: another shape, so not this batch's
            want_rc = None
        seen.add(want_rc)
        if want_rc is None:
            assert status[i] != 0, i
            continue
        assert status[i] == want_rc, (i, int(status[i]), want_rc)
        if want_rc == 0:
            assert np.array_equal(out[i].reshape(-1), want_px), i
    assert 0 in seen and len(seen) >= 3, seen
