"""cfelics / dfelics front ends (src/bin/cfelics.rs, src/bin/dfelics.rs): flags, messages, exit codes, round trip."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def run(tool, *args):
    return subprocess.run([sys.executable, str(ROOT / "bin" / tool), *args], capture_output=True, text=True)


def test_help_and_missing_arguments():
    for tool, about in (("cfelics", "Compresses an image file to a felics file"), ("dfelics", "Decompresses a felics file to another image file")):
        r = run(tool, "--help")
        assert r.returncode == 0 and about in r.stdout and "--input" in r.stdout and "--output" in r.stdout
        assert run(tool).returncode != 0
        assert run(tool, "--version").stdout.startswith(tool)


def test_missing_input_file(tmp_path):
    r = run("cfelics", "-i", str(tmp_path / "nope.png"), "-o", str(tmp_path / "x.fel"))
    assert r.returncode == 1 and r.stdout.startswith("Cannot open file:")
    r = run("dfelics", "-i", str(tmp_path / "nope.fel"), "-o", str(tmp_path / "x.png"))
    assert r.returncode == 1 and r.stdout.startswith("Cannot open input file:")


def test_undecodable_and_unsupported_input(tmp_path):
    import cv2
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not an image")
    r = run("cfelics", "-i", str(bad), "-o", str(tmp_path / "x.fel"))
    assert r.returncode == 1 and r.stdout.startswith("Cannot decode image:")
    rgba = tmp_path / "rgba.png"
    cv2.imwrite(str(rgba), np.zeros((4, 4, 4), np.uint8))
    r = run("cfelics", "-i", str(rgba), "-o", str(tmp_path / "x.fel"))
    assert r.returncode == 1 and r.stdout.startswith("Unsupported image format: Rgba8")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["gray8", "rgb8", "gray16", "rgb16"])
def test_round_trip_through_files(tmp_path, kind):
    import cv2
    from oracle import felics_oracle as fo
    rng = np.random.default_rng(3)
    dtype = np.uint16 if kind.endswith("16") else np.uint8
    shape = (37, 53, 3) if kind.startswith("rgb") else (37, 53)
    top = 65536 if dtype == np.uint16 else 256
    img = (rng.integers(0, top, shape) // 3 + np.arange(shape[1]).reshape((1, -1) + (1,) * (len(shape) - 2)) * 2).astype(dtype)
    src, fel, back = tmp_path / "in.png", tmp_path / "out.fel", tmp_path / "back.png"
    cv2.imwrite(str(src), img[..., ::-1] if img.ndim == 3 else img)
    r = run("cfelics", "-i", str(src), "-o", str(fel))
    words = {"gray8": "8-bit grayscale", "rgb8": "8-bit rgb", "gray16": "16-bit grayscale", "rgb16": "16-bit rgb"}[kind]
    assert r.returncode == 0 and r.stdout.strip() == f"Compressing {words} image...", r.stdout + r.stderr
    assert fel.read_bytes() == fo.compress(img)
    r = run("dfelics", "-i", str(fel), "-o", str(back))
    assert r.returncode == 0, r.stdout + r.stderr
    got = cv2.imread(str(back), cv2.IMREAD_UNCHANGED)
    got = got[..., ::-1] if got.ndim == 3 else got
    assert got.dtype == img.dtype and np.array_equal(got, img)
    # a truncated file reports the reference's error variant
    (tmp_path / "cut.fel").write_bytes(fel.read_bytes()[:40])
    r = run("dfelics", "-i", str(tmp_path / "cut.fel"), "-o", str(back))
    assert r.returncode == 1 and r.stdout.strip() == "Error while decompressing the image: IoError"


@pytest.mark.gpu
def test_round_trip_with_sidecar(tmp_path):
    # --sidecar is our own extension (not in the reference tools): the felics file stays byte-identical
    import cv2
    from oracle import felics_oracle as fo
    from conftest import gnat_image
    img = gnat_image(1024, 300)
    src, fel, side, back = tmp_path / "in.png", tmp_path / "out.fel", tmp_path / "out.flsc", tmp_path / "back.png"
    cv2.imwrite(str(src), img)
    r = run("cfelics", "-i", str(src), "-o", str(fel), "--sidecar", str(side))
    assert r.returncode == 0, r.stdout + r.stderr
    assert fel.read_bytes() == fo.compress(img) and side.read_bytes()[:4] == b"FLSC"
    r = run("dfelics", "-i", str(fel), "-o", str(back), "--sidecar", str(side))
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(cv2.imread(str(back), cv2.IMREAD_UNCHANGED), img)
    r = run("dfelics", "-i", str(fel), "-o", str(back), "--sidecar", str(tmp_path / "nope.flsc"))
    assert r.returncode == 1 and r.stdout.startswith("Cannot open side file:")
