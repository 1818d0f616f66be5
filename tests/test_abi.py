"""CPU-side checks of the drop-in boundary: the C ABI library loads, exports every symbol
include/felics_b200.h declares, its host-only header functions follow format.rs, and the
compute entry points fail loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import io
import re
from pathlib import Path

import numpy as np
import pytest

import felics_b200
from conftest import ROOT


def declared_functions(header="felics_b200.h"):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(felics_[a-z0-9_]+)\s*\(", text)))


def exported_functions():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "felics_b200" / "libfelics_b200.so")], check=True, capture_output=True, text=True).stdout
    return sorted({line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith("felics_")})


def test_library_exports_every_declared_symbol():
    lib = felics_b200.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for name in names + declared_functions("felics_b200_debug.h"):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_every_exported_symbol_is_declared():
    # the other direction: nothing leaves the library that a header does not declare; the drop-in boundary is
    # felics_b200.h, test and bench aids live in felics_b200_debug.h
    api, dbg = set(declared_functions()), set(declared_functions("felics_b200_debug.h"))
    assert not (api & dbg) and all(n.startswith("felics_debug_") for n in dbg)
    exported = set(exported_functions())
    assert exported == api | dbg, f"undeclared exports: {sorted(exported - api - dbg)}; missing: {sorted((api | dbg) - exported)}"


def test_header_roundtrip_and_layout():  # format.rs:51-61
    buf = io.BytesIO()
    hdr = felics_b200.Header(felics_b200.ColorType.Rgb, felics_b200.PixelDepth.Eight, 0x01020304, 7)
    felics_b200.write_header(hdr, buf)
    assert buf.getvalue() == b"FLCS" + bytes([1, 0, 1, 2, 3, 4, 0, 0, 0, 7])
    assert felics_b200.read_header(buf.getvalue()) == hdr
    assert felics_b200.read_header(io.BytesIO(buf.getvalue() + b"tail")) == hdr


@pytest.mark.parametrize("data,kind", [
    (b"", "IoError"), (b"FLC", "IoError"), (b"FLCX" + bytes(10), "InvalidSignature"),
    (b"FLCS" + bytes([2, 0]) + bytes(8), "InvalidColorType"), (b"FLCS" + bytes([0, 9]) + bytes(8), "InvalidPixelDepth"),
    (b"FLCS" + bytes([1, 1]) + bytes(7), "IoError"),
])
def test_read_header_errors(data, kind):  # format.rs:63-84 check order
    with pytest.raises(felics_b200.DecompressionError) as e:
        felics_b200.read_header(data)
    assert e.value.kind == kind


def test_header_agrees_with_oracle():
    from oracle import felics_oracle as fo
    fel = fo.compress(np.zeros((5, 9, 3), np.uint8))
    h = felics_b200.read_header(fel)
    assert (int(h.color_type), int(h.pixel_depth), h.width, h.height) == fo.read_header(fel)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(felics_b200.FelicsError) as e:
        felics_b200.Codec()
    assert e.value.code == -9 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for path in (ROOT / "felics_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            text = path.read_text()
            assert "felics_oracle" not in text and "oracle/" not in text.replace("forbidden", ""), path
