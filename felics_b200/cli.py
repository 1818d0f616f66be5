"""cfelics / dfelics: the reference's command-line tools (src/bin/cfelics.rs, src/bin/dfelics.rs)
over the B200 engine.  Same flags (-i/--input, -o/--output, --version, --help), same messages,
exit status 1 on every failure.  Image files are read and written with OpenCV (any format it
supports); the codec itself runs on the GPU through the C ABI."""
from __future__ import annotations

import argparse
import sys

import numpy as np

VERSION = "0.1.0"


def _read_image(path: str) -> np.ndarray:
    import cv2
    # the reference prints "Cannot open file" for a missing file and "Cannot decode image" for a bad one
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        print(f"Cannot open file: {e}")
        raise SystemExit(1)
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    if img is None:
        print("Cannot decode image: unsupported or corrupt image file")
        raise SystemExit(1)
    return img


def _color_name(img: np.ndarray) -> str:
    ch = 1 if img.ndim == 2 else img.shape[2]
    depth = {np.dtype(np.uint8): "8", np.dtype(np.uint16): "16", np.dtype(np.float32): "32F"}.get(img.dtype, str(img.dtype))
    return {1: "L", 2: "La", 3: "Rgb", 4: "Rgba"}.get(ch, f"{ch}ch") + depth


def cfelics_main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="cfelics", description="Compresses an image file to a felics file")
    ap.add_argument("-i", "--input", required=True, help="The input file.")
    ap.add_argument("-o", "--output", required=True, help="The output felics file.")
    ap.add_argument("-V", "--version", action="version", version=f"cfelics {VERSION}")
    ap.add_argument("--sidecar", metavar="FILE", help="(not in the reference tool) also write a band side file that lets dfelics "
                    "decode the image band-parallel; 8-bit images only. The felics file itself is unchanged.")
    args = ap.parse_args(argv)
    img = _read_image(args.input)
    kinds = {(2, np.dtype(np.uint8)): "8-bit grayscale", (2, np.dtype(np.uint16)): "16-bit grayscale",
             (3, np.dtype(np.uint8)): "8-bit rgb", (3, np.dtype(np.uint16)): "16-bit rgb"}
    key = (img.ndim if img.ndim == 2 or img.shape[2] == 3 else 0, img.dtype)
    if key not in kinds:
        print(f"Unsupported image format: {_color_name(img)}")
        return 1
    print(f"Compressing {kinds[key]} image...")
    if img.ndim == 3:
        img = np.ascontiguousarray(img[..., ::-1])   # OpenCV decodes to B, G, R
    try:
        import felics_b200
        if args.sidecar:
            fel, side = felics_b200._default_codec().compress_with_sidecar(img)
            with open(args.output, "wb") as out:
                out.write(fel)
            with open(args.sidecar, "wb") as out:
                out.write(side)
        else:
            with open(args.output, "wb") as out:
                felics_b200.compress_image(out, img)
    except Exception as e:   # io::Error in the reference
        print(f"Cannot compress image: {e}")
        return 1
    return 0


def dfelics_main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="dfelics", description="Decompresses a felics file to another image file")
    ap.add_argument("-i", "--input", required=True, help="The input felics file.")
    ap.add_argument("-o", "--output", required=True,
                    help="The output file. The output format will be determined using the extension of the output file.")
    ap.add_argument("-V", "--version", action="version", version=f"dfelics {VERSION}")
    ap.add_argument("--sidecar", metavar="FILE", help="(not in the reference tool) the side file written by cfelics --sidecar")
    args = ap.parse_args(argv)
    try:
        f = open(args.input, "rb")
    except OSError as e:
        print(f"Cannot open input file: {e}")
        return 1
    import felics_b200
    with f:
        try:
            if args.sidecar:
                with open(args.sidecar, "rb") as sf:
                    img = felics_b200._default_codec().decompress_with_sidecar(f.read(), sf.read())
            else:
                img = felics_b200.decompress_image(f)
        except OSError as e:
            print(f"Cannot open side file: {e}")
            return 1
        except felics_b200.DecompressionError as e:
            print(f"Error while decompressing the image: {e.kind}")
            return 1
        except felics_b200.FelicsError as e:
            print(f"Error while decompressing the image: {e}")
            return 1
    import cv2
    try:
        ok = cv2.imwrite(args.output, img[..., ::-1] if img.ndim == 3 else img)
    except cv2.error as e:
        print(f"Cannot save image: {e}")
        return 1
    if not ok:
        print("Cannot save image: the image could not be written")
        return 1
    return 0


if __name__ == "__main__":
    tool = sys.argv[1] if len(sys.argv) > 1 else ""
    mains = {"cfelics": cfelics_main, "dfelics": dfelics_main}
    if tool not in mains:
        print("usage: python -m felics_b200.cli {cfelics|dfelics} -i INPUT -o OUTPUT")
        raise SystemExit(2)
    raise SystemExit(mains[tool](sys.argv[2:]))
