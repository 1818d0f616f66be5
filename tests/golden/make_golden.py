"""Generate the committed golden fixtures from the reference's image corpus.

Run HERE (the container that mounts /root/reference):  python tests/golden/make_golden.py
Outputs (committed):
  tests/golden/corpus_manifest.json  per-file: dims, sha256(pixels), oracle .fel size + sha256
  tests/golden/images.npz            a few small corpus images (pixels) so GPU-box tests
                                     have real-image inputs without /root/reference
  tests/golden/published_sizes.json  the size tables of DOC.md:385-396 and :469-477

The .fel digests are ORACLE outputs (regression anchors); the totals they sum to
are the reference's published numbers (DOC.md), which is what pins the oracle.
"""
import glob
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import felics_oracle as fo  # noqa: E402

REF = Path("/root/reference")
FOLDERS = ["image-suite/grayscale/8bit", "image-suite/grayscale/16bit", "image-suite/rgb/8bit", "bench/tiff_files"]
SMALL = {
    "image-suite/grayscale/8bit/5.1.09.tiff": "gray8_5.1.09",
    "image-suite/grayscale/8bit/boat.512.tiff": "gray8_boat.512",
    "image-suite/rgb/8bit/lena_color_256.tif": "rgb8_lena_color_256",
    "bench/tiff_files/pluto.tiff": "rgb8_pluto",
    # larger natural images: mirror-tiled in the GPU tests they give long chains with real statistics
    "image-suite/grayscale/8bit/5.3.01.tiff": "gray8_5.3.01",
    "image-suite/rgb/8bit/mandril_color.tif": "rgb8_mandril",
}


def main():
    manifest = []
    small = {}
    for folder in FOLDERS:
        for p in sorted(glob.glob(str(REF / folder / "*"))):
            a = np.array(Image.open(p))
            if a.ndim == 3 and a.shape[2] == 4:
                a = a[..., :3]
            a = np.ascontiguousarray(a)
            fel = fo.compress(a)
            assert np.array_equal(fo.decompress(fel), a), p
            rel = os.path.relpath(p, REF)
            manifest.append({
                "file": rel, "folder": folder, "height": int(a.shape[0]), "width": int(a.shape[1]),
                "channels": 1 if a.ndim == 2 else 3, "dtype": str(a.dtype),
                "pixels_sha256": hashlib.sha256(a.tobytes()).hexdigest(),
                "fel_bytes": len(fel), "fel_sha256": hashlib.sha256(fel).hexdigest(),
            })
            if rel in SMALL:
                small[SMALL[rel]] = a
    # a 16-bit crop keeps the fixture small
    a16 = np.array(Image.open(sorted(glob.glob(str(REF / "image-suite/grayscale/16bit/*")))[0]))
    small["gray16_crop"] = np.ascontiguousarray(a16[:160, :200])
    here = Path(__file__).resolve().parent
    (here / "corpus_manifest.json").write_text(json.dumps(manifest, indent=1))
    np.savez_compressed(here / "images.npz", **small)
    published = {
        "source": "DOC.md:385-396 (K-set tables), DOC.md:469-477 (RGB with/without transform)",
        "gray8_total_by_kset": {"0-5": 8529509, "0-6": 8530013, "0-4": 8529804, "0-3": 8531913, "0-2": 8563203, "1-5": 8602668, "2-5": 8748943},
        "gray16_total_by_kset": {"0-14": 7543288, "0-12": 7542507, "0-10": 7546086, "0-9": 7636332, "1-11": 7542209,
                                  "3-11": 7542196, "4-11": 7542120, "5-11": 7542011, "6-11": 7543104},
        "rgb8_with_transform": {"house.tiff": 105741, "peppers.tiff": 512290, "tree.tiff": 122246, "lena_color_256.tif": 110707,
                                "sailboat.tiff": 545539, "mandril_color.tif": 617524, "airplane.tiff": 385832},
        "rgb8_without_transform": {"house.tiff": 109047, "peppers.tiff": 504372, "tree.tiff": 130166, "lena_color_256.tif": 118186,
                                   "sailboat.tiff": 544998, "mandril_color.tif": 639373, "airplane.tiff": 413706},
    }
    (here / "published_sizes.json").write_text(json.dumps(published, indent=1))
    print(f"{len(manifest)} files; small fixtures: {list(small)}")


if __name__ == "__main__":
    main()
