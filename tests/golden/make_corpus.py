"""Commit the seven RGB images of the reference's bench corpus (bench/tiff_files, BASELINE.json configs[4]) as pixels, so that
the GPU box -- which has no /root/reference -- can run the corpus workload and its parity test.

Run HERE (the container that mounts /root/reference):  python tests/golden/make_corpus.py
Output (committed): tests/golden/bench_corpus.npz  {name: HxWx3 uint8}, plus the oracle's .fel size of every image in
tests/golden/bench_corpus.json (the sizes sum to 4,984,136 bytes)."""
import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import felics_oracle as fo  # noqa: E402

SRC = Path("/root/reference/bench/tiff_files")


def main():
    images, sizes = {}, {}
    for path in sorted(SRC.glob("*.tiff")):
        img = np.array(Image.open(path).convert("RGB"))
        images[path.stem] = img
        sizes[path.stem] = {"shape": list(img.shape), "fel_bytes": len(fo.compress(img))}
    out = Path(__file__).resolve().parent
    np.savez_compressed(out / "bench_corpus.npz", **images)
    (out / "bench_corpus.json").write_text(json.dumps({"images": sizes, "total_fel_bytes": sum(v["fel_bytes"] for v in sizes.values())}, indent=1))
    print({k: v for k, v in sizes.items()}, sum(v["fel_bytes"] for v in sizes.values()))


if __name__ == "__main__":
    main()
