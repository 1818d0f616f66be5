"""Host-side logic of the bench and its synthetic inputs (no GPU): the mirror tiling of configs[4], the corpus plan and its
sharding, the committed corpus against its manifest, and the traffic lookup of the roofline line."""
import json
import sys

import numpy as np

from conftest import GOLDEN, ROOT

sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from felics_b200 import sharding, synth  # noqa: E402
from oracle import felics_oracle as fo  # noqa: E402


def test_mirror_tile_reflects_without_seams():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (5, 7, 3), dtype=np.uint8)
    out = synth.mirror_tile(img, 23, 30)
    assert out.shape == (23, 30, 3) and out.flags["C_CONTIGUOUS"]
    assert np.array_equal(out[:5, :7], img)
    assert np.array_equal(out[5:10, :7], img[::-1])            # reflected below ...
    assert np.array_equal(out[:5, 7:14], img[:, ::-1])         # ... and to the right
    assert np.array_equal(out[10:15, 14:21], img)              # period 2H x 2W
    gray = rng.integers(0, 256, (4, 4), dtype=np.uint8)
    assert np.array_equal(synth.mirror_tile(gray, 4, 4), gray)


def test_tile_generator_is_deterministic_and_in_range():
    a = synth.tile_batch(3, first=65533)
    b = synth.tile_batch(1, first=65534)
    assert a.shape == (3, 512, 512) and a.dtype == np.uint8
    assert np.array_equal(a[1], b[0])
    assert 60 < a.mean() < 200 and a.std() > 5


def test_corpus_plan_covers_every_source_and_shards_add_up():
    total = 512 * 8
    plans = [bench.corpus_plan(*sharding.shard_range(total, r, 8)) for r in range(8)]
    for side in bench.CORPUS_SIDES:
        assert sum(len(p[side]) for p in plans) == total // 14 * 7 + sum(1 for g in range(total - total % 14, total) if (g % 14) // 7 == bench.CORPUS_SIDES.index(side))
    whole = bench.corpus_plan(0, total)
    for side in bench.CORPUS_SIDES:
        assert sum((p[side] for p in plans), []) == whole[side]
        assert set(whole[side]) == set(range(7))


def test_committed_corpus_matches_its_manifest():
    meta = json.loads((GOLDEN / "bench_corpus.json").read_text())
    images = dict(np.load(GOLDEN / "bench_corpus.npz"))
    assert sorted(images) == sorted(meta["images"]) and len(images) == 7
    assert meta["total_fel_bytes"] == 4984136                  # DOC.md's seven RGB files with the colour transform
    small = min(images, key=lambda n: images[n].size)
    assert list(images[small].shape) == meta["images"][small]["shape"]
    assert len(fo.compress(images[small])) == meta["images"][small]["fel_bytes"]


def test_recorded_traffic_scales_to_the_launch():
    per_launch, src = bench.recorded_traffic("stream", "tiles", units_per_launch=1000)
    whole, _ = bench.recorded_traffic("stream", "tiles")
    table = json.loads((ROOT / "profiles" / "traffic.json").read_text())["tiles"]
    assert abs(per_launch - 1000 * table["_per_unit"]["stream"]) < 1 and whole == table["stream"]
    assert "git" in src and "profiles/" in src
    assert bench.recorded_traffic("stream", "no such workload") == (None, None)
