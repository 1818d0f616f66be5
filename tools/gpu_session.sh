# one GPU session: smoke, tests, benches, the walker-per-SM experiment, then the two-warps-per-bucket patch rebuilt on the box
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
show='import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), round(d["value"]), round(d["e2e"]["value"]), d["parity"], {k:round(v,3) for k,v in d["stages_ms_per_step"].items() if v>0})'
python bench.py --workload gray16 --steps 5 --warmup 3 --no-decode 2>/dev/null | python -c "$show"
python tests/devtools/bw16_variants.py 0 2>&1 | tail -1
for p in 1 2 3; do FELICS_B200_WALK_PER_SM=$p python tests/devtools/natural.py 8 2>&1 | tail -1; done
for p in 2 3; do FELICS_B200_WALK_PER_SM=$p python bench.py --steps 3 --warmup 3 --no-decode 2>/dev/null | python -c "$show"; done
git apply tools/bw_ways2.patch 2>/dev/null || patch -p1 < tools/bw_ways2.patch
python __graft_entry__.py > /dev/null 2>&1
echo "--- two warps per bucket ---"
python -m pytest tests -m gpu -x -q -k 16 2>&1 | tail -1
python bench.py --workload gray16 --steps 5 --warmup 3 --no-decode 2>/dev/null | python -c "$show"
python tests/devtools/bw16_variants.py 0 2>&1 | tail -1
