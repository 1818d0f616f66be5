#!/bin/bash
# quick check of the streaming encoder on the GPU box: its tests, then a short tile bench (kernel time and rate).  usage: quick_stream.sh [tiles]
TILES=${1:-4736}
timeout 600 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -3
python bench.py --tiles $TILES --steps 3 --warmup 2 --no-extras --no-decode --parity-tiles 64 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('GPixel/s %.2f  ms %.3f  stages %s  parity %s  redone %s' % (d['value']/1e3, d['ms_per_step'], d['stages_ms_per_step'], d['parity'], d['stream_redone']))"
FELICS_B200_STREAM_DBG=8 python bench.py --tiles $TILES --steps 2 --warmup 1 --no-extras --no-decode --no-verify 2>&1 | grep "stream phases"
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers --clock-control none -k regex:k_stream_encode -s 1 -c 1 python bench.py --tiles $TILES --steps 2 --warmup 1 --no-extras --no-decode --no-verify 2>&1 | grep -E "inst_executed|duration|issue_active|warps_active|occupancy_limit" | awk -v t=$TILES '{print $1, $2, $3; if ($1 ~ /inst_executed/) {gsub(",","",$3); printf("  warp-instr/px %.3f\n", $3/(t*262144))}}'
