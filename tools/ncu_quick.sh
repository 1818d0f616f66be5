#!/bin/bash
# quick counters of the streaming encoder (one launch): instructions, issue activity, duration, dram bytes, plus the
# source-level capture used by tools/sass_lines.py.  usage: ncu_quick.sh <tag> [tiles]
TAG=$1; TILES=${2:-2368}
CMD="python bench.py --tiles $TILES --steps 2 --warmup 1 --no-extras --no-decode --no-verify"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_stream_encode -s 1 -c 1 -o gpurun_out/${TAG}_stream $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_stream.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
for a,b in zip(h,v):
    if a in want: print(a,b)
"
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_plain.json').read().strip().splitlines()[-1]); print('plain', d['value'], d['ms_per_step'], d['stages_ms_per_step'])"
