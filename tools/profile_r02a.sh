#!/bin/bash
# round 2, first capture of the streaming band encoder (run under gpurun, one GPU)
CMD="python bench.py --tiles 2368 --steps 2 --warmup 1 --no-extras --no-decode --no-verify"
mkdir -p gpurun_out
$CMD > gpurun_out/r02a_plain.json 2> gpurun_out/r02a_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_launches.csv $CMD > gpurun_out/r02a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream_encode -s 1 -c 1 -o gpurun_out/r02a_stream $CMD > gpurun_out/r02a_ncu2.log 2>&1
