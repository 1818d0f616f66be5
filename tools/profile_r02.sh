#!/bin/bash
# round 2 captures (run under gpurun, one GPU): the default bench line, its ncu launch list, and one `--set full` capture each of
# the streaming band encoder and of the gray batch decoder.  Everything lands in gpurun_out/; profiles/ gets the CSV extracts.
TAG=${1:-r02}
CMD="python bench.py"
SHORT="python bench.py --steps 2 --warmup 3"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}_default.err || { tail -5 gpurun_out/${TAG}_default.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stream_encode -s 2 -c 1 -o gpurun_out/${TAG}_stream $SHORT --no-extras --no-decode --no-verify > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_g8 -c 1 -o gpurun_out/${TAG}_decode $SHORT --no-extras --no-verify > gpurun_out/${TAG}_ncu3.log 2>&1
for k in stream decode; do
  ncu -i gpurun_out/${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_${k}_full_raw.csv 2>/dev/null
done
tail -c 400 gpurun_out/${TAG}_ncu1.log; ls -la gpurun_out/${TAG}_*
