set -x
CMD="python bench.py --steps 2 --warmup 1 --no-decode --no-verify"
$CMD > gpurun_out/p_img.json 2> gpurun_out/p_img.err || exit 1
FELICS_B200_NO_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
CMD16="python bench.py --workload gray16 --steps 2 --warmup 1 --no-decode --no-verify"
$CMD16 > gpurun_out/p_16.json 2> gpurun_out/p_16.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_gray16_launches.csv $CMD16 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k16_bwalk|k16_scatter|k16_pack|k16_code" -s 4 -c 4 -o gpurun_out/r01b_gray16 $CMD16 > gpurun_out/ncu3.log 2>&1
FELICS_B200_NO_OVERLAP=1 ncu --set full --clock-control none -k regex:"k_pack|k_prefix|k_kfill|k_scatter|k_code|k_hist" -s 6 -c 6 -o gpurun_out/r01b_pixel $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep
