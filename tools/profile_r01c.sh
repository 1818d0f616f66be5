# End-of-round captures (profiles/r01c_*): launch lists of the default bench and of the 16-bit workload, full captures of
# the kernels that changed after r01b.  Each bench command first exits 0 without ncu.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-decode --no-verify"
$CMD > gpurun_out/pc_img.json 2> gpurun_out/pc_img.err || exit 1
FELICS_B200_NO_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01c_launches.csv $CMD > gpurun_out/ncuc1.log 2>&1
CMD16="python bench.py --workload gray16 --steps 2 --warmup 1 --no-decode --no-verify"
$CMD16 > gpurun_out/pc_16.json 2> gpurun_out/pc_16.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_gray16_launches.csv $CMD16 > gpurun_out/ncuc2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k16_bwalk|k16_scatter" -s 2 -c 2 -o gpurun_out/r01c_gray16 $CMD16 > gpurun_out/ncuc3.log 2>&1
FELICS_B200_NO_OVERLAP=1 ncu --set full --clock-control none -k regex:"k_hist4|k_scatter4|k_code4" -s 3 -c 3 -o gpurun_out/r01c_pixel $CMD > gpurun_out/ncuc4.log 2>&1
ls -la gpurun_out/r01c*.ncu-rep
