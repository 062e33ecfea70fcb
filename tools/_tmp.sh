CMD="python bench.py --frames 500 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lzc_ -c 40 --csv --log-file gpurun_out/r02b_lzc_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lzc_level_k -s 0 -c 1 -f -o gpurun_out/r02b_lvl_L3 $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lzc_level_k -s 5 -c 1 -f -o gpurun_out/r02b_lvl_L8 $CMD > /dev/null 2>&1
ls -la gpurun_out | tail -4
