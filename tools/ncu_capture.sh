#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the hot kernels. Output -> gpurun_out/
# usage: tools/ncu_capture.sh <tag>
TAG=${1:-r01}
CMD="python bench.py --frames 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
}
cap scatter lz_scatter_k 6 2
cap group lz_group_apply_k 4 2
cap small lz_small_k 1 1
cap tiny lz_tiny_k 1 1
cap greduce lz_group_reduce_k 4 1
cap expand expand_mrr_k 0 1
cap quantize quantize_k 1 1
cap hist hist_vec4_k 0 1
cap recon reconstruct_k 0 1
cap pack lz_pack_k 0 1
cap classify classify_k 0 1
cap index "orbit_mark_k<33" 0 1
ls gpurun_out | grep $TAG | wc -l
