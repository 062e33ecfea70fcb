#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of every kernel class that is >= 2 % of the step, at the
# bench configuration (2000 source frames 1080p). Two ncu runs (encode kernels of the first batch, decode kernels); the
# reports are digested on the box (tools/ncu_digest.py) and deleted - only text comes back in gpurun_out/.
# usage: tools/ncu_capture.sh <tag>
TAG=${1:-r02}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ENC='regex:hist_vec4_k|lut_fill_k|quantize8_k|classify_k|emit_k|lzc_hashlink_k|lzc_link3_k|lzc_level_k|lz_emit_mark_k|lz_pack15_k|orbit_mark_k<15|orbit_spec_k<15'
DEC='regex:expand_mrr_k|index_steps_k|orbit_mark_k<33|orbit_spec_k<33|reconstruct_k|recon_p_k'
ncu --set full --clock-control none --import-source on -k "$ENC" -c 24 -f -o /tmp/prof_enc $CMD > gpurun_out/ncu_enc_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "$DEC" -c 6 -f -o /tmp/prof_dec $CMD > gpurun_out/ncu_dec_$TAG.log 2>&1
python tools/ncu_digest.py --top 8 /tmp/prof_enc.ncu-rep /tmp/prof_dec.ncu-rep > gpurun_out/ncu_digest_$TAG.txt 2>&1
ls -la /tmp/prof_enc.ncu-rep /tmp/prof_dec.ncu-rep; wc -l gpurun_out/ncu_digest_$TAG.txt
