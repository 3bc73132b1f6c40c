#!/bin/bash
# Round profile capture (run under gpurun): bench line, ncu launch list of the same command, one
# full capture of the heavy kernels.  Usage: scripts/capture_profiles.sh <tag>
set -u
tag=${1:-r1}
out=gpurun_out
python bench.py --steps 20 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense-reference --no-training-rows > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --import-source on --clock-control none \
    -k regex:"k_pfn_pad_tc|k_pfn_real|k_canvas|k_encode|k_iou_pass|k_emit_dense|k_pfn_stats_tc|k_mean|k_feat|k_rank|k_bn_finalize" \
    -s 60 -c 24 -o $out/${tag}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-training-rows > $out/${tag}_ncu_full.log 2>&1
ls -la $out | tail -8
