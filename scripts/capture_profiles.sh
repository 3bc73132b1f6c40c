#!/bin/bash
# Round profile capture (run under gpurun, ONE ncu use per call: launch list, or full capture).
#   scripts/capture_profiles.sh <tag> bench|launches|full
# bench:    python bench.py (default config, all sub-objects)         -> gpurun_out/<tag>_bench.json
# launches: ncu launch list of a short bench (cold-cache, serialised) -> gpurun_out/<tag>_ncu_launches.csv
# full:     ncu --set full of the hot-path kernels                    -> gpurun_out/<tag>_full.ncu-rep
set -u
tag=${1:-r2}
what=${2:-bench}
out=gpurun_out
mkdir -p $out
short="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense-reference --no-training-rows --no-gpu-comparator --min-timed-s 0.01"
case $what in
  bench)
    python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1 ;;
  launches)
    $short > $out/${tag}_plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_ncu_launches.csv \
        $short > $out/${tag}_ncu_launches.log 2>&1 ;;
  full)
    python scripts/run_fused_once.py 3 > $out/${tag}_plain.log 2>&1 &&
    ncu --set full --import-source on --clock-control none \
        -k regex:"k_pfn_pad_tc|k_pfn_real|k_canvas|k_bn_finalize|k_mean|k_feat|k_rank|k_bin|k_assign|k_scatter" \
        -s 22 -c 11 -o $out/${tag}_full -f python scripts/run_fused_once.py 3 > $out/${tag}_ncu_full.log 2>&1 ;;
esac
ls -la $out | tail -5
