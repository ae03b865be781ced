#!/bin/bash
# Round profile capture on the GPU box (run under gpurun): plain bench first (must exit 0), then the ncu launch list of the
# same command, then --set full captures of the top kernels (every replay pass saves / restores the multi-GB workspace, so a
# handful of launches each).  Outputs land in gpurun_out/ (keep them under 64 MiB); scripts/summarise_profiles.py turns them into
# profiles/<round>_*.  Usage: scripts/capture_profiles.sh r02
# Launches of one training step (c5, joint encoder): 11 forward + 11 dgrad k_tc_gemm2, 15 k_tc_wgrad3 (11 x <32>, 4 x <64>), 13 k_rowpipe; the benchmark runs
# >= 3 warm-up steps first, so skipping three steps' worth of a kernel's launches lands in a steady-state step.
set -u
R=${1:-r02}
OUT=gpurun_out
B="python bench.py --no-cpu --no-others --sustain 0"
mkdir -p $OUT
$B --steps 10 --warmup 3 > $OUT/${R}_bench_c5.json 2> $OUT/${R}_bench_c5.err || { echo "bench failed"; tail -5 $OUT/${R}_bench_c5.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${R}_launches_c5.csv \
    $B --steps 2 --warmup 3 > $OUT/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm2 --launch-skip 66 -c 8 -o $OUT/${R}_full_gemm \
    $B --steps 1 --warmup 3 > $OUT/${R}_ncu_full_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm2 --launch-skip 78 -c 6 -o $OUT/${R}_full_dgrad \
    $B --steps 1 --warmup 3 > $OUT/${R}_ncu_full_dgrad.log 2>&1
[ "${2:-all}" = "gemm" ] && { tail -n 1 $OUT/${R}_ncu_full_gemm.log; du -sh $OUT; exit 0; }
ncu --set full --clock-control none --import-source on -k regex:k_tc_wgrad --launch-skip 45 -c 8 -o $OUT/${R}_full_wgrad \
    $B --steps 1 --warmup 3 > $OUT/${R}_ncu_full_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rowpipe --launch-skip 39 -c 13 -o $OUT/${R}_full_mp \
    $B --steps 1 --warmup 3 > $OUT/${R}_ncu_full_mp.log 2>&1
tail -n 1 $OUT/${R}_ncu_full_gemm.log $OUT/${R}_ncu_full_dgrad.log $OUT/${R}_ncu_full_wgrad.log $OUT/${R}_ncu_full_mp.log
du -sh $OUT
# NOTE: gpurun copies back at most 64 MiB: the four --import-source captures of a whole step's launches came to 66 MB once and were lost.
# Capture fewer launches per kernel (-c) or run the captures in two gpurun calls.
