#!/bin/bash
# Round profile capture on the GPU box (run under gpurun): plain bench first (must exit 0), then the ncu launch list of the
# same command, then --set full captures of the top kernels (a few launches each: every replay pass saves / restores the
# multi-GB workspace).  A second argument `gemm` stops after the GEMM capture.  Outputs land in gpurun_out/ (keep them under 64 MiB); scripts/summarise_profiles.py turns them into
# profiles/<round>_*.  Usage: scripts/capture_profiles.sh r01
set -u
R=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/${R}_bench_c5.json 2> $OUT/${R}_bench_c5.err || { echo "bench failed"; tail -5 $OUT/${R}_bench_c5.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${R}_launches_c5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/${R}_ncu_launches.log 2>&1
# second step of `--steps 1 --warmup 1`: 29 k_tc_gemm2 / 20 k_tc_wgrad2 launches per step
ncu --set full --clock-control none -k regex:k_tc_gemm2 --launch-skip 29 -c 4 -o $OUT/${R}_full_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${R}_ncu_full_gemm.log 2>&1
[ "${2:-all}" = "gemm" ] && { tail -n 1 $OUT/${R}_ncu_full_gemm.log; du -sh $OUT; exit 0; }   # partial re-capture: only the GEMM kernel changed
ncu --set full --clock-control none -k regex:k_tc_wgrad --launch-skip 31 -c 3 -o $OUT/${R}_full_wgrad \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${R}_ncu_full_wgrad.log 2>&1
ncu --set full --clock-control none -k regex:k_rowpipe --launch-skip 24 -c 10 -o $OUT/${R}_full_mp \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${R}_ncu_full_mp.log 2>&1
tail -n 1 $OUT/${R}_ncu_full_gemm.log $OUT/${R}_ncu_full_wgrad.log $OUT/${R}_ncu_full_mp.log
du -sh $OUT
