#!/bin/bash
# wgrad3 ring configurations ("rows per stage, raw slots, bf16 slots") on the model's shapes; one process per configuration
# (the switch is read once per process).  Output: gpurun_out/wgrad_cfg.log
out=gpurun_out/wgrad_cfg.log
: > $out
for cfg in ${WG3_CFGS:-32,2,2 32,3,1 16,4,2 16,5,3 16,6,2 16,5,2}; do
  echo "== RR_WG3_CFG=$cfg" >> $out
  RR_WG3_CFG=$cfg python - >> $out 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
sys.argv = ["bench_gemm.py", "none"]
exec(open("scripts/bench_gemm.py").read())
for name, c in [("wgrad [2B,304]^T[2B,304]", wgrad_case(2 * B, 304, 304)), ("wgrad [2B,304]^T[2B,88]", wgrad_case(2 * B, 304, 88)),
                ("wgrad [2A,304]^T[2A,304]", wgrad_case(2 * A, 304, 304)), ("wgrad [2A,304]^T[2A,64]", wgrad_case(2 * A, 304, 64)),
                ("wgrad [A,304]^T[A,304]", wgrad_case(A, 304, 304)), ("wgrad h600 [A,608]^T[A,608]", wgrad_case(A, 608, 608))]:
    report(name, c[0], c[1], c[2])
PY
done
cat $out
