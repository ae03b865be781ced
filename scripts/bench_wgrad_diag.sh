#!/bin/bash
# wgrad3 with parts switched off (RR_TC_DIAG bits: 1 = no dZ conversion, 2 = no X conversion, 4 = no MMAs, 8 = no write-back):
# which of conversion traffic, MMA operand reads and the pipeline hand-offs paces a stage.  Output: gpurun_out/wgrad_diag.log
out=gpurun_out/wgrad_diag.log
: > $out
python - >> $out 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
sys.argv = ["bench_gemm.py", "none"]
exec(open("scripts/bench_gemm.py").read())
cases = [("wgrad [2B,304]^T[2B,304]", wgrad_case(2 * B, 304, 304)), ("wgrad [2B,304]^T[2B,88]", wgrad_case(2 * B, 304, 88))]
for diag in (0, 1, 2, 3, 4, 5, 6, 7):
    os.environ["RR_TC_DIAG"] = str(diag)
    for name, c in cases:
        report(f"diag={diag} {name}", c[0], c[1], c[2])
PY
cat $out
