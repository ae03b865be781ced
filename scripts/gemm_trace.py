"""Per-work-item timeline of k_tc_gemm2's first CTA (RR_TC_DIAG & 16): the W_h forward [B,304] x [304,304] with bias + residual + ReLU + dropout.
Slots per work item: 0 producer starts the item, 1 its last stage issued; 2 MMA warp has the accumulator, 3 last MMA committed;
epilogue warp 6: 8 first residual requested, 9 accumulator full, 10.. after its sub-blocks; epilogue warp 21: 13 accumulator full, 14 done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.argv = ["bench_gemm.py", "none"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_gemm.py")).read())
diag = int(os.environ.get("TRACE_DIAG", "0"))
os.environ["RR_TC_FAKE_PRESPLIT"] = "1"
os.environ["RR_TC_DIAG"] = str(16 | diag)
L.rr_reload_switches()
run, fl, by, keep = fwd_case(B, 304, 304)
for _ in range(3):
    run()
torch.cuda.synchronize()
S, W = 96, 16
t = np.zeros(S * W, dtype=np.uint64)
_lib.check(L.rr_debug_wgrad_trace(t.ctypes.data, S * W))
t = t.reshape(S, W).astype(np.int64)
t0 = t[0, 0]
names = {0: "P:start", 1: "P:issued", 2: "M:acc", 3: "M:commit", 4: "b0:tmem", 5: "b0:issue", 6: "b0:block", 7: "b0:nextR", 8: "E6:req", 9: "E6:full", 10: "E6:blk0", 11: "E6:blk1", 12: "E6:blk2", 13: "E21:full", 14: "E21:done"}
cols = sorted(names)
print(f"W_h fwd diag={diag}: clocks relative to the producer's first stamp; work items of CTA 0 (17 per CTA)")
print("item " + " ".join(f"{names[c]:>9s}" for c in cols))
for w in range(17):
    print(f"{w:4d} " + " ".join(f"{(t[w, c] - t0) if t[w, c] else 0:9d}" for c in cols))
