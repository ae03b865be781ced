"""Per-stage timeline of k_tc_wgrad3's first CTA (RR_TC_DIAG & 16): where a stage's clocks go.
Slots: 0/1 producer after raw_empty wait / after issuing the TMA boxes; 2/3 MMA warp after `ready` wait / after commit;
4..8 worker warp 2 (dZ + X): after raw_full wait, after mma_done wait, after dZ -> TMEM, after X -> bf16, after fences + arrive;
10..14 worker warp 9 (X only), the same points.  Usage: python scripts/wgrad_trace.py [n k [diag]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reactranker_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 304
k = int(sys.argv[2]) if len(sys.argv) > 2 else 304
diag = int(sys.argv[3]) if len(sys.argv) > 3 else 0
M = 2 * 160889
os.environ["RR_TC_DIAG"] = str(16 | diag)
L = _lib.lib()
L.rr_reload_switches()
dev = torch.device("cuda:0")
dZ, X = torch.randn(M, n, device=dev), torch.randn(M, k, device=dev)
dW, db = torch.zeros(n, k, device=dev), torch.zeros(n, device=dev)
for _ in range(3):
    _lib.check(L.rr_linear_wgrad(M, n, k, dZ.data_ptr(), n, X.data_ptr(), k, dW.data_ptr(), k, db.data_ptr(), _lib.stream_ptr()))
torch.cuda.synchronize()
S, W = 96, 16
t = np.zeros(S * W, dtype=np.uint64)
_lib.check(L.rr_debug_wgrad_trace(t.ctypes.data, S * W))
t = t.reshape(S, W).astype(np.int64)
t0 = t[0, 0]
names = {0: "P:empty", 1: "P:issued", 2: "M:ready", 3: "M:commit", 4: "A:full", 5: "A:mmadone", 6: "A:dZ", 7: "A:X", 8: "A:arrive",
         10: "X:full", 11: "X:mmadone", 12: "X:dZskip", 13: "X:X", 14: "X:arrive"}
cols = sorted(names)
print(f"wgrad [{M},{n}]^T[{M},{k}] diag={diag}: clocks relative to the producer's first stamp")
print("stage " + " ".join(f"{names[c]:>10s}" for c in cols))
for it in list(range(0, 6)) + list(range(40, 52)):
    print(f"{it:5d} " + " ".join(f"{t[it, c] - t0:10d}" for c in cols))
lo, hi = 24, 88
per = (t[hi, 8] - t[lo, 8]) / (hi - lo)
print(f"steady state: {per:.0f} clk per stage (stages {lo}..{hi})")
d = lambda a, b: float(np.mean(t[lo:hi, a] - t[lo:hi, b]))
print(f"  worker A: wait raw_full {d(4, 8) + per:.0f} (from its previous arrive)  wait mma_done {d(5, 4):.0f}  dZ->TMEM {d(6, 5):.0f}  X->bf16 {d(7, 6):.0f}  fences+arrive {d(8, 7):.0f}")
print(f"  worker X: wait raw_full {d(10, 14) + per:.0f}  wait mma_done {d(11, 10):.0f}  X->bf16 {d(13, 12):.0f}  fences+arrive {d(14, 13):.0f}")
print(f"  MMA warp: ready -> commit issued {d(3, 2):.0f};  worker arrive -> MMA sees ready {d(2, 8):.0f}")
print(f"  producer: issue takes {d(1, 0):.0f};  TMA issue -> worker sees raw_full {d(4, 1):.0f} (same stage)")
R = 2
print(f"  raw slot turnaround: worker arrive(it) -> producer wakes for it+{R}: {float(np.mean(t[lo + R:hi + R, 0] - t[lo:hi, 8])):.0f}")
