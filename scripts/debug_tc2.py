"""GPU debugging aid: determinism + accuracy of the tcgen05 GEMM on gradient-like data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reactranker_b200 import _lib
L = _lib.lib(); L.rr_set_gemm_mode(1)
S = torch.cuda.current_stream().cuda_stream
def run(M, n, k, X, W, reps=8):
    outs = []
    for _ in range(reps):
        Y = torch.full((M, n), float("nan"), device="cuda")
        _lib.check(L.rr_linear_fwd(M, n, X.data_ptr(), k, W.data_ptr(), k, None, 0, None, 0, None, None, 0, Y.data_ptr(), n, 0, 0.0, 0, 0, S))
        outs.append(Y)
    torch.cuda.synchronize()
    return outs
for (M, n, k, scale, sparsity) in [(1000, 304, 304, 1.0, 0.0), (1000, 304, 304, 1e-6, 0.5), (500, 304, 304, 1e-6, 0.5), (25, 304, 304, 1e-3, 0.5),
                                   (1000, 608, 608, 1e-6, 0.5), (20000, 304, 304, 1e-6, 0.5)]:
    g = torch.Generator().manual_seed(1)
    X = torch.randn(M, k, generator=g) * scale * (torch.rand(M, k, generator=g) >= sparsity)
    X[0] *= 3000.0            # a padding-row-like outlier
    W = torch.randn(n, k, generator=g) * 0.05
    want = X.double() @ W.double().T
    outs = run(M, n, k, X.cuda(), W.cuda())
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    err = (outs[0].double().cpu() - want).abs()
    rowmax = want.abs().max(dim=1, keepdim=True).values.clamp_min(1e-300)
    print(f"M{M} n{n} k{k} scale {scale} sparsity {sparsity}: deterministic={same}  max err/rowmax {float((err/rowmax).max()):.2e}  "
          f"err/globalmax {float(err.max()/want.abs().max()):.2e}  nan={bool(torch.isnan(outs[0]).any())}")
