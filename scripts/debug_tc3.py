"""GPU debugging aid: isolate the tensor core's accumulation error (inputs exactly TF32-representable => lo terms vanish)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reactranker_b200 import _lib
L = _lib.lib()
S = torch.cuda.current_stream().cuda_stream
def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)
def gemm(mode, M, n, k, X, W):
    L.rr_set_gemm_mode(mode)
    Y = torch.full((M, n), float("nan"), device="cuda")
    _lib.check(L.rr_linear_fwd(M, n, X.data_ptr(), k, W.data_ptr(), k, None, 0, None, 0, None, None, 0, Y.data_ptr(), n, 0, 0.0, 0, 0, S))
    torch.cuda.synchronize()
    return Y
M, n = 512, 304
for k in (32, 96, 304, 608, 1216):
    for kind in ("tf32-exact inputs", "fp32 inputs", "positive tf32-exact"):
        g = torch.Generator().manual_seed(k)
        X, W = torch.randn(M, k, generator=g), torch.randn(n, k, generator=g)
        if kind == "positive tf32-exact":
            X, W = X.abs(), W.abs()
        if "tf32" in kind:
            X, W = tf32(X), tf32(W)
        want = X.double() @ W.double().T
        res = {}
        for mode, name in ((1, "tc"), (0, "simt")):
            Y = gemm(mode, M, n, k, X.cuda(), W.cuda()).double().cpu()
            e = (Y - want)
            res[name] = (float(e.abs().max() / want.abs().max()), float((e / want.abs().clamp_min(1e-300)).mean()) if kind.startswith("positive") else float(e.mean() / want.abs().mean()))
        print(f"K={k:5d} {kind:22s} tc: max {res['tc'][0]:.2e} bias {res['tc'][1]:+.2e} | simt: max {res['simt'][0]:.2e} bias {res['simt'][1]:+.2e}")
