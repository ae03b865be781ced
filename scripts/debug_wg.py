import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reactranker_b200 import _lib
L = _lib.lib(); L.rr_set_gemm_mode(1)
S = torch.cuda.current_stream().cuda_stream
for (M, n, k) in [(256, 128, 32), (1000, 304, 304)]:
    g = torch.Generator().manual_seed(1)
    dZ, X = torch.randn(M, n, generator=g), torch.randn(M, k, generator=g)
    dW = torch.zeros(n, k, device="cuda"); db = torch.zeros(n, device="cuda")
    dZd, Xd = dZ.cuda(), X.cuda()
    print("py ptrs dW %x db %x" % (dW.data_ptr(), db.data_ptr()))
    st = L.rr_linear_wgrad(M, n, k, dZd.data_ptr(), n, Xd.data_ptr(), k, dW.data_ptr(), k, db.data_ptr(), S)
    torch.cuda.synchronize()
    print("db[:4]", db[:4].tolist(), "want", dZ.double().sum(0)[:4].tolist())
    want = dZ.double().T @ X.double()
    got = dW.double().cpu()
    print(M, n, k, "status", st, "dW nonzero frac", float((got != 0).float().mean()), "max got", float(got.abs().max()), "max want", float(want.abs().max()),
          "db err", float((db.double().cpu() - dZ.double().sum(0)).abs().max()))
    if float(got.abs().max()) > 0:
        # where does it match?
        err = (got - want).abs() / want.abs().max()
        print("   err by 32-col block:", [f"{float(err[:, j:j+32].max()):.1e}" for j in range(0, k, 32)])
        print("   err by 32-row block:", [f"{float(err[j:j+32].max()):.1e}" for j in range(0, n, 32)])
