"""GPU micro-benchmark of the dense-layer kernels through the C ABI at the model's shapes (c5 batch: 160 889 bond rows, 82 001 atom rows).
RR_TC_DIAG / RR_TC_FAKE_PRESPLIT switch off parts of the tcgen05 kernel (timing experiments, results are then wrong)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reactranker_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
st = _lib.stream_ptr


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def fwd_case(M, n, k1, k2=0, epi=True, p=0.1):
    X1 = torch.relu(torch.randn(M, k1, device=dev))
    W1 = torch.randn(n, k1, device=dev) * 0.05
    X2 = torch.randn(M, k2, device=dev) if k2 else None
    W2 = torch.randn(n, k2, device=dev) * 0.05 if k2 else None
    bias = torch.randn(n, device=dev) if epi else None
    R = torch.randn(M, n, device=dev) if epi else None
    Y = torch.empty(M, n, device=dev)
    flags = (1 | (2 if p > 0 else 0)) if epi else 0

    def run():
        _lib.check(L.rr_linear_fwd(M, n, X1.data_ptr(), k1, W1.data_ptr(), k1, _lib.ptr(X2), k2, _lib.ptr(W2), k2, _lib.ptr(bias), _lib.ptr(R), n,
                                   Y.data_ptr(), n, flags, p, 1234, 7, st()))
    return run, 2.0 * M * n * (k1 + k2), (M * (k1 + k2) + M * n * (2 if epi else 1)) * 4.0, (X1, W1, X2, W2, bias, R, Y)


def wgrad_case(M, n, k):
    dZ = torch.randn(M, n, device=dev)
    X = torch.randn(M, k, device=dev)
    dW = torch.zeros(n, k, device=dev)
    db = torch.zeros(n, device=dev)

    def run():
        _lib.check(L.rr_linear_wgrad(M, n, k, dZ.data_ptr(), n, X.data_ptr(), k, dW.data_ptr(), k, db.data_ptr(), st()))
    return run, 2.0 * M * n * k, (M * (n + k)) * 4.0, (dZ, X, dW, db)


def report(name, run, flops, bytes_):
    L.rr_reload_switches()          # pick up whatever RR_* switch the caller just set or removed
    us = timeit(run)
    print(f"{name:58s} {us:8.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s  {bytes_ / us * 1e-3:7.0f} GB/s (algorithmic)", flush=True)


B, A = 160889, 82001
which = sys.argv[1] if len(sys.argv) > 1 else "all"
E = os.environ


def setenv(**kw):
    for k, v in kw.items():
        E[k] = str(v)
    L.rr_reload_switches()          # the library reads its RR_* switches once; re-read after every change


if which in ("all", "fwd"):
    cases = [("W_h fwd  [B,304]x[304,304] bias+resid+relu+dropout", fwd_case(B, 304, 304)),
             ("W_h fwd  same, no dropout", fwd_case(B, 304, 304, p=0.0)),
             ("dgrad-like [B,304]x[304,304] plain", fwd_case(B, 304, 304, epi=False)),
             ("W_i fwd  [B,88]x[88,304] plain", fwd_case(B, 304, 88, epi=False)),
             ("W_o fwd  [A,64|304]x[.,304] two-source + relu", fwd_case(A, 304, 64, 304)),
             ("h600: [A,608]x[608,608] plain", fwd_case(A, 608, 608, epi=False))]
    for promo in (128, 256):
        for ew in (4, 8):
            for fake in ("", "1"):
                setenv(RR_TMA_PROMO=promo, RR_TC_EW=ew, RR_TC_DIAG=0)
                if fake:
                    E["RR_TC_FAKE_PRESPLIT"] = "1"
                else:
                    E.pop("RR_TC_FAKE_PRESPLIT", None)
                for name, (run, fl, by, _) in cases:
                    report(f"promo={promo} ew={ew} presplit={fake or 0} {name}", run, fl, by)
        E["RR_TC_FAKE_PRESPLIT"] = "1"
        for diag in (8, 15):
            setenv(RR_TC_DIAG=diag)
            for name, (run, fl, by, _) in cases[:1]:
                report(f"promo={promo} ew=8 presplit=1 diag={diag} {name}", run, fl, by)
    E.pop("RR_TC_FAKE_PRESPLIT", None)
    setenv(RR_TC_DIAG=0, RR_TMA_PROMO=128, RR_TC_EW=8)
if which in ("all", "wgrad"):
    cases = [("wgrad [B,304]^T[B,304]", wgrad_case(B, 304, 304)), ("wgrad [B,304]^T[B,88]", wgrad_case(B, 304, 88)),
             ("wgrad [A,304]^T[A,304]", wgrad_case(A, 304, 304)), ("wgrad h600 [A,608]^T[A,608]", wgrad_case(A, 608, 608))]
    for tf32 in (1, 0):
        if tf32:
            E["RR_WG_TF32"] = "1"
        else:
            E.pop("RR_WG_TF32", None)
        setenv(RR_TC_DIAG=0)
        for name, c in cases:
            report(f"{'3xtf32' if tf32 else '3xbf16'} {name}", c[0], c[1], c[2])
        for diag in (4, 3, 7, 8):
            setenv(RR_TC_DIAG=diag)
            report(f"{'3xtf32' if tf32 else '3xbf16'} diag={diag} {cases[0][0]}", cases[0][1][0], cases[0][1][1], cases[0][1][2])
    setenv(RR_TC_DIAG=0)
if which == "diag":    # what bounds the forward kernel: switch its parts off one at a time (pre-split weights as in the model)
    E["RR_TC_FAKE_PRESPLIT"] = "1"
    cases = [("W_h fwd bias+resid+relu+dropout", fwd_case(B, 304, 304)), ("W_h fwd bias+resid+relu", fwd_case(B, 304, 304, p=0.0)),
             ("plain [B,304]x[304,304]", fwd_case(B, 304, 304, epi=False))]
    names = {0: "full", 1: "no A split", 4: "no MMA", 8: "no epilogue traffic", 5: "no A split, no MMA", 12: "no MMA, no epilogue", 9: "no A split, no epilogue",
             13: "TMA ring only"}
    for name, (run, fl, by, _) in cases:
        for diag, what in names.items():
            setenv(RR_TC_DIAG=diag)
            report(f"{name}: {what}", run, fl, by)
    setenv(RR_TC_DIAG=0)
    E.pop("RR_TC_FAKE_PRESPLIT", None)
if which == "one":     # one launch of each headline kernel, for ncu
    run, *_ = fwd_case(B, 304, 304)
    run2, *_ = wgrad_case(B, 304, 304)
    for _ in range(2):
        run()
        run2()
    torch.cuda.synchronize()
