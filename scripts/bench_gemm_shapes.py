"""k_tc_gemm2 on the model's launch shapes (c5 batch, joint encoder): forward with its epilogues, dgrad plain / accumulating."""
import os, sys
sys.path.insert(0, os.getcwd())
sys.argv = ["bench_gemm.py", "none"]
exec(open("scripts/bench_gemm.py").read())
os.environ["RR_TC_FAKE_PRESPLIT"] = "1"


def dgrad_case(M, n, k, accumulate):
    dZ = torch.randn(M, n, device=dev)
    W = torch.randn(n, k, device=dev) * 0.05
    dX = torch.zeros(M, k, device=dev)
    scratch = torch.empty(int(L.rr_linear_dgrad_tc_scratch_bytes(n, k)), dtype=torch.uint8, device=dev)

    def run():
        _lib.check(L.rr_linear_dgrad_tc(M, n, k, dZ.data_ptr(), n, W.data_ptr(), k, dX.data_ptr(), k, accumulate, scratch.data_ptr(), scratch.numel(), st()))
    return run, 2.0 * M * n * k, (M * n + M * k * (2 if accumulate else 1)) * 4.0, (dZ, W, dX, scratch)


cases = [("W_h fwd [2B,304]x[304,304] bias+resid+relu+dropout", fwd_case(2 * B, 304, 304)),
         ("W_h fwd [B,304]x[304,304] bias+resid+relu+dropout", fwd_case(B, 304, 304)),
         ("W_i fwd [2B,88]x[88,304] plain", fwd_case(2 * B, 304, 88, epi=False)),
         ("W_o fwd [2A,64|304]x[.,304] two-source + relu", fwd_case(2 * A, 304, 64, 304)),
         ("plain [2B,304]x[304,304]", fwd_case(2 * B, 304, 304, epi=False)),
         ("h600 resid [A,608]x[608,608]", fwd_case(A, 608, 608)),
         ("dgrad bf16 accumulate [2B,304]x[304,304]", dgrad_case(2 * B, 304, 304, 1)),
         ("dgrad bf16 plain [2B,304]x[304,304]", dgrad_case(2 * B, 304, 304, 0))]
for rep in range(2):
    for name, c in cases:
        report(name, c[0], c[1], c[2])
