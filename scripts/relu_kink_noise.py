#!/usr/bin/env python
"""How much gradient noise do ReLU-kink mask flips cause?  CPU experiments on the oracle / the reference's own arithmetic (no GPU).

    python scripts/relu_kink_noise.py floor 82 50        # the reference's fp32 arithmetic vs fp64 at the c5 batch (~5 min)
    python scripts/relu_kink_noise.py law 24 50 3e-6,1e-6,1e-7     # fp64 oracle with its linear layers perturbed by eps (~6 min)
    python scripts/relu_kink_noise.py which 24 50 1e-7   # which pre-activations flip under a 1e-7 perturbation

Results of the runs made in the build container are kept in profiles/r02_relu_kink_noise.md; tests/test_gpu_model.py
(test_gradients_at_bench_scale_vs_fp64_oracle) and DESIGN.md section 2 quote them.
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import reactranker_oracle as O  # noqa: E402
from reactranker_b200 import synthetic  # noqa: E402


def setup(groups, n):
    sizes = [n] * groups
    ds = synthetic.make_dataset(4242, sizes)
    sd = O.init_state_dict(300, 1, 1, True, seed=11)
    return sizes, ds, sd, O.OracleBatch([ds.mols[t] for t in ds.rsmi]), O.OracleBatch([ds.mols[t] for t in ds.psmi])


def grads(sd, r_o, p_o, ds, sizes, dt, lin=None, relu=None):
    s = {k: v.to(dt) for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in s.items() if "cached_zero" not in k}
    full = dict(s)
    full.update(params)
    orig_lin, orig_relu = O._lin, torch.relu
    if lin is not None:
        O._lin = lambda x, sd_, name: lin(orig_lin, x, sd_, name)
    if relu is not None:
        torch.relu = lambda x: relu(orig_relu, x)
    try:
        out = O.model_forward(full, r_o, p_o, ds.temp.reshape(-1, 1))
        loss = O.loss_for_task("mle", out, sizes, torch.tensor(ds.lgk.astype(np.float32)).to(dt))
        loss.backward(torch.ones_like(loss))
    finally:
        O._lin, torch.relu = orig_lin, orig_relu
    return out.detach().double().numpy(), {k: v.grad.double().numpy() for k, v in params.items()}


def report(tag, g1, g0):
    gs = max(np.abs(v).max() for v in g0.values())
    live = [k for k in g0 if np.abs(g0[k]).max() >= 1e-6 * gs]
    mr = max(np.abs(g1[k] - g0[k]).max() / np.abs(g0[k]).max() for k in live)
    l2 = max(np.linalg.norm(g1[k] - g0[k]) / np.linalg.norm(g0[k]) for k in live)
    print(f"{tag}: worst max-rel {mr:.2e}  worst rel-L2 {l2:.2e}", flush=True)


def main():
    mode, groups, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    sizes, ds, sd, r_o, p_o = setup(groups, n)
    t0 = time.time()
    _, g64 = grads(sd, r_o, p_o, ds, sizes, torch.float64)
    print(f"fp64 oracle, {groups * n} reactions: {time.time() - t0:.0f} s", flush=True)
    if mode == "floor":
        _, g32 = grads(sd, r_o, p_o, ds, sizes, torch.float32)
        report(f"PyTorch fp32 vs fp64 (the reference's own arithmetic), {groups * n} reactions", g32, g64)
        return
    noises = [float(x) for x in sys.argv[4].split(",")]
    for eps in noises:
        for kind in ("gaussian", "toward-zero"):
            gen = torch.Generator().manual_seed(1)

            def lin(orig, x, sd_, name, eps=eps, kind=kind, gen=gen):
                out = orig(x, sd_, name)
                if name.startswith("ffn"):
                    return out
                g = out.detach() - sd_[name + ".bias"]
                if kind == "gaussian":      # unbiased rounding noise relative to the row's largest entry
                    return out + torch.randn(out.shape, generator=gen, dtype=out.dtype) * eps * g.abs().amax(-1, keepdim=True)
                return out - g * eps * (0.5 + torch.rand(out.shape, generator=gen, dtype=out.dtype))   # RZ accumulation: the product shrinks
            if mode == "law":
                _, g1 = grads(sd, r_o, p_o, ds, sizes, torch.float64, lin=lin)
                report(f"{groups * n} reactions, linear layers perturbed by {eps:g} ({kind})", g1, g64)
            elif mode == "which" and kind == "toward-zero":
                pre = {}

                def relu(orig, x, store):
                    store.append(x.detach().clone())
                    return orig(x)
                a, b = [], []
                grads(sd, r_o, p_o, ds, sizes, torch.float64, relu=lambda o, x: relu(o, x, a))
                grads(sd, r_o, p_o, ds, sizes, torch.float64, lin=lin, relu=lambda o, x: relu(o, x, b))
                total = 0
                for i, (u, v) in enumerate(zip(a, b)):
                    for r, c in ((u > 0) != (v > 0)).nonzero().tolist():
                        print(f"relu #{i} {tuple(u.shape)}: row {r} col {c} pre-activation {u[r, c].item():.3e} (row max {u[r].abs().max().item():.2f})")
                        total += 1
                print(f"{total} mask flips among {sum(x.numel() for x in a)} pre-activations at eps {eps:g}")
                _ = pre


if __name__ == "__main__":
    main()
