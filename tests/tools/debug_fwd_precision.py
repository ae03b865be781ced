"""Scores / loss / gradient error of the forward operand splits (3 x tf32, 3 x bf16) against the fp64 oracle, at the two north-star
model sizes, on a batch large enough for accumulation effects (usage: python tests/tools/debug_fwd_precision.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import reactranker_oracle as O
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train.train_listwise import batch_loss
L = _lib.lib()
for task, tn, tt, hidden, depth in (("mle", 1, None, 300, 3), ("evidential_ranking", 2, "evidential_ranking", 600, 5)):
    sizes = [9, 7, 11, 6, 8, 10, 12, 5]
    ds = synthetic.make_dataset(5, sizes, star_leaves_in_group={1: 7})
    torch.manual_seed(0)
    model = build_model(hidden_size=hidden, mpnn_depth=depth, mpnn_diff_depth=depth, task_num=tn, task_type=tt, ffn_last_layer="with_softplus",
                        add_features_dim=1, dropout=0.0).cuda(0)
    sd64 = {k: v.double().cpu() for k, v in model.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd64.items() if "cached_zero" not in k}
    full = dict(sd64)
    full.update(params)
    r_o, p_o = O.OracleBatch([ds.mols[t] for t in ds.rsmi]), O.OracleBatch([ds.mols[t] for t in ds.psmi])
    want = O.model_forward(full, r_o, p_o, ds.temp.reshape(-1, 1), mpnn_depth=depth, mpnn_diff_depth=depth, head=O.resolve_task_type(tn, "with_softplus", tt))
    t64 = torch.tensor(ds.lgk.astype(np.float32)).double()
    wl = O.loss_for_task(task, want, sizes, t64)
    wl.backward()
    r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
    targets = torch.tensor(ds.lgk, dtype=torch.float32)
    for name, fb in (("3xtf32", 0), ("3xbf16", 1)):
        L.rr_set_forward_bf16(fb)
        model.zero_grad()
        out = model(r_g, p_g, gpu=0, add_features=ds.temp.reshape(-1, 1))
        loss = batch_loss(task, out, sizes, targets, 0)
        loss.backward()
        o = out.detach().double().cpu()
        es = float((o - want.detach()).abs().max() / want.detach().abs().max())
        el = float((loss.detach().double().cpu().reshape(-1)[0] - wl.detach().reshape(-1)[0]).abs() / wl.detach().abs().reshape(-1)[0])
        gscale = max(float(p.grad.abs().max()) for p in params.values())
        eg = max(float((m.grad.double().cpu() - params[k].grad).abs().max()) / max(float(params[k].grad.abs().max()), 1e-2 * gscale)
                 for k, m in model.named_parameters() if m.requires_grad)
        print(f"{task} h{hidden} d{depth} forward {name}: scores {es:.2e}  loss {el:.2e}  worst gradient tensor {eg:.2e}")
L.rr_set_forward_bf16(0)
