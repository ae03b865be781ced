"""Per-tensor gradient error of the tensor-core backward splits (3 x bf16, 3 x tf32) against the exact-fp32 SIMT GEMM path, h = 300."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train.loss import MLEloss
L = _lib.lib()
sizes = [9, 7, 11, 6, 8, 10]
ds = synthetic.make_dataset(5, sizes, star_leaves_in_group={1: 7})
torch.manual_seed(0)
model = build_model(hidden_size=300, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0).cuda(0)
r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
targets = torch.tensor(ds.lgk, dtype=torch.float32)
feats = ds.temp.reshape(-1, 1)
res = {}
for name, mode, bf in (("simt", 0, 0), ("tf32", 1, 0), ("bf16", 1, 1)):
    L.rr_set_gemm_mode(mode)
    L.rr_set_backward_bf16(bf)
    model.zero_grad()
    out = model(r_g, p_g, gpu=0, add_features=feats)
    MLEloss()(out, sizes, targets, 0).backward()
    res[name] = {k: p.grad.double().cpu() for k, p in model.named_parameters() if p.requires_grad}
gscale = max(float(v.abs().max()) for v in res["simt"].values())
for k, w in res["simt"].items():
    den = max(float(w.abs().max()), 1e-3 * gscale)
    print(f"{k:28s} max|g| {float(w.abs().max()):.3e}  tf32 err {float((res['tf32'][k] - w).abs().max()) / den:.2e}  bf16 err {float((res['bf16'][k] - w).abs().max()) / den:.2e}")
