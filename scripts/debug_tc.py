"""GPU debugging aid: per-tensor gradient error of the SIMT and tcgen05 paths against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import reactranker_oracle as O
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train import loss as RL

hidden, depth = int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 3
sizes = [7, 5, 9, 4]
ds = synthetic.make_dataset(77, sizes)
torch.manual_seed(3)
model = build_model(hidden_size=hidden, mpnn_depth=depth, mpnn_diff_depth=depth, ffn_depth=3, use_bias=True, dropout=0.0, task_num=1,
                    ffn_last_layer="with_softplus", add_features_dim=1).cuda(0)
sd64 = {k: v.double().cpu() for k, v in model.state_dict().items()}
params = {k: v.clone().requires_grad_(True) for k, v in sd64.items() if "cached_zero" not in k}
full = dict(sd64); full.update(params)
want = O.model_forward(full, O.OracleBatch([ds.mols[t] for t in ds.rsmi]), O.OracleBatch([ds.mols[t] for t in ds.psmi]), ds.temp.reshape(-1, 1),
                       mpnn_depth=depth, mpnn_diff_depth=depth, head="with_softplus")
targets = torch.tensor(ds.lgk.astype(np.float32))
wl = O.listmle_loss(want, sizes, targets.double()); wl.backward()
r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
for mode in (0, 1, 2, 3):
    _lib.lib().rr_set_gemm_mode(mode)
    model.zero_grad()
    out = model(r_g, p_g, gpu=0, add_features=ds.temp.reshape(-1, 1))
    loss = RL.MLEloss()(out, sizes, targets, 0); loss.backward()
    es = float((out.detach().cpu().double() - want.detach()).abs().max() / want.detach().abs().max())
    print(f"mode {mode}: scores rel err {es:.2e} loss rel err {abs(float(loss)-float(wl))/abs(float(wl)):.2e}")
    gmax = max(float(v.grad.abs().max()) for v in params.values())
    worst = max((float((p.grad.double().cpu() - params[k].grad).abs().max()) / max(float(params[k].grad.abs().max()), 1e-30), k)
                for k, p in model.named_parameters() if p.requires_grad and k != "ffn.ffn.7.bias")
    print(f"   worst per-tensor rel err {worst[0]:.2e} ({worst[1]})")
