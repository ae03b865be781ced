"""GPU debugging aid: run-to-run determinism of the model in forward-TC mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train import loss as RL
hidden, depth = int(sys.argv[1]), int(sys.argv[2])
sizes = [7, 5, 9, 4]
ds = synthetic.make_dataset(77, sizes)
torch.manual_seed(3)
model = build_model(hidden_size=hidden, mpnn_depth=depth, mpnn_diff_depth=depth, ffn_depth=3, use_bias=True, dropout=0.0, task_num=1,
                    ffn_last_layer="with_softplus", add_features_dim=1).cuda(0)
targets = torch.tensor(ds.lgk.astype(np.float32))
r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
for mode in (0, 2):
    _lib.lib().rr_set_gemm_mode(mode)
    runs = []
    for rep in range(4):
        model.zero_grad()
        out = model(r_g, p_g, gpu=0, add_features=ds.temp.reshape(-1, 1))
        loss = RL.MLEloss()(out, sizes, targets, 0); loss.backward()
        runs.append((out.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.requires_grad}))
    same_out = all(torch.equal(runs[0][0], r[0]) for r in runs[1:])
    worst = max(float((runs[0][1][k] - r[1][k]).abs().max() / runs[0][1][k].abs().max().clamp_min(1e-30)) for r in runs[1:] for k in runs[0][1] if k != "ffn.ffn.7.bias")
    print(f"mode {mode}: scores bitwise equal across runs: {same_out}; worst run-to-run grad rel diff {worst:.2e}")
    if mode == 0:
        base = runs[0]
    else:
        for k in base[1]:
            e = float((runs[0][1][k] - base[1][k]).abs().max() / base[1][k].abs().max().clamp_min(1e-30))
            print(f"    {k:28s} tc-fwd vs simt rel diff {e:.2e}")
        # reactant duplicates: identical rows in, identical rows out?
