"""GPU debugging aid: per-layer activations, forward-TC vs SIMT."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
from reactranker_b200.models.base_model import build_model
hidden, depth = int(sys.argv[1]), int(sys.argv[2])
sizes = [7, 5, 9, 4]
ds = synthetic.make_dataset(77, sizes)
torch.manual_seed(3)
model = build_model(hidden_size=hidden, mpnn_depth=depth, mpnn_diff_depth=depth, ffn_depth=3, use_bias=True, dropout=0.0, task_num=1,
                    ffn_last_layer="with_softplus", add_features_dim=1).cuda(0)
r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
hp = _lib.lib().rr_padded(hidden)
names = []
for k, g in ((0, r_g), (1, p_g)):
    names += [(f"enc{k}.inp", g.n_bonds)] + [(f"enc{k}.pre{t}", g.n_bonds) for t in range(depth - 1)] + [(f"enc{k}.m{t+1}", g.n_bonds) for t in range(depth - 1)]
    names += [(f"enc{k}.am", g.n_atoms), (f"enc{k}.hid", g.n_atoms)]
A = p_g.n_atoms
names += [("d", A), ("inp2", A)] + [(f"nm{t}", A) for t in range(depth - 1)] + [(f"m2_{t+1}", A) for t in range(depth - 1)] + [("am2", A), ("hid2", A)]
acts = {}
for mode in (0, 2):
    _lib.lib().rr_set_gemm_mode(mode)
    out = model(r_g, p_g, gpu=0, add_features=ds.temp.reshape(-1, 1))
    acts[mode] = {n: model.saved_activation(out, n, rows, hp).clone() for n, rows in names}
for n, rows in names:
    a, b = acts[0][n].double(), acts[2][n].double()
    diff = (a - b).abs()
    rowmax = a.abs().max(dim=1).values.clamp_min(1e-30)
    worst_row = int((diff.max(dim=1).values / rowmax).argmax())
    flips = int(((a != 0) != (b != 0)).sum())
    print(f"{n:12s} max|diff|/max {float(diff.max()/a.abs().max()):.2e}  worst row {worst_row} rel {float(diff[worst_row].max()/rowmax[worst_row]):.2e}  "
          f"zero-pattern flips {flips}  row0 max {float(a[0].abs().max()):.2e}")
