"""torchrun worker: one ListMLE step on 2 GPUs (groups sharded, global max_num_bonds, global normaliser, SUM all-reduce)
equals the same step on one GPU over the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reactranker_b200 import synthetic  # noqa: E402
from reactranker_b200.features.featurization import BatchMolGraph, DeviceGraph  # noqa: E402
from reactranker_b200.models.base_model import build_model  # noqa: E402
from reactranker_b200.parallel import GradSync, broadcast_parameters, shard_groups, shard_rows  # noqa: E402
from reactranker_b200.train.loss import MLEloss  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sizes = [9, 7, 11, 6, 8, 10]
ds = synthetic.make_dataset(5, sizes, star_leaves_in_group={1: 7})
torch.manual_seed(0)
model = build_model(hidden_size=300, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0).cuda(local)
broadcast_parameters(model)
r_all, p_all = [ds.mols[t] for t in ds.rsmi], [ds.mols[t] for t in ds.psmi]
r_g, p_g = BatchMolGraph(r_all), BatchMolGraph(p_all)
targets = torch.tensor(ds.lgk, dtype=torch.float32)
feats = ds.temp.reshape(-1, 1)
# single device, whole batch
out = model(r_g, p_g, gpu=local, add_features=feats)
MLEloss()(out, sizes, targets, local).backward()
want = [p.grad.clone() for p in model.parameters() if p.requires_grad]
model.zero_grad()
# sharded
atoms = [sum(ds.mols[ds.rsmi[i]].n_atoms for i in range(o, o + n)) for o, n in zip(np.cumsum([0] + sizes[:-1]), sizes)]
lo, hi = shard_groups(atoms, world)[rank]
a, b = shard_rows(sizes, lo, hi)
rs, ps = BatchMolGraph(r_all[a:b]), BatchMolGraph(p_all[a:b])
out = model(DeviceGraph.from_batches([rs], dev, [r_g.max_num_bonds]), DeviceGraph.from_batches([ps], dev, [p_g.max_num_bonds]), gpu=local,
            add_features=feats[a:b])
MLEloss(global_norm=len(sizes))(out, sizes[lo:hi], targets[a:b], local).backward()
GradSync(model.parameters())()
got = [p.grad for p in model.parameters() if p.requires_grad]
# same criterion as tests/helpers.grads_close: per-tensor error <= rtol * max|want_k| + 1e-2 * rtol * (largest gradient entry of the model);
# the second term absorbs tensors whose true gradient is zero (the last bias under the shift-invariant ListMLE: ~1e-8 of rounding noise)
gscale = max(float(w.abs().max()) for w in want)
rtol = 2e-4
worst = max(float((g - w).abs().max()) / (float(w.abs().max()) + 1e-2 * gscale) for g, w in zip(got, want))
assert worst < rtol, worst
dist.barrier()
if rank == 0:
    print("DP-EQUIVALENCE-OK worst rel err", worst)
dist.destroy_process_group()
