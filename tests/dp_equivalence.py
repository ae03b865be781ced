"""torchrun worker (2+ GPUs): the PRODUCT's data-parallel step -- ``TrainStep.prepare`` / ``run`` as ``train()`` and ``bench.py`` drive it:
global batch plan, shard of whole groups, global ``max_num_bonds``, global normalisers, in-place all-reduce of the flat gradient buffer --
gives the gradients of the same step on one GPU over the whole batch, for a group-normalised loss (ListMLE), an item-normalised one
(ListNet) and a composite of both kinds (mle_gaussian)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reactranker_b200 import synthetic  # noqa: E402
from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features  # noqa: E402
from reactranker_b200.models.base_model import build_model  # noqa: E402
from reactranker_b200.train.step import TrainStep  # noqa: E402
from reactranker_b200.train.train_listwise import batch_loss  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
sizes = [9, 7, 11, 6, 8, 10, 5, 12]
ds = synthetic.make_dataset(5, sizes, star_leaves_in_group={1: 7})       # the star molecule sets the global max_num_bonds
fz = Parsing_features(ds.mols)
planner = DataProcessor(ds.to_dataframe())
worst_all = 0.0
for task, task_num in (("mle", 1), ("listnet", 1), ("mle_gaussian", 2)):
    torch.manual_seed(0)
    model = build_model(hidden_size=300, task_num=task_num, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0).cuda(local)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    step = TrainStep(model, opt, sched, task, local)
    assert (step.rank, step.world) == (rank, world)
    batch = next(iter(planner.generate_batch_reactions(smiles_list=["rsmi_mapped", "psmi_mapped"], target_name="lgk", batch_size=sum(sizes),
                                                       seed=3, add_features_name="temp")))
    reactions, targets, scope, feats = batch
    # single device, whole batch, through the same public calls
    r_g, p_g = fz.parsing_reactions(reactions)
    out = model(r_g, p_g, gpu=local, add_features=feats)
    whole = batch_loss(task, out, scope, torch.FloatTensor(targets).squeeze(), local)
    whole.backward()
    want = [p.grad.clone() for p in model.hot_parameters()]
    model.zero_grad()
    # the product's data-parallel step, through both host paths: the gathered-SMILES tuples of generate_batch_reactions (any featuriser
    # with the reference's interface) and the planner's row positions (what train() and bench.py feed it: per-frame id vectors + rr_batch_build)
    rows_scope = next(iter(planner.plan_batch_reactions(batch_size=sum(sizes), seed=3)))
    assert list(rows_scope[1]) == list(scope)
    steps_before = 0
    for how in ("smiles", "rows"):
        if how == "smiles":
            prepared = step.prepare(batch, fz)
        else:
            prepared = step.prepare_rows(planner, rows_scope[0], rows_scope[1], fz, ["rsmi_mapped", "psmi_mapped"], "lgk", "temp")
        assert prepared.groups == len(scope) and prepared.items == sum(scope) and 0 < prepared.rows < sum(scope)
        term = step.run(prepared)
        steps_before += 1
        assert step.sync.fast_path_steps == steps_before and step.sync.copy_path_steps == 0   # p.grad aliases the flat buffer: no staging copies
        total = step.global_loss(term)
        assert abs(total - float(whole.detach().reshape(-1)[0])) <= 1e-5 * abs(float(whole.detach().reshape(-1)[0])), (task, how, total, float(whole))
        got = [p.grad for p in model.hot_parameters()]
        # same criterion as tests/helpers.grads_close
        gscale = max(float(w.abs().max()) for w in want)
        worst = max(float((g - w).abs().max()) / (float(w.abs().max()) + 1e-2 * gscale) for g, w in zip(got, want))
        assert worst < 1e-3, (task, how, worst)
        worst_all = max(worst_all, worst)
# fewer groups than ranks (an epoch's tail batch): the ranks without a group take part in the all-reduce with zero gradients, through the
# staging bucket, while rank 0 all-reduces its aliased flat buffer in place -- one collective, same layout on every rank
torch.manual_seed(0)
model = build_model(hidden_size=64, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0).cuda(local)
opt = torch.optim.SGD(model.parameters(), lr=0.0)
step = TrainStep(model, opt, torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0), "mle", local)
rows, scope = next(iter(planner.plan_batch_reactions(batch_size=5, seed=1)))           # one (truncated) group
assert len(scope) == 1
smiles, tg, feats = planner._gather(planner.df, rows, ["rsmi_mapped", "psmi_mapped"], "lgk", "temp")
r_g, p_g = fz.parsing_reactions(smiles)
whole = batch_loss("mle", model(r_g, p_g, gpu=local, add_features=feats), scope, torch.FloatTensor(tg.reshape(-1, 1)).squeeze(), local)
whole.backward()
want = [p.grad.clone() for p in model.hot_parameters()]
model.zero_grad()
prepared = step.prepare_rows(planner, rows, scope, fz, ["rsmi_mapped", "psmi_mapped"], "lgk", "temp")
mine = torch.tensor([1.0 if prepared.rows > 0 else 0.0], device=f"cuda:{local}")
dist.all_reduce(mine)
assert float(mine) == 1.0                                                               # exactly one rank owns the group
step.run(prepared)
assert (step.sync.fast_path_steps, step.sync.copy_path_steps) == ((1, 0) if prepared.rows > 0 else (0, 1))
gscale = max(float(w.abs().max()) for w in want)
worst = max(float((p.grad - w).abs().max()) / (float(w.abs().max()) + 1e-2 * gscale) for p, w in zip(model.hot_parameters(), want))
assert worst < 1e-3, ("empty shard", worst)
dist.barrier()
if rank == 0:
    print("DP-EQUIVALENCE-OK worst rel err", worst_all)
dist.destroy_process_group()
