"""GPU parity of the whole path through the public (reference-shaped) API against
(1) golden vectors produced by the REAL reference (tests/golden/model.npz, fp32 and fp64 runs) and
(2) the CPU oracle on fresh inputs.

Tolerances (north star: 1e-4 relative on losses, 1e-3 on gradients): the fp32 SIMT path is held to
2e-5 on scores/losses and 2e-4 on gradients against the reference's fp64 run -- the reference's own
fp32 run differs from its fp64 run by ~1e-6 / 1e-5 on the same cases.
"""
import numpy as np
import pytest
import torch

from oracle import reactranker_oracle as O
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph, DeviceGraph
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train import loss as RL
from helpers import (TOL, check_grads, dataset_from_golden, forced_relu_masks, gpu_relu_masks, grads_close, rel_err, sd_from_golden,
                     star_dict)

pytestmark = pytest.mark.gpu
GPU = 0
TASKS = {"mle": (1, None), "listnet": (1, None), "evidential_ranking": (2, "evidential_ranking"),
         "gauss_regression": (2, None), "regression": (1, None),
         "mle_gaussian": (2, None), "listnet_gauss": (2, None), "mle_regression": (1, None), "listnet_regression": (1, None),
         "regression_exploss": (1, None), "mledis_gaussian": (2, None), "listnetdis_gauss": (2, None),
        "listnet_uq": (1, "listnet"), "listnetdis_lognorm": (2, "listnetdis_lognorm"), "dirichlet_uq": (1, "listnet", "with_uncertainty"),
        "evidential": (4, None), "mle_evidential": (4, None), "mledis_evidential": (4, None), "listnet_evidential": (4, None)}


def product_loss(task, out, scope, targets, gpu=GPU):
    from reactranker_b200.train.train_listwise import batch_loss
    if task in ("listnet_uq", "dirichlet_uq"):   # the golden's point on the annealing schedule (tests/golden/make_golden.py)
        return batch_loss(task, out, scope, targets, gpu, 0.05, 3, 5)
    return batch_loss(task, out, scope, targets, gpu)


def make_model(hidden, task, depth, ddepth, sd=None, dropout=0.0, last="with_softplus"):
    tn, tt, *ll = TASKS[task]
    last = ll[0] if ll else last
    m = build_model(hidden_size=hidden, mpnn_depth=depth, mpnn_diff_depth=ddepth, ffn_depth=3, use_bias=True, dropout=dropout,
                    task_num=tn, ffn_last_layer=last, task_type=tt, add_features_dim=1)
    if sd is not None:
        m.load_state_dict(sd)
    return m.cuda(GPU)


def oracle_run(sd, ds, sizes, task, depth, ddepth, masks=None, last="with_softplus", feats=None):
    """fp64 oracle forward + loss + backward on the data set of a test; with ``masks`` the ReLU decisions are the GPU's
    (helpers.forced_relu_masks).  Returns (scores, loss, {name: gradient}, mask flips)."""
    sd64 = {k: v.double().cpu() for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd64.items() if "cached_zero" not in k}
    full = dict(sd64)
    full.update(params)
    tn, tt, *ll = TASKS[task]
    head = O.resolve_task_type(tn, ll[0] if ll else last, tt)
    r_o, p_o = O.OracleBatch([ds.mols[t] for t in ds.rsmi]), O.OracleBatch([ds.mols[t] for t in ds.psmi])
    feats = ds.temp.reshape(-1, 1) if feats is None else feats
    targets = torch.tensor(ds.lgk.astype(np.float32)).double()
    flips = 0
    if masks is None:
        want = O.model_forward(full, r_o, p_o, feats, mpnn_depth=depth, mpnn_diff_depth=ddepth, head=head)
    else:
        with forced_relu_masks(masks) as fm:
            want = O.model_forward(full, r_o, p_o, feats, mpnn_depth=depth, mpnn_diff_depth=ddepth, head=head)
        flips = fm.flips
    wl = O.loss_for_task(task, want, sizes, targets)
    wl.backward(torch.ones_like(wl))
    return want.detach().numpy(), wl.detach().numpy(), {k: v.grad.numpy() for k, v in params.items()}, flips


def masked_oracle_factory(model, out, r_g, p_g, sd, ds, sizes, task, hidden, depth, ddepth, **kw):
    """Reads the GPU's ReLU masks NOW (before backward recycles the workspace); the returned callable re-runs the oracle with them."""
    masks = gpu_relu_masks(model, out, (r_g.n_atoms, r_g.n_bonds), (p_g.n_atoms, p_g.n_bonds), hidden, depth, ddepth)

    def run():
        _, _, g, flips = oracle_run(sd, ds, sizes, task, depth, ddepth, masks=masks, **kw)
        return g, flips
    return run


CASES = ["mle.h40", "listnet.h40", "evidential_ranking.h40", "gauss_regression.h40", "regression.h40", "mle.star.h40",
         "evidential_ranking.h24d5"]
COMPOSITE = ["mle_gaussian.h40", "listnet_gauss.h40", "mle_regression.h40", "listnet_regression.h40", "regression_exploss.h40",
             "mledis_gaussian.h40", "listnetdis_gauss.h40", "listnet_uq.h40", "listnetdis_lognorm.h40", "dirichlet_uq.h40", "evidential.h40",
             "mle_evidential.h40", "mledis_evidential.h40", "listnet_evidential.h40"]


@pytest.mark.parametrize("name", CASES + COMPOSITE)
def test_scores_loss_grads_vs_reference_golden(golden, name, gemm_mode):
    """All 21 golden cases of the real reference, on the exact-fp32 SIMT GEMMs and on the product's tcgen05 path."""
    g = golden("model_composite" if name in COMPOSITE else "model")
    task = str(g[name + ".task"])
    ds, sizes, hidden, depth, ddepth = dataset_from_golden(g, name)
    sd = sd_from_golden(g, name + ".sd")
    model = make_model(hidden, task, depth, ddepth, sd)
    model.train()
    r_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi])
    p_g = BatchMolGraph([ds.mols[t] for t in ds.psmi])
    out = model(r_g, p_g, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    targets = torch.FloatTensor(ds.lgk.reshape(-1, 1)).squeeze()          # train_listwise.py:187
    loss = product_loss(task, out, sizes, targets)
    redo = masked_oracle_factory(model, out, r_g, p_g, sd, ds, sizes, task, hidden, depth, ddepth) if gemm_mode == 1 else None
    model.zero_grad()
    loss.backward()
    tol = TOL[gemm_mode]
    assert tuple(out.shape) == tuple(g[name + ".f32.scores"].shape)
    assert tuple(loss.shape) == tuple(g[name + ".f32.loss"].shape)       # [1] for mle/evidential, 0-d otherwise
    assert rel_err(out.detach().cpu().numpy(), g[name + ".f64.scores"]) < tol["score"]
    assert rel_err(loss.detach().cpu().numpy(), g[name + ".f64.loss"]) < tol["loss"]
    got = {k: p.grad.cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}
    want = {k: g[f"{name}.f64.grad.{k}"] for k in got}
    check_grads(gemm_mode, got, want, redo)
    assert model.encoder.cached_zero_vector.grad is None


def test_h300_against_reference_golden(golden, gemm_mode):
    """hidden 300 (the north-star width): weights re-created from the torch seed (checked bit-exact on CPU in
    test_host_cpu), outputs/loss/large-gradient checksums from the reference run."""
    g = golden("model")
    name = "mle.h300"
    ds, sizes, hidden, depth, ddepth = dataset_from_golden(g, name)
    torch.manual_seed(int(g[name + ".meta"][1]))
    model = make_model(hidden, "mle", depth, ddepth)
    r_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi])
    p_g = BatchMolGraph([ds.mols[t] for t in ds.psmi])
    out = model(r_g, p_g, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    loss = RL.MLEloss()(out, sizes, torch.FloatTensor(ds.lgk), GPU)
    loss.backward()
    tol = TOL[gemm_mode]
    assert rel_err(out.detach().cpu().numpy(), g[name + ".f64.scores"]) < tol["score"]
    assert rel_err(loss.detach().cpu().numpy(), g[name + ".f64.loss"]) < tol["loss"]
    gscale = max(float(np.abs(g[k]).max()) for k in g.files if k.startswith(name + ".f64.grad."))
    gt = tol["grad"]
    for k, p in model.named_parameters():
        if not p.requires_grad:
            continue
        v = p.grad.double().cpu().numpy()
        if f"{name}.f64.grad.{k}" in g.files:
            w = g[f"{name}.f64.grad.{k}"]
            assert np.abs(v - w).max() <= gt * np.abs(w).max() + 0.5 * gt * gscale, k
        else:
            s = g[f"{name}.f64.gradsum.{k}"]
            got = np.asarray([v.sum(), np.abs(v).sum(), (v ** 2).sum()])
            assert abs(got[1] - s[1]) <= gt * s[1] and abs(got[2] - s[2]) <= 2 * gt * s[2], k


@pytest.mark.parametrize("algo,pre", [("sum_session", ""), ("accelerate_grad", "acc.")])
def test_ranknet_window_vs_reference_golden(golden, algo, pre, gemm_mode):
    """One accumulation window of factorized_training_loop (train_pairwise.py:81-160), both training_algo values: every group keeps its
    OWN max_num_bonds (one reference forward per group) but all groups share one launch here."""
    g = golden("ranknet")
    sizes = [int(x) for x in g["sizes"]]
    ds = synthetic.make_dataset(int(g["seed"]), sizes, star_leaves_in_group=star_dict(g["star"]))
    model = build_model(hidden_size=40, mpnn_depth=3, mpnn_diff_depth=3, ffn_depth=3, use_bias=True, dropout=0.0, task_num=1,
                        ffn_last_layer="no_softplus", add_features_dim=1)
    model.load_state_dict(sd_from_golden(g, "sd"))
    model = model.cuda(GPU)
    dev = torch.device("cuda", GPU)
    r_batches, p_batches, o = [], [], 0
    for n in sizes:
        r_batches.append(BatchMolGraph([ds.mols[t] for t in ds.rsmi[o:o + n]]))
        p_batches.append(BatchMolGraph([ds.mols[t] for t in ds.psmi[o:o + n]]))
        o += n
    rg, pg = DeviceGraph.from_batches(r_batches, dev), DeviceGraph.from_batches(p_batches, dev)
    y = model(rg, pg, gpu=GPU, add_features=ds.lgk.reshape(-1, 1))       # the target-column leak (load_reactions.py:264-267)
    pairs = sum(RL.count_ordered_pairs(ds.lgk[a:a + n]) for a, n in zip(np.cumsum([0] + sizes[:-1]), sizes))
    assert pairs == float(g["f64.pairs"])
    loss = RL.ranknet_window_loss(y, sizes, ds.lgk.astype(np.float32), pairs, sigma=1.0, gpu=GPU, training_algo=algo)
    loss.backward()
    tol = TOL[gemm_mode]
    assert rel_err(y.detach().cpu().numpy(), g["f64.scores"]) < tol["score"]
    assert rel_err(loss.detach().cpu().numpy(), g[pre + "f64.loss"]) < tol["loss"]
    got = {k: p.grad.cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}
    assert not grads_close(got, {k: g[f"{pre}f64.grad.{k}"] for k in got}, tol["grad"])


def test_three_optimizer_steps_vs_reference_golden(golden, gemm_mode):
    """Forward, ListMLE, backward, Adam, NoamLR for three steps (train_listwise.py:177-290)."""
    from reactranker_b200.train.utils import build_lr_scheduler, build_optimizer
    g = golden("steps")
    ds = synthetic.make_dataset(51, [6, 5, 4, 6, 3, 6])
    model = make_model(40, "mle", 3, 3, sd_from_golden(g, "sd0"))
    opt = build_optimizer(model)
    sched = build_lr_scheduler(opt, warmup_epochs=2, total_epochs=4, train_data_size=30, batch_size=10, init_lr=1e-4, max_lr=1e-3, final_lr=1e-4)
    model.train()
    losses = []
    for lo, hi, scope in [(0, 11, [6, 5]), (11, 21, [4, 6]), (21, 30, [3, 6])]:
        r_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi[lo:hi]])
        p_g = BatchMolGraph([ds.mols[t] for t in ds.psmi[lo:hi]])
        out = model(r_g, p_g, gpu=GPU, add_features=ds.temp[lo:hi].reshape(-1, 1))
        loss = RL.MLEloss()(out, scope, torch.FloatTensor(ds.lgk[lo:hi]), GPU)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss))
    assert np.allclose(losses, g["losses"], rtol=1e-4)      # steps 2 and 3 see the updated weights
    # Adam turns the SIGN of a near-zero gradient into a full +-lr move, so rounding noise on degenerate entries (e.g. the
    # last bias under a shift-invariant loss) is amplified: the reference's own fp32 and fp64 runs differ by 1e-2 relative
    # there.  Weights are therefore held to: typical entry within 1e-6, no entry further than Adam's bound 2*sum(lr).
    lr_sum = float(np.sum(g["lrs"][:3]))
    # ffn.ffn.4.bias / ffn.ffn.7.bias have an exactly-zero true gradient here (units are on or off for every row of a group and
    # ListMLE is shift invariant: 1e-18 in the fp64 oracle), so their trajectory is pure rounding noise times Adam.
    for k, v in model.state_dict().items():
        d = np.abs(v.cpu().numpy() - g["sd3." + k])
        assert float(d.max()) <= 2 * lr_sum, k
        if k not in ("ffn.ffn.4.bias", "ffn.ffn.7.bias"):
            assert float(np.median(d)) <= 1e-6, k


@pytest.mark.parametrize("task,hidden,depth", [("mle", 300, 3), ("evidential_ranking", 600, 5), ("mle", 296, 3)])
def test_vs_oracle_fresh_inputs(task, hidden, depth, gemm_mode):
    """The two north-star widths (and one that is not a multiple of 16) on inputs the goldens do not cover, oracle run in fp64 on the
    host, in both GEMM modes.  Mode 1 (the product default) is held to the north star's 1e-4 on scores / loss and 1e-3 on every
    gradient tensor; see helpers.check_grads for the one accepted exception (a mask decided by a pre-activation within 2e-5 of zero)."""
    sizes = [7, 5, 9, 4]
    ds = synthetic.make_dataset(77, sizes)
    torch.manual_seed(3)
    model = make_model(hidden, task, depth, depth)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ws, wl, wg, _ = oracle_run(sd, ds, sizes, task, depth, depth)
    r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
    out = model(r_g, p_g, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    loss = product_loss(task, out, sizes, torch.tensor(ds.lgk.astype(np.float32)))
    redo = masked_oracle_factory(model, out, r_g, p_g, sd, ds, sizes, task, hidden, depth, depth) if gemm_mode == 1 else None
    loss.backward()
    tol = TOL[gemm_mode]
    assert rel_err(out.detach().cpu().numpy(), ws) < tol["score"]
    assert rel_err(loss.detach().cpu().numpy(), wl) < tol["loss"]
    got = {k: p.grad.cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}
    check_grads(gemm_mode, got, {k: wg[k] for k in got}, redo)


# Gradient tolerance on a FULL-SIZE batch.  A ReLU whose pre-activation lies within the forward pass's rounding error of zero may get the
# other mask than exact arithmetic gives it, and each such flip adds or removes ONE row's contribution to a weight gradient.  A gradient row is
# a sum over N rows with cancellation (~ sqrt(N) x a typical term), so a single flip moves it by ~ 1 / sqrt(N): 0.7 % at 23 k atom rows, 0.35 %
# at 82 k.  With forward error eps there are ~ eps * N flips per output unit, so the relative gradient noise is ~ sqrt(eps) whatever the batch
# size.  NO fp32 implementation escapes this: the reference's own fp32 run differs from its fp64 run by up to 1.55e-3 of a tensor's maximum
# on the c5 batch (scripts/relu_kink_noise.py; rel-L2 4e-4), and perturbing the fp64 oracle's linear layers by 1e-7 flips 4 of 1e8
# pre-activations at 1200 reactions and moves one gradient tensor by 1.2e-2 of its maximum.  Hence two checks per GEMM mode:
#   (1) unconditional bounds, stated per mode (measured: SIMT fp32 4e-4 / 1e-4, tcgen05 3 x TF32 1.2e-2 / 3.6e-3 max-rel / rel-L2 at h300);
#   (2) the exact statement: GIVEN the GPU's own ReLU decisions -- every one of which differs from the fp64 oracle's only where the
#       pre-activation is within 2e-5 (relative to its row) of zero, asserted -- all gradients equal the fp64 oracle's to 2e-4.
BENCH_TOL = {0: dict(max_rel=4e-3, l2=2e-3), 1: dict(max_rel=2.5e-2, l2=6e-3)}


@pytest.mark.parametrize("task,hidden,depth,groups,n", [("mle", 300, 3, 24, 50), ("evidential_ranking", 600, 5, 12, 32)])
def test_gradients_at_bench_scale_vs_fp64_oracle(task, hidden, depth, groups, n, gemm_mode):
    """c5-shaped (ListMLE h300 d3, 50 candidates per group) and c4-shaped (UC-Listwise h600 d5, 32 per group) batches of ~24 k / ~8 k atom
    rows per graph -- large enough that no single row dominates a gradient, small enough for the fp64 CPU oracle -- at dropout 0, in both
    GEMM modes.  Scores and loss to the north star's 1e-4; gradients as described above.  The measured figures go to
    gpurun_out/parity_bench_scale.json when that directory exists."""
    import json
    import os
    sizes = [n] * groups
    ds = synthetic.make_dataset(4242, sizes)
    torch.manual_seed(11)
    model = make_model(hidden, task, depth, depth)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ws, wl, wg, _ = oracle_run(sd, ds, sizes, task, depth, depth)
    r_g, p_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi])
    out = model(r_g, p_g, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    loss = product_loss(task, out, sizes, torch.tensor(ds.lgk.astype(np.float32)))
    redo = masked_oracle_factory(model, out, r_g, p_g, sd, ds, sizes, task, hidden, depth, depth)
    loss.backward()
    e_s, e_l = rel_err(out.detach().cpu().numpy(), ws), rel_err(loss.detach().cpu().numpy(), wl)
    gscale = max(float(np.abs(v).max()) for v in wg.values())
    got = {k: p.grad.double().cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}
    live = [k for k in got if float(np.abs(wg[k]).max()) >= 1e-6 * gscale]
    rec = {k: (float(np.abs(got[k] - wg[k]).max() / np.abs(wg[k]).max()), float(np.linalg.norm(got[k] - wg[k]) / np.linalg.norm(wg[k]))) for k in live}
    worst = (max(v[0] for v in rec.values()), max(v[1] for v in rec.values()))
    # (2) conditional on the GPU's masks (forced_relu_masks asserts that each differing mask sits within helpers.KINK of a kink)
    wg2, flips = redo()
    n_pre = sum(int(m.numel()) for m in gpu_relu_masks_count(model, r_g, p_g, hidden, depth, out.shape[0]))
    cond = {k: float(np.abs(got[k] - wg2[k]).max() / np.abs(wg2[k]).max()) for k in live}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_bench_scale.json"), "a") as f:
            f.write(json.dumps({"task": task, "hidden": hidden, "depth": depth, "reactions": groups * n, "gemm_mode": gemm_mode, "score_rel": e_s,
                                "loss_rel": e_l, "grad_max_rel_worst": worst[0], "grad_rel_l2_worst": worst[1], "relu_mask_flips": flips,
                                "pre_activations": n_pre, "grad_max_rel_worst_given_gpu_masks": max(cond.values()), "per_tensor": rec}) + "\n")
    assert e_s < 1e-4 and e_l < 1e-4, (e_s, e_l)
    tol = BENCH_TOL[gemm_mode]
    assert worst[0] <= tol["max_rel"] and worst[1] <= tol["l2"], (worst, rec)
    assert flips <= 2e-4 * n_pre, (flips, n_pre)
    assert not grads_close(got, {k: wg2[k] for k in live}, 2e-4), cond


def gpu_relu_masks_count(model, r_g, p_g, hidden, depth, n_mols):
    """Shapes of the ReLU inputs of one forward (for the flip-rate denominator)."""
    shapes = []
    for g in (r_g, p_g):
        shapes += [(g.n_bonds, hidden)] * depth + [(g.n_atoms, hidden)]
    shapes += [(p_g.n_atoms, hidden)] * (depth + 1) + [(n_mols, hidden)] * 2
    return [torch.empty(s, device="meta") for s in shapes]


def test_eval_mode_is_deterministic_and_dropout_is_unbiased():
    ds = synthetic.make_dataset(5, [16] * 8)
    torch.manual_seed(0)
    model = make_model(300, "mle", 3, 3, dropout=0.1)
    r_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi])
    p_g = BatchMolGraph([ds.mols[t] for t in ds.psmi])
    feats = ds.temp.reshape(-1, 1)
    model.eval()
    with torch.no_grad():
        a = model(r_g, p_g, gpu=GPU, add_features=feats)
        b = model(r_g, p_g, gpu=GPU, add_features=feats)
    assert torch.equal(a, b)
    model.train()
    with torch.no_grad():
        runs = torch.stack([model(r_g, p_g, gpu=GPU, add_features=feats) for _ in range(64)])
    assert not torch.equal(runs[0], runs[1])
    # inverted dropout keeps the first moment: the mean over masks approaches the eval output
    assert float((runs.mean(0) - a).abs().mean()) < 0.35 * float(runs.std(0).mean())


def test_full_batch_properties_at_config2_size():
    """Config-2 scale (4096 reactions = 128 groups x 32, hidden 300): size-independent checks --
    (a) a data-parallel shard packed with the global max_num_bonds reproduces its rows of the global batch,
    (b) scores are invariant to the order of the groups in the batch, (c) everything is finite."""
    G, n = 128, 32
    ds = synthetic.make_dataset(123, [n] * G)
    torch.manual_seed(1)
    model = make_model(300, "listnet", 3, 3).eval()
    feats = ds.temp.reshape(-1, 1)
    r_all = [ds.mols[t] for t in ds.rsmi]
    p_all = [ds.mols[t] for t in ds.psmi]
    dev = torch.device("cuda", GPU)
    with torch.no_grad():
        r_g, p_g = BatchMolGraph(r_all), BatchMolGraph(p_all)
        full = model(r_g, p_g, gpu=GPU, add_features=feats)
        assert bool(torch.isfinite(full).all())
        half = G // 2 * n
        rs, ps = BatchMolGraph(r_all[:half]), BatchMolGraph(p_all[:half])
        shard = model(DeviceGraph.from_batches([rs], dev, [r_g.max_num_bonds]), DeviceGraph.from_batches([ps], dev, [p_g.max_num_bonds]),
                      gpu=GPU, add_features=feats[:half])
        assert rel_err(shard.cpu().numpy(), full[:half].cpu().numpy()) < 1e-6
        perm = np.concatenate([np.arange(half, G * n), np.arange(half)])
        swapped = model(BatchMolGraph([r_all[i] for i in perm]), BatchMolGraph([p_all[i] for i in perm]), gpu=GPU, add_features=feats[perm])
        assert rel_err(swapped.cpu().numpy(), full.cpu().numpy()[perm]) < 1e-5
    model.train()
    out = model(r_g, p_g, gpu=GPU, add_features=feats)
    loss = RL.ListnetLoss()(out, [n] * G, torch.FloatTensor(ds.lgk), GPU)
    loss.backward()
    assert all(bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.requires_grad)
    assert float(loss) > 0


@pytest.mark.parametrize("task,hidden,depth", [("mle", 300, 3), ("evidential_ranking", 600, 5)])
def test_tcgen05_backward_with_exact_forward(tc_mode, task, hidden, depth):
    """gemm mode 3 = exact-fp32 forward (the oracle's masks up to fp32 rounding) + tensor-core dgrad / wgrad (3 x bf16): isolates the
    backward GEMMs' accuracy from the forward's mask decisions.  Held to the SIMT path's 2e-4."""
    sizes = [7, 5, 9, 4]
    ds = synthetic.make_dataset(77, sizes)
    torch.manual_seed(3)
    model = make_model(hidden, task, depth, depth)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    _, _, wg, _ = oracle_run(sd, ds, sizes, task, depth, depth)
    _lib.check(_lib.lib().rr_set_gemm_mode(3))
    out = model(BatchMolGraph([ds.mols[t] for t in ds.rsmi]), BatchMolGraph([ds.mols[t] for t in ds.psmi]), gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    product_loss(task, out, sizes, torch.tensor(ds.lgk.astype(np.float32))).backward()
    got = {k: p.grad.double().cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}
    gscale = max(float(np.abs(v).max()) for v in wg.values())
    live = {k: w for k, w in wg.items() if float(np.abs(w).max()) >= 1e-6 * gscale}   # drop structurally-zero gradients
    assert not grads_close(got, live, 2e-4)


def test_padding_rows_are_exact_on_the_tensor_core_path(tc_mode):
    """After every tensor-core forward GEMM the segment padding rows are recomputed from the unsplit fp32 weights with fp64 accumulation
    (rr_model.cu: k_pad_rows_linear): one such row collects the gradient of every padded neighbour slot of its segment, so its ReLU masks
    must be the reference's.  Check: padding rows of every saved activation agree with the fp64 oracle to fp32 rounding (5e-7 of the row's
    maximum; the tensor-core rows carry 2-5e-6), also with several segments in one launch."""
    sizes = [6, 4, 5]
    ds = synthetic.make_dataset(31, sizes, star_leaves_in_group={1: 7})
    torch.manual_seed(4)
    model = make_model(300, "mle", 3, 3).eval()
    sd64 = {k: v.double().cpu() for k, v in model.state_dict().items()}
    dev = torch.device("cuda", GPU)
    hp = _lib.lib().rr_padded(300)
    o, rb, pb = 0, [], []
    for n in sizes:
        rb.append(BatchMolGraph([ds.mols[t] for t in ds.rsmi[o:o + n]]))
        pb.append(BatchMolGraph([ds.mols[t] for t in ds.psmi[o:o + n]]))
        o += n
    rg, pg = DeviceGraph.from_batches(rb, dev), DeviceGraph.from_batches(pb, dev)
    for p_ in model.parameters():
        p_.requires_grad_(True)
    out = model(rg, pg, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    a0 = b0 = 0
    for s_, n in enumerate(sizes):              # one oracle forward per segment = the reference's per-group forward
        lo = sum(sizes[:s_])
        trace = {}
        O.model_forward(sd64, O.OracleBatch([ds.mols[t] for t in ds.rsmi[lo:lo + n]]), O.OracleBatch([ds.mols[t] for t in ds.psmi[lo:lo + n]]),
                        ds.temp[lo:lo + n].reshape(-1, 1), trace=trace)
        for name, rows, row0, key in (("enc1.m1", pg.n_bonds, b0, "encoder.msg1"), ("enc1.m2", pg.n_bonds, b0, "encoder.msg2"),
                                      ("enc1.hid", pg.n_atoms, a0, "p_hiddens"), ("hid2", pg.n_atoms, a0, "diff_encoder.atom_hiddens")):
            got = model.saved_activation(out, name, rows, hp)[row0, :300].double().cpu()
            want = trace[key][0]
            err = float((got - want).abs().max()) / max(float(want.abs().max()), 1e-30)
            assert err <= (1e-6 if name == "hid2" else 5e-7), (s_, name, err)
        a0 += pb[s_].n_atoms
        b0 += pb[s_].n_bonds


def test_joint_encoder_paths_agree_bit_for_bit():
    """The shared-weight encoder runs once over [reactant rows | product rows] (rr_model.cu: joint graph).  Store-backed batches are
    assembled into one blob with adjacent feature arrays (no copy); host-packed batches, or two independently assembled DeviceGraphs,
    get their features copied next to each other inside the workspace.  Same kernels, same rows: scores and gradients are identical, with
    one segment and with several, with and without different max_num_bonds on the two sides (a star molecule among the reactants only)."""
    from reactranker_b200.data.load_reactions import Parsing_features
    dev = torch.device("cuda", GPU)
    for star in (None, {1: 7}):
        sizes = [6, 4, 5]
        ds = synthetic.make_dataset(61, sizes, star_leaves_in_group=star)
        fz = Parsing_features(ds.mols)
        feats = ds.temp.reshape(-1, 1)
        torch.manual_seed(2)
        model = make_model(300, "mle", 3, 3).train()
        model.dedup_reactants = False
        res = []
        for how in ("pair", "separate", "host"):
            if how == "host":
                rg = DeviceGraph.from_batches([BatchMolGraph([ds.mols[t] for t in ds.rsmi])], dev)
                pg = DeviceGraph.from_batches([BatchMolGraph([ds.mols[t] for t in ds.psmi])], dev)
            else:
                r_b, p_b = fz.parsing_smiles(list(ds.rsmi)), fz.parsing_smiles(list(ds.psmi))
                rg, pg = DeviceGraph.pair_from_batches([r_b], [p_b], dev) if how == "pair" else (r_b.to_device(dev), p_b.to_device(dev))
            adjacent = pg.c.f_atoms == rg.c.f_atoms + rg.n_atoms * 64 * 4 and pg.c.f_bonds == rg.c.f_bonds + rg.n_bonds * 88 * 4
            assert adjacent == (how == "pair")
            model.zero_grad()
            out = model(rg, pg, gpu=GPU, add_features=feats)
            product_loss("mle", out, sizes, torch.tensor(ds.lgk.astype(np.float32))).backward()
            res.append((out.detach().clone(), [p.grad.clone() for p in model.hot_parameters()]))
        for other in res[1:]:
            assert torch.equal(other[0], res[0][0])
            # weight gradients are accumulated with float atomics (split-K reductions): equal up to their order
            for a, b in zip(other[1], res[0][1]):
                assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-12
        # several segments per side (the RankNet window / evaluation layout)
        o, rs, ps = 0, [], []
        for n in sizes:
            rs.append(fz.parsing_smiles(list(ds.rsmi[o:o + n])))
            ps.append(fz.parsing_smiles(list(ds.psmi[o:o + n])))
            o += n
        with torch.no_grad():
            a = model.eval()(*DeviceGraph.pair_from_batches(rs, ps, dev), gpu=GPU, add_features=feats)
            b = model(DeviceGraph.from_batches(rs, dev), DeviceGraph.from_batches(ps, dev), gpu=GPU, add_features=feats)
        assert torch.equal(a, b)


# ---- optional reactant de-duplication (rr_model_cfg.r_atom_map): exact without dropout ----------------------------------------------
def _store_batches(ds):
    from reactranker_b200.data.load_reactions import Parsing_features
    fz = Parsing_features(ds.mols)
    return fz, fz.parsing_smiles(list(ds.rsmi)), fz.parsing_smiles(list(ds.psmi))


@pytest.mark.parametrize("star", [None, {1: 7}])
def test_dedup_reactants_eval_scores_identical(star):
    from reactranker_b200 import synthetic
    ds = synthetic.make_dataset(51, [6, 4, 5], star_leaves_in_group=star)
    model = make_model(300, "mle", 3, 3).eval()
    fz, r_b, p_b = _store_batches(ds)
    feats = ds.temp.reshape(-1, 1)
    with torch.no_grad():
        model.dedup_reactants = False
        want = model(r_b, p_b, gpu=GPU, add_features=feats)
        model.dedup_reactants = True
        got = model(r_b, p_b, gpu=GPU, add_features=feats)
        assert torch.equal(got, want)                         # same rows, same arithmetic per row
        # one segment per group (the evaluation path): every segment keeps its own padding rows and max_num_bonds
        from reactranker_b200.features.featurization import DeviceGraph
        sizes, o, rs, ps = [6, 4, 5], 0, [], []
        for n in sizes:
            rs.append(fz.parsing_smiles(list(ds.rsmi[o:o + n])))
            ps.append(fz.parsing_smiles(list(ds.psmi[o:o + n])))
            o += n
        want_seg = model(DeviceGraph.from_batches(rs, "cuda:0"), DeviceGraph.from_batches(ps, "cuda:0"), gpu=GPU, add_features=feats)
        rg, pg = DeviceGraph.from_batches_dedup(rs, ps, "cuda:0")
        assert rg.n_mols == 3 and getattr(rg, "atom_map", None) is not None
        assert torch.equal(model(rg, pg, gpu=GPU, add_features=feats), want_seg)


def test_dedup_reactants_training_gradients_match_at_dropout_zero():
    from reactranker_b200 import synthetic
    ds = synthetic.make_dataset(52, [5, 7, 3])
    sizes = [5, 7, 3]
    fz, r_b, p_b = _store_batches(ds)
    feats = ds.temp.reshape(-1, 1)
    targets = torch.FloatTensor(ds.lgk.reshape(-1, 1)).squeeze()
    grads = {}
    for dedup in (False, True):
        torch.manual_seed(5)
        model = make_model(64, "mle", 3, 3, dropout=0.0).train()
        model.dedup_reactants = dedup
        out = model(r_b, p_b, gpu=GPU, add_features=feats)
        loss = product_loss("mle", out, sizes, targets)
        loss.backward()
        grads[dedup] = ({k: p.grad.double().cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}, out.detach().cpu())
    assert torch.equal(grads[True][1], grads[False][1])
    assert not grads_close(grads[True][0], grads[False][0], 2e-4)     # summation order of the shared rows' gradient differs, nothing else


def test_dedup_is_refused_with_dropout():
    from reactranker_b200 import synthetic
    from reactranker_b200.features.featurization import DeviceGraph
    ds = synthetic.make_dataset(53, [4, 4])
    fz, r_b, p_b = _store_batches(ds)
    model = make_model(40, "mle", 3, 3, dropout=0.2).train()
    rg, pg = DeviceGraph.from_batches_dedup([r_b], [p_b], "cuda:0")
    with pytest.raises(Exception, match="dropout"):
        model(rg, pg, gpu=GPU, add_features=ds.temp.reshape(-1, 1))
    out = model(r_b, p_b, gpu=GPU, add_features=ds.temp.reshape(-1, 1))      # BatchMolGraph inputs: falls back to one row per candidate
    assert out.shape[0] == 8


def test_ranknet_window_properties_at_config3_size():
    """Config-3 scale (RankNet, one accumulation window of 64 groups x 64 candidates as 64 segments of ONE launch, hidden 300):
    (a) a group's scores equal its own single-group forward (the reference's one forward per group), (b) the pairwise cost is invariant
    to a shift of a group's scores: dL/dscore sums to zero inside every group, (c) loss and gradients are finite."""
    from reactranker_b200.features.featurization import DeviceGraph
    from reactranker_b200.train.loss import count_ordered_pairs, ranknet_window_loss
    G, n = 64, 64
    ds = synthetic.make_dataset(321, [n] * G)
    torch.manual_seed(2)
    model = make_model(300, "mle", 3, 3, last="no_softplus").eval()
    feats = ds.temp.reshape(-1, 1)
    r_all = [ds.mols[t] for t in ds.rsmi]
    p_all = [ds.mols[t] for t in ds.psmi]
    dev = torch.device("cuda", GPU)
    rb = [BatchMolGraph(r_all[g * n:(g + 1) * n]) for g in range(G)]
    pb = [BatchMolGraph(p_all[g * n:(g + 1) * n]) for g in range(G)]
    with torch.no_grad():
        window = model(DeviceGraph.from_batches(rb, dev), DeviceGraph.from_batches(pb, dev), gpu=GPU, add_features=feats)
        for g in (0, 17, 63):
            alone = model(rb[g], pb[g], gpu=GPU, add_features=feats[g * n:(g + 1) * n])
            assert rel_err(window[g * n:(g + 1) * n].cpu().numpy(), alone.cpu().numpy()) < 1e-6
    model.train()
    y = model(DeviceGraph.from_batches(rb, dev), DeviceGraph.from_batches(pb, dev), gpu=GPU, add_features=feats)
    y.retain_grad()
    pairs = sum(count_ordered_pairs(ds.lgk[g * n:(g + 1) * n]) for g in range(G))
    assert pairs == G * n * (n - 1)
    loss = ranknet_window_loss(y, [n] * G, ds.lgk.astype(np.float32), pairs, sigma=1.0, gpu=GPU)
    loss.backward()
    d = y.grad.double().reshape(G, n)
    assert float(d.sum(1).abs().max()) < 1e-5 * float(d.abs().sum(1).max())
    assert bool(torch.isfinite(loss).all()) and float(loss.detach()) > 0
    assert all(bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.requires_grad)


def test_config4_model_properties_at_full_size():
    """Config-4 scale (UC-Listwise, hidden 600, depth 5, 128 groups x 32 = 4096 reactions per GPU): a shard packed with the global
    max_num_bonds reproduces its rows, the variance column is positive, ListMLE-style shift of the score column leaves the softmax terms
    unchanged, loss and gradients are finite."""
    from reactranker_b200.features.featurization import DeviceGraph
    G, n = 128, 32
    ds = synthetic.make_dataset(77, [n] * G)
    torch.manual_seed(3)
    model = make_model(600, "evidential_ranking", 5, 5).eval()
    feats = ds.temp.reshape(-1, 1)
    r_all = [ds.mols[t] for t in ds.rsmi]
    p_all = [ds.mols[t] for t in ds.psmi]
    dev = torch.device("cuda", GPU)
    with torch.no_grad():
        r_g, p_g = BatchMolGraph(r_all), BatchMolGraph(p_all)
        full = model(r_g, p_g, gpu=GPU, add_features=feats)
        assert tuple(full.shape) == (G * n, 2) and bool(torch.isfinite(full).all()) and bool((full[:, 1] > 0).all())
        q = G // 4 * n
        rs, ps = BatchMolGraph(r_all[q:2 * q]), BatchMolGraph(p_all[q:2 * q])
        shard = model(DeviceGraph.from_batches([rs], dev, [r_g.max_num_bonds]), DeviceGraph.from_batches([ps], dev, [p_g.max_num_bonds]),
                      gpu=GPU, add_features=feats[q:2 * q])
        assert rel_err(shard.cpu().numpy(), full[q:2 * q].cpu().numpy()) < 1e-6
    model.train()
    out = model(r_g, p_g, gpu=GPU, add_features=feats)
    loss = RL.evidential_ranking()(out, [n] * G, torch.FloatTensor(ds.lgk), 0.0001, 0, 1, GPU)
    loss.backward()
    assert bool(torch.isfinite(loss).all())
    assert all(bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.requires_grad)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """The `gpu` argument may name any device: per-device kernel attributes (dynamic shared memory limits), SM counts, workspaces and the
    molecule store mirror must not leak from cuda:0 to cuda:1.  Same step on both devices -> same scores, loss and gradients."""
    ds = synthetic.make_dataset(9, [6, 5, 7])
    sizes = [6, 5, 7]
    res = []
    for dev in (0, 1, 0):
        torch.manual_seed(5)
        model = build_model(hidden_size=300, mpnn_depth=3, mpnn_diff_depth=3, ffn_depth=3, use_bias=True, dropout=0.0, task_num=1,
                            ffn_last_layer="with_softplus", add_features_dim=1).cuda(dev)
        r_g = BatchMolGraph([ds.mols[t] for t in ds.rsmi])
        p_g = BatchMolGraph([ds.mols[t] for t in ds.psmi])
        out = model(r_g, p_g, gpu=dev, add_features=ds.temp.reshape(-1, 1))
        loss = RL.MLEloss()(out, sizes, torch.FloatTensor(ds.lgk), dev)
        loss.backward()
        assert out.device.index == dev
        res.append((out.detach().cpu(), loss.detach().cpu(), {k: p.grad.cpu().numpy() for k, p in model.named_parameters() if p.requires_grad}))
    for other in res[1:]:
        assert rel_err(other[0].numpy(), res[0][0].numpy()) < 1e-6 and rel_err(other[1].numpy(), res[0][1].numpy()) < 1e-6
        assert not grads_close(other[2], res[0][2], 1e-4)          # float-atomic order in the split-K weight gradients
