"""End-to-end training through the reference-shaped entry points on synthetic reactions (config 1 plumbing), and the
data-parallel step (2 GPUs, skipped on a one-GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, timeout=600):
    res = subprocess.run([sys.executable] + cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


def test_save_metric_mse_trains_and_checkpoints(tmp_path):
    """save_metric='mse' (train_listwise.py:345-352): the validation MSE is computed every epoch and the best checkpoint kept -- the
    reference's own calculate_mse cannot run; here it does."""
    out = run(["main.py", "--synthetic", "40,10", "--path", str(tmp_path), "--gpu", "0", "--task_type", "regression", "--batch_size", "100",
               "--total_epochs", "3", "--hidden_size", "64", "--save_metric", "mse"])
    assert out.count("the validation MSE over") == 3
    assert os.path.exists(os.path.join(str(tmp_path), "0.pt"))


@pytest.mark.parametrize("task", ["mle", "listnet", "evidential_ranking", "gauss_regression", "regression",
                                  "mledis_gaussian", "listnet_uq", "dirichlet_uq", "listnetdis_lognorm", "evidential", "mledis_evidential"])
def test_main_trains_every_task_key(tmp_path, task):
    out = run(["main.py", "--synthetic", "40,10", "--path", str(tmp_path), "--gpu", "0", "--task_type", task, "--batch_size", "100",
               "--total_epochs", "3", "--hidden_size", "64", "--max_lr", "3e-3"])
    losses = [float(l.rsplit(":", 1)[1]) for l in out.splitlines() if "train loss" in l]
    assert len(losses) == 3 and all(np.isfinite(losses))
    assert os.path.exists(os.path.join(str(tmp_path), "T1", "0.pt"))
    assert "test score for k_fold vailidation" in out


def test_resume_continues_the_interrupted_run(tmp_path):
    """two epochs, stop, resume for two more == four epochs in one go (same batches, Adam moments, Noam position).
    Rounding order differs between runs only through the kernels' float atomics, so the late losses agree to 1e-3, not bit for bit."""
    args = ["main.py", "--synthetic", "40,10", "--gpu", "0", "--task_type", "mle", "--batch_size", "100", "--total_epochs", "4",
            "--hidden_size", "64", "--max_lr", "3e-3", "--dropout", "0.0"]
    whole = run(args + ["--path", str(tmp_path / "a")])
    run(args + ["--path", str(tmp_path / "b"), "--resume", "--stop_after", "2"])
    assert os.path.exists(str(tmp_path / "b" / "0.state.pt"))
    rest = run(args + ["--path", str(tmp_path / "b"), "--resume"])
    assert "resuming after epoch 2" in rest
    want = [float(l.rsplit(":", 1)[1]) for l in whole.splitlines() if "train loss" in l]
    got = [float(l.rsplit(":", 1)[1]) for l in rest.splitlines() if "train loss" in l]
    assert len(want) == 4 and len(got) == 2
    assert np.allclose(got, want[2:], rtol=1e-3), (got, want)
    # and a weights-only checkpoint cannot be resumed from
    from reactranker_b200.utils import load_checkpoint
    assert "optimizer" in load_checkpoint(str(tmp_path / "b" / "0.state.pt"))


@pytest.mark.parametrize("algo", ["sum_session", "accelerate_grad"])
def test_main_ranknet_trains(tmp_path, algo):
    out = run(["main_ranknet.py", "--synthetic", "40,8", "--path", str(tmp_path), "--gpu", "0", "--batch_size", "64", "--total_epochs", "3",
               "--train_strategy", algo])
    losses = [float(l.rsplit(":", 1)[1]) for l in out.splitlines() if "train loss" in l]
    assert len(losses) == 3 and all(np.isfinite(losses)) and 0.3 < losses[0] < 1.5      # ~log 2 per pair at initialisation
    assert "test score for k_fold vailidation" in out


def test_listmle_training_reduces_loss_and_checkpoint_roundtrip(tmp_path):
    from reactranker_b200 import synthetic
    from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features
    from reactranker_b200.models.base_model import build_model
    from reactranker_b200.train.loss import MLEloss
    from reactranker_b200.train.utils import build_optimizer
    from reactranker_b200.utils import load_checkpoint, save_checkpoint
    ds = synthetic.make_dataset(3, [12] * 20)
    fz = Parsing_features(ds.mols)
    df = ds.to_dataframe()
    torch.manual_seed(0)
    model = build_model(hidden_size=64, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0).cuda(0)
    opt = build_optimizer(model)
    opt.param_groups[0]["lr"] = 3e-3
    dp, crit, first, last = DataProcessor(df), MLEloss(), None, None
    for epoch in range(12):
        for reactions, targets, scope, feats in dp.generate_batch_reactions(smiles_list=["rsmi_mapped", "psmi_mapped"], target_name="lgk",
                                                                             batch_size=120, seed=epoch, add_features_name="temp"):
            r, p = fz.parsing_reactions(reactions)
            loss = crit(model(r, p, gpu=0, add_features=feats), scope, torch.FloatTensor(targets).squeeze(), 0)
            opt.zero_grad()
            loss.backward()
            opt.step()
            first = float(loss) if first is None else first
            last = float(loss)
    assert last < 0.9 * first, (first, last)
    path = os.path.join(str(tmp_path), "ck.pt")
    save_checkpoint(path, model, np.float64(1.5), np.float64(0.5))
    state = torch.load(path)                                       # weights_only=True default must work
    assert state["data_scaler"] == {"means": 1.5, "stds": 0.5} and len(state["state_dict"]) == 20
    m2 = build_model(hidden_size=64, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=0.0)
    m2.load_state_dict(load_checkpoint(path)["state_dict"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_step_equals_one_gpu_step(tmp_path):
    out = run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29611",
               "tests/dp_equivalence.py"])
    assert "DP-EQUIVALENCE-OK" in out


def _loss_curve(out):
    return [float(l.rsplit("=", 1)[1]) for l in out.splitlines() if "full precision" in l]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("task", ["mle", "listnet"])
def test_torchrun_main_reproduces_the_one_gpu_loss_curve(tmp_path, task):
    """The product entry point under torchrun (2 ranks: global batch plan, shards of whole groups, global max_num_bonds and normalisers,
    in-place gradient all-reduce, rank 0 validates / checkpoints / tests) trains the same model as the plain one-GPU run: the per-epoch
    losses agree to 1e-4 over four epochs of Adam + NoamLR at dropout 0."""
    args = ["main.py", "--synthetic", "40,10", "--task_type", task, "--batch_size", "100", "--total_epochs", "4", "--hidden_size", "64",
            "--max_lr", "3e-3", "--dropout", "0.0"]
    one = run(args + ["--gpu", "0", "--path", str(tmp_path / "one")])
    two = run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29633"]
              + args + ["--path", str(tmp_path / "two")])
    want, got = _loss_curve(one), _loss_curve(two)
    assert len(want) == 4 and len(got) == 4, (one[-2000:], two[-2000:])
    assert np.allclose(got, want, rtol=1e-4), (got, want)
    assert os.path.exists(os.path.join(str(tmp_path / "two"), "T1", "0.pt"))
    assert two.count("test score for k_fold vailidation") == 1                  # rank 0 alone tests and reports


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_torchrun_main_ranknet_reproduces_the_one_gpu_loss_curve(tmp_path):
    args = ["main_ranknet.py", "--synthetic", "40,8", "--batch_size", "64", "--total_epochs", "3"]
    env_note = "RankNet's entry script fixes dropout 0.2 (main_ranknet.py:113-121): the curves agree statistically, not to rounding"
    one = run(args + ["--gpu", "0", "--path", str(tmp_path / "one")])
    two = run(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29634"]
              + args + ["--path", str(tmp_path / "two")])
    want, got = _loss_curve(one), _loss_curve(two)
    assert len(want) == 3 and len(got) == 3, env_note
    assert np.allclose(got, want, rtol=0.15), (got, want, env_note)
    assert two.count("test score for k_fold vailidation") == 1


@pytest.mark.parametrize("two", [False, True])
def test_validation_metrics_on_device_match_reference_golden(two, monkeypatch):
    """ranking_metrics / evaluate_top_scores as the product runs them (scores on the device, rr_rank_metrics, one host wait) return what
    the reference's functions returned for the same stub scorer (tests/golden/metrics.npz, tests/golden/make_golden.py golden_metrics)."""
    from test_host_cpu import _StubScorer
    from reactranker_b200 import synthetic
    from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features
    from reactranker_b200.train import eval as E
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics.npz"))
    ds = synthetic.make_dataset(int(g["seed"]), [int(x) for x in g["sizes"]], atoms_lo=3, atoms_hi=4)
    fz = Parsing_features(ds.mols)
    dp = DataProcessor(ds.to_dataframe())
    cols = ["rsmi_mapped", "psmi_mapped"]
    m = _StubScorer(two).eval()
    tag = "two." if two else "one."

    def stub_chunks(model, gpu, chunks, smiles2graph_dic):       # the stub scores tokens, not graphs: same chunking, scores put on the device
        for lo in range(0, len(chunks), 2):
            part = chunks[lo:lo + 2]
            preds = [model(None, smiles2graph_dic.parsing_smiles([s[1] for s in X]), add_features=f) for X, f in part]
            yield lo, lo + len(part), torch.cat(preds).cuda(0)
    monkeypatch.setattr(E, "_forward_chunks", stub_chunks)
    got = E.evaluate_top_scores(m, gpu=0, data_processor=dp, smiles2graph_dic=fz, ratio=0.25, batch_size=3, smiles_list=cols, target_name="lgk",
                                add_features_name="temp")
    assert np.allclose(got, g[tag + "top_scores"], rtol=0, atol=1e-12)
    r = E.ranking_metrics(m, gpu=0, data_processor=dp, smiles2graph_dic=fz, show_info=False, smiles_list=cols, target_name="lgk", add_features_name="temp")
    assert np.allclose([r[0], r[1], r[2]] + list(r[3]), g[tag + "ranking"], rtol=1e-12, atol=1e-12)


def test_planned_rows_path_equals_the_gathered_smiles_path():
    """TrainStep.prepare_rows (the planner's row positions -> per-frame store-id vectors -> rr_batch_build -> pair assembly on the side
    stream; what train() and bench.py use) builds byte-for-byte the graphs, targets and extra features that TrainStep.prepare builds from
    the gathered (smiles, targets, scope, add_features) tuples of generate_batch_reactions, with and without reactant de-duplication;
    the two steps then give the same loss."""
    from reactranker_b200 import synthetic
    from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features
    from reactranker_b200.features.featurization import DeviceGraph
    from reactranker_b200.models.base_model import build_model
    from reactranker_b200.train.step import TrainStep
    ds = synthetic.make_dataset(91, [7, 5, 9, 6, 8], star_leaves_in_group={2: 6})
    fz = Parsing_features(ds.mols)
    proc = DataProcessor(ds.to_dataframe())
    cols = ["rsmi_mapped", "psmi_mapped"]
    for dropout in (0.0, 0.1):                       # 0.0: repeated reactants are encoded once (r_atom_map)
        torch.manual_seed(1)
        model = build_model(hidden_size=64, task_num=1, ffn_last_layer="with_softplus", add_features_dim=1, dropout=dropout).cuda(0).train()
        opt = torch.optim.SGD(model.parameters(), lr=0.0)
        step = TrainStep(model, opt, torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0), "mle", 0)
        batches = list(proc.generate_batch_reactions(smiles_list=cols, target_name="lgk", batch_size=20, seed=2, add_features_name="temp"))
        plans = list(proc.plan_batch_reactions(batch_size=20, seed=2))
        assert len(batches) == len(plans) >= 2
        for batch, (rows, scope) in zip(batches, plans):
            a = step.prepare(batch, fz, device_graphs=True)
            b = step.prepare_rows(proc, rows, scope, fz, cols, "lgk", "temp")
            torch.cuda.synchronize()
            assert a.scope == b.scope and (a.groups, a.items, a.rows) == (b.groups, b.items, b.rows)
            assert torch.equal(a.targets, b.targets) and torch.equal(a.feats, b.feats)
            for g1, g2 in ((a.r, b.r), (a.p, b.p)):
                assert (g1.n_atoms, g1.n_bonds, g1.n_mols, g1.c.wmax, g1.c.n_segments) == (g2.n_atoms, g2.n_bonds, g2.n_mols, g2.c.wmax, g2.c.n_segments)
                for name, nbytes in DeviceGraph._sections(g1.n_atoms, g1.n_bonds, g1.n_mols, g1.c.wmax, g1.c.n_segments):
                    assert torch.equal(g1.blob[g1._offs[name]:g1._offs[name] + nbytes], g2.blob[g2._offs[name]:g2._offs[name] + nbytes]), name
            assert (getattr(a.r, "atom_map", None) is None) == (dropout > 0)
            if dropout == 0:
                assert torch.equal(a.r.atom_map, b.r.atom_map)
                la, lb = float(step.run(a)), float(step.run(b))
                assert la == lb
