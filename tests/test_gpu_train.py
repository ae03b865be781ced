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


@pytest.mark.parametrize("task", ["mle", "listnet", "evidential_ranking", "gauss_regression", "regression",
                                  "mledis_gaussian", "listnet_uq", "dirichlet_uq", "listnetdis_lognorm", "evidential", "mledis_evidential"])
def test_main_trains_every_task_key(tmp_path, task):
    out = run(["main.py", "--synthetic", "40,10", "--path", str(tmp_path), "--gpu", "0", "--task_type", task, "--batch_size", "100",
               "--total_epochs", "3", "--hidden_size", "64", "--max_lr", "3e-3"])
    losses = [float(l.rsplit(":", 1)[1]) for l in out.splitlines() if "train loss" in l]
    assert len(losses) == 3 and all(np.isfinite(losses))
    assert os.path.exists(os.path.join(str(tmp_path), "T1", "0.pt"))
    assert "test score for k_fold vailidation" in out


def test_main_ranknet_trains(tmp_path):
    out = run(["main_ranknet.py", "--synthetic", "40,8", "--path", str(tmp_path), "--gpu", "0", "--batch_size", "64", "--total_epochs", "3"])
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
