"""Parity against the LIVE reference (imported from /root/reference with rdkit stubbed, oracle/ref_loader.py) on freshly seeded inputs
that the committed goldens do not contain.  Runs in the build container only: on the GPU box /root/reference does not exist and the
whole module is skipped (the goldens under tests/golden/ carry the parity there).

  * host logic of the PRODUCT (batch planners, BatchMolGraph tensors): bit-exact against the reference's own classes;
  * the ORACLE (model forward + every runnable task key's loss + gradients): against the reference's modules in fp64, so that the
    checker the GPU tests rely on is held to the reference on more than the golden cases.
"""
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import reactranker_oracle as O
from reactranker_b200 import synthetic

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (build container only)")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
COLS = ["rsmi_mapped", "psmi_mapped"]


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_planners_bit_exact_on_fresh_frames(seed):
    """generate_batch_reactions (incl. truncation and tail yields), generate_batch_per_query, generate_batch_querys
    (load_reactions.py:235-421): same rows, same order, same scope, same targets / extra features, for random group sizes and seeds."""
    from reactranker_b200.data.load_reactions import DataProcessor
    lr = ref_loader.ref("data.load_reactions")
    rng = np.random.default_rng(seed)
    sizes = [int(x) for x in rng.integers(1, 28, size=int(rng.integers(5, 14)))]
    ds = synthetic.make_dataset(seed, sizes, atoms_lo=3, atoms_hi=4)
    df = ds.to_dataframe()
    ours, ref = DataProcessor(df), lr.DataProcessor(df)
    for batch_size in (int(rng.integers(2, 12)), int(rng.integers(12, 40)), 200):
        for epoch in (0, int(rng.integers(1, 50))):
            kw = dict(smiles_list=COLS, target_name="lgk", batch_size=batch_size, seed=epoch, add_features_name="temp")
            a, b = list(ours.generate_batch_reactions(**kw)), list(ref.generate_batch_reactions(**kw))
            assert len(a) == len(b)
            for (s1, t1, c1, f1), (s2, t2, c2, f2) in zip(a, b):
                assert np.array_equal(np.asarray(s1, dtype=object), np.asarray(s2, dtype=object)) and list(c1) == list(c2)
                assert np.array_equal(np.asarray(t1, np.float64), np.asarray(t2, np.float64))
                assert np.array_equal(np.asarray(f1, np.float64), np.asarray(f2, np.float64))
    for epoch in (0, 7):
        kw = dict(smiles_list=COLS, target_name="lgk", seed=epoch, add_features_name="temp")
        for (s1, t1, f1), (s2, t2, f2) in zip(ours.generate_batch_per_query(**kw), ref.generate_batch_per_query(**kw)):
            assert np.array_equal(np.asarray(s1, dtype=object), np.asarray(s2, dtype=object))
            assert np.array_equal(np.asarray(t1, np.float64).reshape(-1), np.asarray(t2, np.float64).reshape(-1))
            assert np.array_equal(np.asarray(f1, np.float64).reshape(-1), np.asarray(f2, np.float64).reshape(-1))
        kw = dict(smiles_list=COLS, target_name="lgk", batch_size=3, seed=epoch, add_features_name="temp", shuffle_query=False, shuffle_batch=False)
        a, b = list(ours.generate_batch_querys(**kw)), list(ref.generate_batch_querys(**kw))
        assert len(a) == len(b)
        for (s1, t1, c1, f1), (s2, t2, c2, f2) in zip(a, b):
            assert np.array_equal(np.asarray(s1, dtype=object), np.asarray(s2, dtype=object)) and list(c1) == list(c2)
            assert np.array_equal(np.asarray(t1, np.float64).reshape(-1), np.asarray(t2, np.float64).reshape(-1))


@pytest.mark.parametrize("seed,sizes,star", [(201, [4, 1, 6], None), (202, [3, 5], {1: 9}), (203, [1], None), (204, [2, 2, 2, 7], {3: 5})])
def test_batchmolgraph_bit_exact_on_fresh_batches(seed, sizes, star):
    """BatchMolGraph (featurization.py:246-329) of the product, host-packed and store-backed, against the reference's class."""
    from reactranker_b200.data.load_reactions import Parsing_features
    from reactranker_b200.features.featurization import BatchMolGraph
    ds = synthetic.make_dataset(seed, sizes, star_leaves_in_group=star)
    ref_fz = ref_loader.RefFeaturizer(ds.mols)
    fz = Parsing_features(ds.mols)
    for col in (ds.rsmi, ds.psmi):
        want = ref_fz.parsing_smiles(list(col))
        for got in (BatchMolGraph([ds.mols[t] for t in col]), fz.parsing_smiles(list(col))):
            fa, fb, a2b, b2a, b2revb, a_scope, b_scope = got.get_components()
            wa, wb, w2b, wb2a, wrev, wa_scope, wb_scope = want.get_components()
            assert torch.equal(fa, wa) and torch.equal(fb, wb)
            assert torch.equal(a2b.long(), w2b.long()) and torch.equal(b2a.long(), wb2a.long()) and torch.equal(b2revb.long(), wrev.long())
            assert list(a_scope) == list(wa_scope) and list(b_scope) == list(wb_scope)
            assert torch.equal(got.get_a2a().long(), want.get_a2a().long())
            assert (got.n_atoms, got.n_bonds, got.max_num_bonds) == (want.n_atoms, want.n_bonds, want.max_num_bonds)


def _tasks():
    import make_golden as MG
    return MG.TASKS_ALL


@pytest.mark.parametrize("task", ["mle", "listnet", "evidential_ranking", "gauss_regression", "regression", "mle_gaussian", "listnet_gauss",
                                  "mle_regression", "listnet_regression", "regression_exploss", "mledis_gaussian", "listnetdis_gauss",
                                  "listnet_uq", "listnetdis_lognorm", "dirichlet_uq", "evidential", "mle_evidential", "mledis_evidential",
                                  "listnet_evidential"])
def test_oracle_equals_live_reference_on_fresh_inputs(task):
    """Every runnable task key (train_listwise.py:196-285): the oracle's scores, loss and all gradients against the reference's own modules
    in fp64, on a model / batch the goldens do not hold (hidden 32, depth 4 / 2, ragged groups with a single-candidate group and a star)."""
    import make_golden as MG
    tn, last, tt = MG.TASKS_ALL[task]
    seed, sizes, star, hidden, depth, ddepth = 300 + len(task), [4, 1, 7, 3], {2: 6}, 32, 4, 2
    with _quiet():
        ds, model, scores, loss, grads = MG.run_case(task, hidden, seed, sizes, star, torch.float64, depth, ddepth)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "cached_zero" not in k}
    full = dict(sd)
    full.update(params)
    r_g = O.OracleBatch([ds.mols[t] for t in ds.rsmi])
    p_g = O.OracleBatch([ds.mols[t] for t in ds.psmi])
    out = O.model_forward(full, r_g, p_g, ds.temp.reshape(-1, 1), mpnn_depth=depth, mpnn_diff_depth=ddepth, head=O.resolve_task_type(tn, last, tt))
    targets = torch.tensor(ds.lgk.astype(np.float32)).double()
    got_loss = O.loss_for_task(task, out, sizes, targets)
    got_loss.backward()
    assert tuple(out.shape) == tuple(scores.shape) and np.abs(out.detach().numpy() - scores).max() <= 1e-10 * max(1.0, np.abs(scores).max())
    # mle / evidential_ranking accumulate into an fp32 ``torch.Tensor([0])`` (loss.py:79, 493) even in an fp64 run
    tol = 2e-7 if np.asarray(loss).dtype == np.float32 else 1e-10
    assert tuple(got_loss.shape) == tuple(np.asarray(loss).shape)
    assert abs(float(got_loss.detach().reshape(-1)[0]) - float(np.asarray(loss).reshape(-1)[0])) <= tol * max(1.0, abs(float(np.asarray(loss).reshape(-1)[0])))
    gscale = max(float(np.abs(g).max()) for g in grads.values())
    for k, g in grads.items():
        assert np.abs(params[k].grad.numpy() - g).max() <= 1e-6 * gscale, k
