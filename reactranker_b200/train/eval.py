"""Validation / test metrics with the reference's names and return values (train/eval.py:76-177, 475-555).

The reference runs ONE model forward per reactant group.  Here up to ``groups_per_launch`` groups share a
launch: every group is its own segment of the DeviceGraph, so it keeps its own padding rows and its own
``max_num_bonds`` and the scores equal the one-forward-per-group scores of the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from ..features.featurization import DeviceGraph

GROUPS_PER_LAUNCH = 256


def _scores_per_group(model, gpu, groups, smiles2graph_dic):
    """groups: list of (smiles [n,2], add_features | None).  Returns a list of 1-D numpy score arrays."""
    out = []
    dev = torch.device("cuda", gpu) if isinstance(gpu, int) else torch.device(gpu)
    for lo in range(0, len(groups), GROUPS_PER_LAUNCH):
        chunk = groups[lo:lo + GROUPS_PER_LAUNCH]
        r_b = [smiles2graph_dic.parsing_smiles([s[0] for s in X]) for X, _ in chunk]
        p_b = [smiles2graph_dic.parsing_smiles([s[1] for s in X]) for X, _ in chunk]
        feats = None
        if chunk[0][1] is not None:
            feats = np.concatenate([np.asarray(f, dtype=np.float64).reshape(len(X), -1) for X, f in chunk], axis=0)
        if getattr(model, "dedup_reactants", False) and not (model.training and getattr(model, "_dropout", 0) > 0):
            rg, pg = DeviceGraph.from_batches_dedup(r_b, p_b, dev)       # a group's candidates share one reactant graph
        else:
            rg, pg = DeviceGraph.from_batches(r_b, dev), DeviceGraph.from_batches(p_b, dev)
        preds = model(rg, pg, gpu=gpu, add_features=feats)
        preds = preds[:, 0] if preds.dim() > 1 else preds
        flat = preds.detach().float().cpu().numpy()
        o = 0
        for X, _ in chunk:
            out.append(flat[o:o + len(X)])
            o += len(X)
    return out


def _desc_order(x):
    """``sorted(enumerate(x), key=lambda t: t[1], reverse=True)`` indices (stable for ties)."""
    x = np.asarray(x, dtype=np.float64)
    return np.argsort(-x, kind="stable")


def compute_NDCG(truth, pred):
    """eval.py:460-472 (exponential gain, log2 discount)."""
    truth, pred = np.asarray(truth, dtype=np.float64), np.asarray(pred, dtype=np.float64)
    length = len(truth)
    disc = np.log2(np.arange(2, length + 2))
    return float(np.sum(np.exp(pred) / disc) / np.sum(np.exp(truth) / disc))


def ranking_metrics(model, gpu, data_processor, smiles2graph_dic, show_info=True, smiles_list=None, target_name: str = 'ea',
                    logger=None, add_features_name=None):
    """top-1 hit, recall@25 %, top-25 % hit, [NDCG@1, NDCG@2, NDCG@25 %, NDCG@all] (eval.py:475-555).  As in the
    reference the groups come from ``generate_batch_per_query`` (so the extra feature is the target column, load_reactions.py:264)."""
    was_training = model.training
    model.eval()
    groups, targets = [], []
    for X, t, feats in data_processor.generate_batch_per_query(smiles_list=smiles_list, target_name=target_name, shuffle_query=False,
                                                                shuffle_batch=False, add_features_name=add_features_name):
        groups.append((X, feats))
        targets.append(np.asarray(t, dtype=np.float64))
    with torch.no_grad():
        scores = _scores_per_group(model, gpu, groups, smiles2graph_dic)
    top1 = top25 = 0
    recall, ndcgs = [], []
    for pred, targ in zip(scores, targets):
        n = len(targ)
        p_idx, t_idx = _desc_order(pred), _desc_order(targ)
        t_sorted = targ[t_idx]
        top1 += int(p_idx[0] == t_idx[0])
        len25 = max(1, round(n * 0.25))
        p25, t25 = p_idx[:len25], set(t_idx[:len25].tolist())
        top25 += int(p25[0] in t25)
        recall.append(sum(int(i in t25) for i in p25.tolist()) / len25)
        by_pred = targ[p_idx]
        # NDCG@2 in the reference wraps its two-item slices in a list (eval.py:544): one position, both gains summed
        ndcg2 = float(np.sum(np.exp(by_pred[:2])) / np.sum(np.exp(t_sorted[:2])))
        ndcgs.append([compute_NDCG(t_sorted[:1], by_pred[:1]), ndcg2, compute_NDCG(t_sorted[:len25], by_pred[:len25]),
                      compute_NDCG(t_sorted, by_pred)])
    model.train(was_training)
    k = max(len(scores), 1)
    return top1 / k, float(np.mean(recall)), top25 / k, np.mean(ndcgs, axis=0)


def evaluate_top_scores(model, gpu, data_processor, smiles2graph_dic, ratio=0.25, batch_size=2, show_info=False, smiles_list=None,
                        target_name: str = 'ea', add_features_name=None):
    """top-1 accuracy, mean overlap of predicted/true top-``ratio`` sets, true top-1 inside predicted top-``ratio``
    (eval.py:76-177).  Groups come from ``generate_batch_querys`` (real ``add_features_name`` column); in the reference
    ``batch_size`` groups share one BatchMolGraph and therefore one ``max_num_bonds`` -- reproduced by packing each
    ``batch_size`` chunk as one segment."""
    score, overlap, top1_in = [], [], []
    with torch.no_grad():
        for X, targets, scope, feats in data_processor.generate_batch_querys(smiles_list=smiles_list, target_name=target_name,
                                                                              batch_size=batch_size, shuffle_query=False, shuffle_batch=False,
                                                                              add_features_name=add_features_name):
            r_b = smiles2graph_dic.parsing_smiles([s[0] for s in X])
            p_b = smiles2graph_dic.parsing_smiles([s[1] for s in X])
            preds = model(r_b, p_b, gpu=gpu, add_features=feats)
            preds = (preds[:, 0] if preds.dim() > 1 else preds).detach().float().cpu().numpy()
            targets = np.asarray(targets, dtype=np.float64).reshape(-1)
            o = 0
            for n in scope:
                t, p = targets[o:o + n], preds[o:o + n]
                o += n
                t_idx, p_idx = _desc_order(t), _desc_order(p)
                score.append(int(int(np.argmax(t)) == int(np.argmax(p))))
                length = max(1, round(n * ratio))
                tset = set(t_idx[:length].tolist())
                overlap.append(sum(int(i in tset) for i in p_idx[:length].tolist()) / length)
                top1_in.append(int(int(np.argmax(t)) in set(p_idx[:length].tolist())))
    return sum(score) / len(score), sum(overlap) / len(overlap), sum(top1_in) / len(top1_in)


def calculate_ndcg(model, gpu, data_processor, smiles2graph_dic, batch_size=2, NDCG_cut=0.5, show_info=False, smiles_list=None,
                   target_name: str = 'ea', logger=None, is_order=True, means=None, stds=None, add_features_name=None):
    """Rank-based NDCG@cut, KL(softmax(targets) || softmax(scores)), the per-item order table and the re-ordered SMILES of the
    test report (eval.py:329-457).  ``batch_size`` groups share one forward (one BatchMolGraph, one ``max_num_bonds``) as in the
    reference; everything after the forward is the reference's host arithmetic in fp32:
      * ``means/stds`` de-normalise scores (and scale the variance column by ``stds**2``) first (eval.py:379-387);
      * per group: items sorted by target (descending); ``pred_order`` = rank of each item's score among the group (1 = best);
        gains are ``n + 1 - order`` and NDCG uses the first ``ceil(n * NDCG_cut)`` positions with a log2 discount (eval.py:309-326, 426-430);
      * rows of ``total_order``: [target, score, (uncertainty,) true_order, pred_order]; ``smiles_and_idx``: [iteration, rsmi, psmi].
    ``is_order=False`` only collects [target, score...] rows and SMILES pairs and returns ``None`` for both means."""
    import math
    ndcg_list, kl_list, total_order, smiles_and_idx = [], [], [], []
    it = 0
    with torch.no_grad():
        for X, targets, scope, add_features in data_processor.generate_batch_querys(smiles_list=smiles_list, target_name=target_name,
                                                                                     batch_size=batch_size, shuffle_query=False, shuffle_batch=False,
                                                                                     add_features_name=add_features_name):
            rsmi, psmi = [s[0] for s in X], [s[1] for s in X]
            preds_ini = model(smiles2graph_dic.parsing_smiles(rsmi), smiles2graph_dic.parsing_smiles(psmi), gpu=gpu, add_features=add_features)
            it += 1
            preds_ini = preds_ini.detach().float().cpu()
            if means is not None:
                if preds_ini.dim() > 1:
                    preds_ini = torch.stack((preds_ini[:, 0] * stds + means, preds_ini[:, 1] * (stds ** 2)), dim=1)
                else:
                    preds_ini = preds_ini * stds + means
            with_unc = preds_ini.dim() > 1
            preds = preds_ini[:, 0] if with_unc else preds_ini
            unc = preds_ini[:, 1] if with_unc else None
            t_all = torch.FloatTensor(np.squeeze(targets).tolist())
            if not is_order:
                rows = preds_ini.numpy().tolist()
                total_order.extend([[t] + r for t, r in zip(t_all.tolist(), rows)] if with_unc else [[t, r] for t, r in zip(t_all.tolist(), rows)])
                smiles_and_idx.extend([[a, b] for a, b in zip(rsmi, psmi)])
                continue
            o = 0
            for n in scope:
                bt, bp = t_all[o:o + n], preds[o:o + n]
                P, Q = torch.softmax(bt, 0), torch.softmax(bp, 0)
                # eval.py:400-402 writes exp(x)/sum(exp(x)) without a max shift; identical in exact arithmetic, and fp32-identical to ~1e-7
                P = torch.exp(bt) / torch.sum(torch.exp(bt))
                Q = torch.exp(bp) / torch.sum(torch.exp(bp))
                kl_list.append(float(torch.sum(P * torch.log(P / Q))))
                st, idx = torch.sort(bt, descending=True)
                sp = bp[idx]
                pred_order = torch.argsort(torch.argsort(sp, descending=True)) + 1
                true_order = torch.arange(n, dtype=torch.float32) + 1
                cols = [st, sp] + ([unc[o:o + n][idx]] if with_unc else []) + [true_order, pred_order.float()]
                total_order.extend(torch.stack(cols, dim=1).tolist())
                k = math.ceil(n * NDCG_cut)
                disc = torch.log2(torch.arange(min(k, n), dtype=torch.float32) + 2)
                pred_gain, true_gain = (n + 1 - pred_order).float()[:k], (n + 1 - true_order)[:k]
                ndcg_list.append(float(torch.sum(pred_gain / disc) / torch.sum(true_gain / disc)))
                order = idx.tolist()
                smiles_and_idx.extend([[it, rsmi[o + i], psmi[o + i]] for i in order])
                o += n
    if not is_order:
        return None, None, total_order, smiles_and_idx
    return np.mean(np.array(ndcg_list), axis=0), np.mean(np.array(kl_list)), total_order, smiles_and_idx


def _not_built(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name} is an inference-side metric outside the hot path (SURVEY.md §2 row 8)")
    f.__name__ = name
    return f


calculate_mse = _not_built("calculate_mse")
pairwise_acc = _not_built("pairwise_acc")
pairwise_baseline_acc = _not_built("pairwise_baseline_acc")
eval_cross_entropy_loss = _not_built("eval_cross_entropy_loss")
