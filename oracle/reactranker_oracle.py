"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ReactRanker training hot path.

A plain PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm for the
path named in BASELINE.json.  It is the checker for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it; the
product package ``reactranker_b200`` never does (it raises when its CUDA extension
is missing -- there is no CPU fallback).

Pinning: the reference ships no tests / golden vectors for this path,
so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; checked by
``tests/test_oracle_golden.py`` and, wherever the reference is present, live against its own
modules by ``tests/test_live_reference.py``).

Every function cites the reference lines it restates (paths relative to the
reference root, package ``reactranker/``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import math

import numpy as np
import torch
import torch.nn.functional as F

ATOM_FDIM = 61      # features/featurization.py:63
BOND_FDIM = 22      # features/featurization.py:64
FBOND_TOTAL = ATOM_FDIM + BOND_FDIM   # models/base_model.py:129


# ----------------------------------------------------------------------------------------
# graph batching  (features/featurization.py:246-329)
# ----------------------------------------------------------------------------------------
class OracleBatch:
    """Restates ``BatchMolGraph.__init__`` (featurization.py:246-290): row 0 of atoms and
    bonds is padding, per-molecule indices are offset by the running totals, ``a2b`` is
    right-padded with 0 up to ``max_num_bonds = max(1, max in-degree)``."""

    def __init__(self, mols: Sequence):
        f_atoms = [[0.0] * ATOM_FDIM]
        f_bonds = [[0.0] * FBOND_TOTAL]
        a2b: List[List[int]] = [[]]
        b2a = [0]
        b2revb = [0]
        self.a_scope: List[Tuple[int, int]] = []
        self.b_scope: List[Tuple[int, int]] = []
        na, nb = 1, 1
        for m in mols:
            f_atoms += list(m.f_atoms)
            f_bonds += list(m.f_bonds)
            a2b += [[b + nb for b in m.a2b[a]] for a in range(m.n_atoms)]
            b2a += [na + m.b2a[b] for b in range(m.n_bonds)]
            b2revb += [nb + m.b2revb[b] for b in range(m.n_bonds)]
            self.a_scope.append((na, m.n_atoms))
            self.b_scope.append((nb, m.n_bonds))
            na += m.n_atoms
            nb += m.n_bonds
        self.n_atoms, self.n_bonds, self.n_mols = na, nb, len(mols)
        self.max_num_bonds = max(1, max(len(x) for x in a2b))
        W = self.max_num_bonds
        self.f_atoms = torch.tensor(f_atoms, dtype=torch.float32).reshape(na, ATOM_FDIM)
        self.f_bonds = torch.tensor(f_bonds, dtype=torch.float32).reshape(nb, FBOND_TOTAL)
        self.a2b = torch.tensor([row + [0] * (W - len(row)) for row in a2b], dtype=torch.int64).reshape(na, W)
        self.b2a = torch.tensor(b2a, dtype=torch.int64)
        self.b2revb = torch.tensor(b2revb, dtype=torch.int64)

    def get_components(self):
        return self.f_atoms, self.f_bonds, self.a2b, self.b2a, self.b2revb, self.a_scope, self.b_scope

    def get_a2a(self):
        return self.b2a[self.a2b]          # featurization.py:326-327


def gather_sum(source: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``index_select_ND(source, index).sum(dim=1)`` (utils.py:176-193 + mpn.py:89-90):
    padded slots hold 0 and therefore add row 0 of ``source``."""
    return source.index_select(0, index.reshape(-1)).reshape(index.shape + source.shape[1:]).sum(dim=1)


# ----------------------------------------------------------------------------------------
# model  (models/mpn.py:61-240, models/base_model.py:10-171, 235-297)
# ----------------------------------------------------------------------------------------
def _lin(x, sd, name):
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"], b)


def _drop(x, p, training):
    return F.dropout(x, p=p, training=training) if (training and p > 0) else x


def mpn_forward(sd: Dict[str, torch.Tensor], g, depth: int, dropout: float = 0.0, training: bool = False,
                prefix: str = "encoder", trace: Optional[dict] = None) -> torch.Tensor:
    """Bond-message D-MPNN returning per-atom hiddens (mpn.py:61-108 with
    ``return_atom_hiddens=True``, base_model.py:135)."""
    dt = sd[prefix + ".W_i.weight"].dtype
    f_atoms, f_bonds, a2b, b2a, b2revb, _, _ = g.get_components()
    f_atoms, f_bonds = f_atoms.to(dt), f_bonds.to(dt)
    inp = _lin(f_bonds, sd, prefix + ".W_i")                      # mpn.py:80
    msg = torch.relu(inp)                                          # mpn.py:81
    for t in range(depth - 1):                                     # mpn.py:84-97
        a_msg = gather_sum(msg, a2b)
        pre = a_msg[b2a] - msg[b2revb]
        if trace is not None:
            trace[f"{prefix}.pre{t}"] = pre
        msg = _drop(torch.relu(inp + _lin(pre, sd, prefix + ".W_h")), dropout, training)
        if trace is not None:
            trace[f"{prefix}.msg{t + 1}"] = msg
    a_msg = gather_sum(msg, a2b)                                   # mpn.py:100-102
    hid = torch.relu(_lin(torch.cat([f_atoms, a_msg], dim=1), sd, prefix + ".W_o"))   # mpn.py:103-104
    return _drop(hid, dropout, training)                           # mpn.py:105


def mpndiff_forward(sd, diff, g, depth: int, add_features=None, dropout: float = 0.0, training: bool = False,
                    prefix: str = "diff_encoder", trace: Optional[dict] = None) -> torch.Tensor:
    """Atom-message encoder over the product graph + scope-mean readout (mpn.py:170-240)."""
    dt = diff.dtype
    _, f_bonds, a2b, _, _, a_scope, _ = g.get_components()
    f_bonds = f_bonds.to(dt)
    a2a = g.get_a2a()
    inp = _lin(diff, sd, prefix + ".W_i")                          # mpn.py:194
    msg = torch.relu(inp)
    if depth > 0:
        for _ in range(depth - 1):                                 # mpn.py:199-213
            nm = gather_sum(msg, a2a)
            nf = gather_sum(f_bonds, a2b)                          # the [-bond_fdim:] slice keeps all 83 columns
            msg = _drop(torch.relu(inp + _lin(torch.cat([nm, nf], dim=1), sd, prefix + ".W_h")), dropout, training)
        a_msg = gather_sum(msg, a2a)                               # mpn.py:215-216
        hid = _drop(torch.relu(_lin(torch.cat([diff, a_msg], dim=1), sd, prefix + ".W_o")), dropout, training)
    else:
        hid = _drop(msg, dropout, training)                        # mpn.py:220-221
    if trace is not None:
        trace[prefix + ".atom_hiddens"] = hid
    vecs = []
    for (start, size) in a_scope:                                  # mpn.py:224-235
        if size == 0:
            vecs.append(sd[prefix + ".cached_zero_vector"].to(dt))
        else:
            vecs.append(hid.narrow(0, start, size).sum(dim=0) / size)
    vecs = torch.stack(vecs, dim=0)
    if add_features is not None:                                   # mpn.py:237-238 (cast mpn.py:183)
        vecs = torch.cat([vecs, torch.as_tensor(np.asarray(add_features), dtype=torch.float32).to(device=vecs.device, dtype=dt)], dim=1)
    return vecs


def resolve_task_type(task_num: int, ffn_last_layer: str, task_type: Optional[str]) -> str:
    """``build_model``'s head-name resolution (base_model.py:252-264)."""
    if task_type is None:
        if task_num == 2:
            return "gaussian_" + ffn_last_layer
        if task_num == 4:
            return "evidential_" + ffn_last_layer
        return ffn_last_layer
    if task_type == "evidential_ranking":
        return task_type
    return task_type + "_" + ffn_last_layer


def ffn_forward(sd, x, head: str, ffn_depth: int = 3, dropout: float = 0.0, training: bool = False,
                prefix: str = "ffn.ffn") -> torch.Tensor:
    """``FFN`` (base_model.py:10-108): Sequential(drop, Lin, [relu, drop, Lin]*) -> heads."""
    idx = 1
    x = _lin(_drop(x, dropout, training), sd, f"{prefix}.{idx}")
    for _ in range(ffn_depth - 1):
        idx += 3
        x = _lin(_drop(torch.relu(x), dropout, training), sd, f"{prefix}.{idx}")
    out = x.squeeze(-1)                                            # base_model.py:60
    if head == "evidential_ranking":                               # base_model.py:91-98
        score, u = torch.split(out, out.shape[1] // 2, dim=1)
        return torch.stack((score, F.softplus(u) + 1e-6), dim=2).view(out.size())
    if head in ("gauss_regression_with_softplus", "gaussian_with_softplus"):   # base_model.py:71-83
        mu, lv = torch.split(out, out.shape[1] // 2, dim=1)
        return torch.stack((mu, F.softplus(lv)), dim=2).view(out.size())
    if head in ("listnet_with_softplus",):                         # base_model.py:99-100
        return F.softplus(out)
    if head in ("listnet_with_uncertainty", "evidential"):         # base_model.py:101-104
        return F.softplus(out) + 1
    if head == "listnetdis_lognorm_with_softplus":                 # base_model.py:83-90
        mu, lv = torch.split(out, out.shape[1] // 2, dim=1)
        return torch.stack((F.softplus(mu) + 1e-6, F.softplus(lv) + 1e-6), dim=2).view(out.size())
    if head == "evidential_with_softplus":                         # base_model.py:61-70
        mu, ll, la, lb = torch.split(out, out.shape[1] // 4, dim=1)
        return torch.stack((mu, F.softplus(ll) + 1e-6, F.softplus(la) + 1e-6 + 1, F.softplus(lb) + 1e-6), dim=2).view(out.size())
    return out                                                     # base_model.py:105-106


def model_forward(sd, r_g, p_g, add_features, *, mpnn_depth=3, mpnn_diff_depth=3, ffn_depth=3, head="with_softplus",
                  dropout: float = 0.0, training: bool = False, trace: Optional[dict] = None) -> torch.Tensor:
    """``ReactionModel.forward`` (base_model.py:150-171): shared encoder on both graphs,
    atom-wise difference, diff encoder on the product graph, FFN."""
    r = mpn_forward(sd, r_g, mpnn_depth, dropout, training, trace=None)
    p = mpn_forward(sd, p_g, mpnn_depth, dropout, training, trace=trace)
    if trace is not None:
        trace["r_hiddens"], trace["p_hiddens"] = r, p
    vec = mpndiff_forward(sd, p - r, p_g, mpnn_diff_depth, add_features, dropout, training, trace=trace)
    if trace is not None:
        trace["readout"] = vec
    return ffn_forward(sd, vec, head, ffn_depth, dropout, training)


# ----------------------------------------------------------------------------------------
# losses  (train/loss.py, train/train_pairwise.py)
# ----------------------------------------------------------------------------------------
def listmle_loss(score: torch.Tensor, scope: Sequence[int], targets: torch.Tensor) -> torch.Tensor:
    """``MLEloss`` (loss.py:64-99) with the forward of ``LogCumsumExp`` (loss.py:28-34).
    Written with autograd-native ops: its gradient equals the reference's hand-written
    backward (loss.py:57-61) because the upstream gradient of ``mean`` is uniform.
    Returns shape [1] like the reference."""
    total = score.new_zeros(1)
    for s, t in zip(score.split(list(scope)), targets.split(list(scope))):
        order = torch.argsort(t, descending=True)
        x = s[order]
        m = x.max()
        lcse = torch.log(torch.flip(torch.cumsum(torch.flip(torch.exp(x - m), [0]), 0), [0])) + m
        total = total + torch.mean(lcse - x)
    return total / len(scope)


def listnet_loss(score, scope, targets) -> torch.Tensor:
    """``ListnetLoss`` (loss.py:317-352): mean over ALL items of -softmax(t) * log softmax(s)."""
    parts = []
    for s, t in zip(score.split(list(scope)), targets.split(list(scope))):
        parts.append(-F.softmax(t, dim=0) * torch.log(F.softmax(s, dim=0)))
    return torch.mean(torch.cat(parts))


def evidential_ranking_loss(out2, scope, targets) -> torch.Tensor:
    """``evidential_ranking`` live branch (loss.py:526-554); returns shape [1]."""
    total = out2.new_zeros(1)
    for o, t in zip(out2.split(list(scope)), targets.split(list(scope))):
        mean, var = o[:, 0], o[:, 1]
        q = F.softmax(mean, dim=0)
        p = F.softmax(t, dim=0)
        unc = 0.5 * (torch.log(p) - torch.log(q)) ** 2 / var + 0.5 * torch.log(2 * 3.141592653 * var)
        total = total + torch.mean(-torch.log(p) + unc + torch.abs(mean - t))
    return total / len(scope)


def gauss_loss(mu, var, targets) -> torch.Tensor:
    """``GaussDisLoss`` (loss.py:144-162)."""
    pi = torch.tensor([np.pi], dtype=torch.float32, device=var.device)         # ``torch.Tensor([np.pi])``: the constant term is evaluated in fp32 even in an fp64 run
    return torch.mean(0.5 * torch.log(2 * pi) + 0.5 * torch.log(var) + (mu - targets) ** 2 / (2 * var))


def mse_loss(out, targets) -> torch.Tensor:
    """``nn.MSELoss`` default branch (train_listwise.py:166-167, 282-285)."""
    return F.mse_loss(out, targets)


def ranknet_group_cost(y_pred: torch.Tensor, targets: np.ndarray, sigma: float = 1.0):
    """One group of ``factorized_training_loop``/'sum_session' (train_pairwise.py:98-122).
    Returns (sum of C over ordered pairs, num_pairs) or (None, 0) for a skipped group."""
    Y = np.asarray(targets).reshape(-1, 1)
    rel = Y - Y.T
    pos = (rel > 0).astype(np.float32)
    npos = pos.sum()
    if npos == 0:
        return None, 0.0
    neg = (rel < 0).astype(np.float32)
    pos_t = torch.from_numpy(pos).to(y_pred.dtype)
    neg_t = torch.from_numpy(neg).to(y_pred.dtype)
    if y_pred.dim() > 1:
        y_pred = y_pred[:, 0]
    y = y_pred.unsqueeze(1)
    c_pos = torch.log(1 + torch.exp(-sigma * (y - y.t())))
    c_neg = torch.log(1 + torch.exp(sigma * (y - y.t())))
    return torch.sum(pos_t * c_pos + neg_t * c_neg), 2.0 * float(npos)


def ranknet_group_lambda(y_pred: torch.Tensor, targets: np.ndarray, sigma: float = 1.0):
    """One group of ``factorized_training_loop``/'accelerate_grad' (train_pairwise.py:123-137): the cost evaluated without a graph and
    the hand-written gradient ``back`` [n, 1] the loop later feeds to ``y_pred.backward(back / pairs)`` (line 152).
    Returns (cost, back, num_pairs) or (None, None, 0) for a skipped group."""
    Y = np.asarray(targets).reshape(-1, 1)
    rel = Y - Y.T
    pos = (rel > 0).astype(np.float32)
    npos = pos.sum()
    if npos == 0:
        return None, None, 0.0
    neg = (rel < 0).astype(np.float32)
    if y_pred.dim() > 1:
        y_pred = y_pred[:, 0]
    y = y_pred.detach().unsqueeze(1)
    pos_t = torch.from_numpy(pos).to(y.dtype)
    neg_t = torch.from_numpy(neg).to(y.dtype)
    l_pos = 1 + torch.exp(sigma * (y - y.t()))
    l_neg = 1 + torch.exp(-sigma * (y - y.t()))
    lam = -sigma * pos_t / l_pos + sigma * neg_t / l_neg
    cost = torch.sum(torch.log(l_neg) * pos_t + torch.log(l_pos) * neg_t)
    return cost, torch.sum(lam, dim=1, keepdim=True), 2.0 * float(npos)


def mledis_loss(mean, variance, scope, targets) -> torch.Tensor:
    """MLEDisLoss (loss.py:102-141), restated as written: per group, items sorted by target descending, the n x n matrix
    exp(s_i - s_j + (v_i + v_j) / 2) restricted to i >= j, column sums, mean of the logs; mean over groups, shape [1]."""
    total = torch.zeros(1, dtype=mean.dtype)
    o = 0
    for n in scope:
        t = targets[o:o + n]
        idx = torch.argsort(t, descending=True)
        s, v = mean[o:o + n][idx], variance[o:o + n][idx]
        x1 = -s.repeat(n, 1)
        x2 = -x1.t()
        y1 = v.repeat(n, 1)
        y2 = y1.t()
        total = total + torch.mean(-torch.log(1 / torch.sum(torch.tril(torch.exp(x1 + x2 + (y1 + y2) / 2)), 0)))
        o += n
    return total / len(scope)


def listnet_gauss_loss(mean, variance, scope, targets) -> torch.Tensor:
    """Listnet_For_Gauss (loss.py:233-272), restated as written; mean over groups, shape [1]."""
    total = torch.zeros(1, dtype=mean.dtype)
    o = 0
    for n in scope:
        m, v, t = mean[o:o + n], variance[o:o + n], targets[o:o + n]
        x1 = m.repeat(n, 1)
        x2 = x1.t()
        y1 = v.repeat(n, 1)
        y2 = y1.t()
        pred = 1 / torch.sum(torch.exp(x1 - x2 + (y1 + y2) / 2), dim=1)
        z1 = t.repeat(n, 1)
        z2 = z1.t()
        targ = 1 / torch.sum(torch.exp(z1 - z2), dim=1)
        total = total + (-torch.mean(targ * torch.log(pred)))
        o += n
    return total / len(scope)


def listnet_uq_loss(score, scope, targets, max_coeff, epoch, epochs) -> torch.Tensor:
    """Listnet_with_uq (loss.py:355-399), restated as written (KLDivLoss(reduction='batchmean') of 1-D tensors = sum / n)."""
    total = torch.zeros(1, dtype=score.dtype)
    o = 0
    for n in scope:
        item, t = score[o:o + n], targets[o:o + n]
        pred_p = item / torch.sum(item)
        targ_p = torch.softmax(t, dim=0)
        real_loss = torch.sum(targ_p * (torch.log(targ_p) - torch.log(pred_p))) / n
        consist = torch.log(targ_p / pred_p)
        penalty = torch.abs(consist * (item - torch.ones(n, dtype=score.dtype)))
        coef = max_coeff * (epoch / (epochs - 1)) ** 3
        total = total + torch.mean(real_loss + coef * penalty)
        o += n
    return total / len(scope)


def dirichlet_uq_loss(concentration, scope, targets, max_coeff, epoch, epochs) -> torch.Tensor:
    """Dirichlet_uq (loss.py:440-474), restated as written for the 1-D concentration a task_num = 1 model emits."""
    total = torch.zeros(1, dtype=concentration.dtype)
    o = 0
    for n in scope:
        alpha, t = concentration[o:o + n], targets[o:o + n]
        pred_p = alpha / torch.sum(alpha)
        targ_p = torch.softmax(t, dim=0)
        err = (pred_p - targ_p) ** 2
        var = pred_p * (1 - pred_p) / (torch.sum(alpha) + 1)
        residue = torch.log(targ_p / pred_p) * (alpha - 1)
        coef = max_coeff * (epoch / (epochs - 1)) ** 3
        total = total + torch.mean(err + var + coef * torch.abs(residue))
        o += n
    return total / len(scope)


def lognorm_loss(scores, std_scores, targets) -> torch.Tensor:
    """Lognorm (loss.py:165-184)."""
    pi = torch.tensor([np.pi], dtype=torch.float32, device=scores.device)     # loss.py:174, as in GaussDisLoss
    mse = 0.5 * torch.log(2 * pi) + 0.5 * torch.log(std_scores * (scores ** 2)) + torch.pow(torch.log(scores) - targets, 2) / (2 * std_scores)
    return torch.mean(mse)


def nig_loss(mu, v, alpha, beta, targets, lam=1.0, epsilon=1e-4) -> torch.Tensor:
    """evidential_loss_new (loss.py:402-437), restated as written.  NOTE the shapes of its call sites (train_listwise.py:229-260):
    the four parameters are [N,1] column slices, ``targets`` is [N]; ``targets - mu`` therefore broadcasts to [N,N] and the mean runs
    over every (reaction, target) pair of the batch.  Nothing here corrects that: pass the same shapes and the same thing happens."""
    two_b_lambda = 2 * beta * (1 + v)
    nll = (0.5 * torch.log(math.pi / v) - alpha * torch.log(two_b_lambda) + (alpha + 0.5) * torch.log(v * (targets - mu) ** 2 + two_b_lambda)
           + torch.lgamma(alpha) - torch.lgamma(alpha + 0.5))
    reg = torch.abs(targets - mu) * (2 * v + alpha)
    return torch.mean(nll + lam * (reg - epsilon))


def loss_for_task(task_type: str, output, scope, targets) -> torch.Tensor:
    """Loss dispatch of ``train()`` for the five north-star keys and the composite keys built from them (train_listwise.py:196-285)."""
    if task_type == "mle":
        return listmle_loss(output, scope, targets)
    if task_type == "listnet":
        return listnet_loss(output, scope, targets)
    if task_type == "evidential_ranking":
        return evidential_ranking_loss(output, scope, targets)
    if task_type == "gauss_regression":
        return gauss_loss(output[:, 0], output[:, 1], targets)
    # composite keys: sums of the terms above (train_listwise.py:204-210, 224-227, 263-266, 276-281)
    if task_type == "mle_gaussian":
        return listmle_loss(output[:, 0], scope, targets) + gauss_loss(output[:, 0], output[:, 1], targets)
    if task_type == "listnet_gauss":
        return listnet_loss(output[:, 0], scope, targets) + gauss_loss(output[:, 0], output[:, 1], targets)
    if task_type == "mle_regression":
        return mse_loss(output, targets) + listmle_loss(output, scope, targets)
    if task_type == "listnet_regression":
        return listnet_loss(output, scope, targets) + mse_loss(output, targets)
    if task_type == "regression_exploss":
        return torch.mean((torch.exp(targets) - torch.exp(output)) ** 2)
    if task_type == "mledis_gaussian":          # train_listwise.py:196-203
        return mledis_loss(output[:, 0], torch.exp(output[:, 1]), scope, targets) + gauss_loss(output[:, 0], output[:, 1], targets)
    if task_type == "listnet_uq":               # train_listwise.py:228-229 with the golden's (max_coeff, epoch, epochs) = (0.05, 3, 5)
        return listnet_uq_loss(output, scope, targets, 0.05, 3, 5)
    if task_type == "listnetdis_lognorm":       # train_listwise.py:215-219
        return lognorm_loss(output[:, 0], output[:, 1], targets)
    if task_type == "dirichlet_uq":             # train_listwise.py:269-270, same schedule point as listnet_uq
        return dirichlet_uq_loss(output, scope, targets, 0.05, 3, 5)
    if task_type in ("evidential", "mle_evidential", "mledis_evidential", "listnet_evidential"):   # train_listwise.py:229-260
        mu, lambdas, alphas, betas = (output[:, k::4] for k in range(4))         # [N,1] each
        if task_type == "evidential":
            return nig_loss(mu, lambdas, alphas, betas, targets, lam=0.1)
        if task_type == "mle_evidential":
            return listmle_loss(output[:, 0], scope, targets) + nig_loss(mu, lambdas, alphas, betas, targets, lam=0.2)
        variance = betas / (lambdas * (alphas - 1))
        rank = mledis_loss if task_type == "mledis_evidential" else listnet_gauss_loss
        return rank(mu[:, 0], variance[:, 0], scope, targets) + nig_loss(mu, lambdas, alphas, betas, targets, lam=0.1)
    if task_type == "listnetdis_gauss":         # train_listwise.py:211-215
        return listnet_gauss_loss(output[:, 0], output[:, 1], scope, targets) + gauss_loss(output[:, 0], output[:, 1], targets)
    return mse_loss(output, targets)


# ----------------------------------------------------------------------------------------
# batch planners  (data/load_reactions.py:235-273, 336-421)
# ----------------------------------------------------------------------------------------
def plan_batch_reactions(df, batch_size: int, seed: int, target_name: str, smiles_list=None,
                         add_features_name=None, shuffle_query=True, shuffle_batch=True):
    """``DataProcessor.generate_batch_reactions`` (load_reactions.py:336-421), pandas/sklearn
    calls kept so the RNG streams are the reference's.  Yields
    (smiles [n,2], targets [n,1], scope, add_features, row_index)."""
    from sklearn.utils import shuffle
    cols = smiles_list if smiles_list is not None else ["rsmi", "psmi"]
    reactants = df.rsmi.unique()
    if shuffle_query:
        reactants = shuffle(reactants, random_state=seed)
    room = batch_size
    chunks, scope = [], []

    def flush():
        part = chunks[0] if len(chunks) == 1 else __import__("pandas").concat(chunks)
        feats = None
        if add_features_name is not None:
            feats = part[add_features_name].values
            if feats.ndim == 1:
                feats = feats.reshape(-1, 1)
        return part[cols].values, part[target_name].values.reshape(-1, 1), list(scope), feats, part.index.values

    for reactant in reactants:
        grp = df[df.rsmi == reactant]
        n = grp.shape[0]
        if room - n >= 0:                                           # load_reactions.py:370-395
            if shuffle_batch:
                grp = grp.sample(frac=1, random_state=seed)
            chunks.append(grp)
            scope.append(n)
            room -= n
            if room < 2:
                yield flush()
                room, chunks, scope = batch_size, [], []
        else:                                                       # load_reactions.py:396-418
            grp = grp.sample(n=room, random_state=seed)
            chunks.append(grp)
            scope.append(room)
            yield flush()
            room, chunks, scope = batch_size, [], []
    if room < batch_size:                                           # load_reactions.py:420-421
        yield flush()


def plan_batch_per_query(df, seed: int, target_name: str, smiles_list=None, add_features_name=None,
                         shuffle_query=True, shuffle_batch=True):
    """``DataProcessor.generate_batch_per_query`` (load_reactions.py:235-273).  The quirk at
    line 264-267 (add_features taken from the TARGET column) is preserved."""
    from sklearn.utils import shuffle
    cols = smiles_list if smiles_list is not None else ["rsmi", "psmi"]
    reactants = df.rsmi.unique()
    if shuffle_query:
        reactants = shuffle(reactants, random_state=seed)
    for reactant in reactants:
        grp = df[df.rsmi == reactant]
        if shuffle_batch:
            grp = grp.sample(frac=1, random_state=seed)
        feats = None
        if add_features_name is not None:
            feats = grp[target_name].values
            if feats.ndim == 1:
                feats = feats.reshape(-1, 1)
        yield grp[cols].values, grp[target_name].values, feats, grp.index.values


# ----------------------------------------------------------------------------------------
# one training step, reference order (train/train_listwise.py:177-290) -- used for parity
# of updated weights and as the CPU-baseline leg of bench.py
# ----------------------------------------------------------------------------------------
def noam_lr(step: int, warmup_steps: int, total_steps: int, init_lr: float, max_lr: float, final_lr: float) -> float:
    """``NoamLR.step`` (train/utils.py:61-81) after ``step`` calls."""
    if step <= warmup_steps:
        return init_lr + step * (max_lr - init_lr) / warmup_steps
    if step <= total_steps:
        gamma = (final_lr / max_lr) ** (1 / (total_steps - warmup_steps))
        return max_lr * gamma ** (step - warmup_steps)
    return final_lr


def init_state_dict(hidden: int, task_num: int, add_features_dim: int, use_bias: bool = True, seed: int = 0,
                    mpnn_depth: int = 3, mpnn_diff_depth: int = 3, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """State dict with the reference's 20 key names and nn.Linear default
    init, created in module-construction order (base_model.py:128-148) so that the same
    torch seed reproduces ``build_model``'s weights."""
    torch.manual_seed(seed)
    import torch.nn as nn
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, i, o, bias):
        l = nn.Linear(i, o, bias=bias)
        sd[name + ".weight"] = l.weight.detach().to(dtype)
        if bias:
            sd[name + ".bias"] = l.bias.detach().to(dtype)

    sd["encoder.cached_zero_vector"] = torch.zeros(hidden, dtype=dtype)
    lin("encoder.W_i", FBOND_TOTAL, hidden, use_bias)
    if mpnn_depth > 1:
        lin("encoder.W_h", hidden, hidden, use_bias)
    lin("encoder.W_o", ATOM_FDIM + hidden, hidden, True)
    sd["diff_encoder.cached_zero_vector"] = torch.zeros(hidden, dtype=dtype)
    lin("diff_encoder.W_i", hidden, hidden, use_bias)
    if mpnn_diff_depth > 1:
        lin("diff_encoder.W_h", hidden + FBOND_TOTAL, hidden, use_bias)
    if mpnn_diff_depth > 0:
        lin("diff_encoder.W_o", 2 * hidden, hidden, True)
    lin("ffn.ffn.1", hidden + add_features_dim, hidden, use_bias)
    lin("ffn.ffn.4", hidden, hidden, use_bias)
    lin("ffn.ffn.7", hidden, task_num, use_bias)
    return sd


# ----------------------------------------------------------------------------------------
# validation metrics  (train/eval.py)
# ----------------------------------------------------------------------------------------
def _desc_order(x):
    """``sorted(enumerate(x), key=lambda t: t[1], reverse=True)`` indices (eval.py:500-503): stable, ties keep the earlier item."""
    return np.argsort(-np.asarray(x, dtype=np.float64), kind="stable")


def _ndcg(truth, pred):
    """compute_NDCG (eval.py:460-472)."""
    truth, pred = np.asarray(truth, dtype=np.float64), np.asarray(pred, dtype=np.float64)
    disc = np.log2(np.arange(2, len(truth) + 2))
    return float(np.sum(np.exp(pred) / disc) / np.sum(np.exp(truth) / disc))


def group_metrics(pred, targ, ratio: float = 0.25) -> np.ndarray:
    """The eight per-group numbers behind ``ranking_metrics`` (eval.py:497-553) and ``evaluate_top_scores`` (eval.py:112-165), host
    restatement with Python's ``round`` and stable sorts; column order of include/rr_sm100.h ``rr_rank_metrics``."""
    pred, targ = np.asarray(pred, dtype=np.float64), np.asarray(targ, dtype=np.float64)
    n = len(targ)
    p_idx, t_idx = _desc_order(pred), _desc_order(targ)
    k = max(1, round(n * ratio))
    p_k, t_k = p_idx[:k].tolist(), set(t_idx[:k].tolist())
    t_sorted, by_pred = targ[t_idx], targ[p_idx]
    return np.asarray([
        float(p_idx[0] == t_idx[0]),
        sum(int(i in t_k) for i in p_k) / k,
        float(p_idx[0] in t_k),
        float(int(np.argmax(targ)) in set(p_k)),
        _ndcg(t_sorted[:1], by_pred[:1]),
        float(np.sum(np.exp(by_pred[:2])) / np.sum(np.exp(t_sorted[:2]))),     # eval.py:544 wraps both items in one list position
        _ndcg(t_sorted[:k], by_pred[:k]),
        _ndcg(t_sorted, by_pred)], dtype=np.float64)
