"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from reactranker_b200 import synthetic


def star_dict(arr):
    return {int(k): int(v) for k, v in np.asarray(arr).reshape(-1, 2)} or None


def dataset_from_golden(g, name):
    hidden, seed, depth, ddepth = (int(x) for x in g[name + ".meta"])
    sizes = [int(x) for x in g[name + ".sizes"]]
    ds = synthetic.make_dataset(seed, sizes, star_leaves_in_group=star_dict(g[name + ".star"]))
    return ds, sizes, hidden, depth, ddepth


def sd_from_golden(g, prefix, dtype=torch.float32):
    pre = prefix + "."
    return {k[len(pre):]: torch.tensor(g[k]).to(dtype) for k in g.files if k.startswith(pre)}


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def grads_close(got: dict, want: dict, rtol: float):
    """Per-tensor max-abs error <= rtol * max|want_k| + 1e-2 * rtol * (largest gradient entry
    of the whole model).  The second term absorbs tensors whose true gradient is zero (e.g. the
    last bias under a shift-invariant ranking loss).  Returns a list of offending keys."""
    gscale = max(float(np.abs(v).max()) for v in want.values())
    bad = []
    for k, w in want.items():
        a = np.asarray(got[k], np.float64)
        w = np.asarray(w, np.float64)
        err = np.abs(a - w).max()
        if not err <= rtol * np.abs(w).max() + 1e-2 * rtol * gscale:
            bad.append((k, float(err), float(np.abs(w).max())))
    return bad
