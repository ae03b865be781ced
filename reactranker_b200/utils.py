"""``save_checkpoint`` and ``index_select_ND`` with the reference's signatures (utils.py:152-193)."""
import ctypes

import torch

from . import _lib


def save_checkpoint(path, model, means=None, stds=None):
    """Same dict layout as the reference (utils.py:152-173): ``{'state_dict', 'data_scaler': {'means','stds'}}``.
    ``means``/``stds`` are stored as python floats so that ``torch.load(weights_only=True)`` (the default since
    torch 2.6) can read the file; numerically identical to the reference's numpy scalars."""
    scaler = {'means': float(means), 'stds': float(stds)} if means is not None and stds is not None else None
    torch.save({'state_dict': model.state_dict(), 'data_scaler': scaler}, path)


def load_checkpoint(path):
    """Reads checkpoints written by this package or by the reference (numpy-scalar means/stds)."""
    try:
        return torch.load(path, map_location=lambda storage, loc: storage)
    except Exception:
        return torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)


def index_select_ND(source: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``source[index]`` with the reference's shape contract (utils.py:176-193): [R, h], [A, W] -> [A, W, h].
    The hot path never materialises this tensor (the gather kernels sum on the fly); kept for API parity."""
    if not source.is_cuda:
        raise _lib.RRError("index_select_ND: reactranker_b200 tensors live on the GPU (no CPU fallback)")
    return source.index_select(0, index.reshape(-1).to(source.device)).view(index.shape + source.shape[1:])
