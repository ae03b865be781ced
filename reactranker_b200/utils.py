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


def save_train_state(path, model, optimizer, scheduler, epoch, means=None, stds=None, best=None):
    """A checkpoint one can RESUME from (the reference saves weights only and restarts from scratch).
    A superset of ``save_checkpoint``'s dict, so the reference's loaders (test_listwise.py:27-38) read it unchanged; the extra keys carry
    Adam's moments and step, the NoamLR position, the last finished epoch, the best validation scores so far and torch's host RNG state
    (dropout seeds are drawn from it).  Batches are planned from ``seed=epoch`` (train_listwise.py:179-182), so a resumed run sees the
    same batches as an uninterrupted one.  Written to a temporary name and renamed: a crash mid-write cannot destroy the last state."""
    import os
    scaler = {'means': float(means), 'stds': float(stds)} if means is not None and stds is not None else None
    state = {'state_dict': model.state_dict(), 'data_scaler': scaler, 'optimizer': optimizer.state_dict(),
             'scheduler': {k: v for k, v in scheduler.state_dict().items() if not k.startswith('_')} if scheduler is not None else None,
             'epoch': int(epoch), 'best': best, 'rng_state': torch.get_rng_state()}
    tmp = str(path) + '.tmp'
    torch.save(state, tmp)
    os.replace(tmp, path)


def load_train_state(path, model, optimizer, scheduler=None):
    """Inverse of ``save_train_state``: restores model / optimizer / scheduler in place and returns ``(next_epoch, best, data_scaler)``."""
    state = load_checkpoint(path)
    if 'optimizer' not in state:
        raise _lib.RRError(f"{path} is a weights-only checkpoint (save_checkpoint); resuming needs one written by save_train_state")
    model.load_state_dict(state['state_dict'])
    optimizer.load_state_dict(state['optimizer'])
    if scheduler is not None and state.get('scheduler') is not None:
        for k, v in state['scheduler'].items():
            setattr(scheduler, k, v)
        optimizer.param_groups[0]['lr'] = scheduler.lr
    torch.set_rng_state(state['rng_state'])
    return state['epoch'] + 1, state.get('best'), state['data_scaler']


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
