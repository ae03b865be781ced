"""RankNet epoch driver ``run_train`` with the reference's signature (train/run_train_pairwise.py:18-140): target z-scoring,
'sum_session' training, ``evaluate_top_scores`` validation, best-metric checkpointing."""
from __future__ import annotations

import copy
import os
from logging import Logger

import torch
import torch.nn as nn
from pandas import DataFrame
from torch.optim.lr_scheduler import _LRScheduler
from tqdm import trange

from .. import _lib, parallel
from ..data.load_reactions import DataProcessor
from ..utils import load_train_state, save_checkpoint, save_train_state
from .eval import evaluate_top_scores
from .train_pairwise import factorized_training_loop

try:
    from torch.utils.tensorboard import SummaryWriter
except Exception:  # pragma: no cover
    SummaryWriter = None


def run_train(model: nn.Module, scheduler: _LRScheduler, train_data_ini: DataFrame, val_data_ini: DataFrame, path_checkpoints: str, optimizer,
              epochs: int, smiles2graph_dic, batch_size: int, seed: int, gpu: int, train_strategy: str = 'baseline', task_type: str = 'baseline',
              writer=SummaryWriter, logger: Logger = None, smiles_list=None, target_name: str = 'ea', save_metric=None, add_features_name=None,
              resume_path=None):
    """``run_train`` of run_train_pairwise.py:20-110; ``resume_path`` as in train_listwise.train (None = the reference's behaviour)."""
    if train_strategy not in ('sum_session', 'accelerate_grad') or task_type != 'baseline':
        raise NotImplementedError("train_strategy 'sum_session' / 'accelerate_grad' with task_type='baseline' (main_ranknet.py's model) is built")
    gpu = _lib.require_device(gpu)
    torch.cuda.set_device(gpu)
    rank, world = parallel.init_from_env(device=torch.device("cuda", gpu))       # torchrun: data-parallel over the window's groups
    main = rank == 0
    if world > 1:
        parallel.broadcast_parameters(model, 0)
        torch.manual_seed(seed + 1000003 * rank)         # own dropout masks per rank
    train_data, val_data = copy.deepcopy(train_data_ini), copy.deepcopy(val_data_ini)
    mean, std = train_data[target_name].mean(), train_data[target_name].std(ddof=0)
    sign = 1.0 if target_name == 'lgk' else -1.0          # run_train_pairwise.py:40-45
    train_data['std' + target_name] = train_data[target_name].map(lambda x: sign * (x - mean) / std)
    val_data['std' + target_name] = val_data[target_name].map(lambda x: sign * (x - mean) / std)
    print('stds is: ', std)
    print('mean is: ', mean)
    train_proc, val_proc = DataProcessor(train_data), DataProcessor(val_data)
    score_old = [0, 0, 0] if save_metric == 'all' else float(0)
    first_epoch = 0
    if resume_path is not None and os.path.exists(resume_path):
        first_epoch, best, _ = load_train_state(resume_path, model, optimizer, scheduler)
        score_old = best if best is not None else score_old
        print('Note: resuming after epoch {} from {}'.format(first_epoch, resume_path))
    for epoch in trange(first_epoch, epochs, disable=not main):
        lr = optimizer.state_dict()['param_groups'][0]['lr']
        if main:
            print('learning rate: ', lr)
        if logger is not None and main:
            logger.info('learning rate is: {}'.format(lr))
        model.zero_grad()
        model.train()
        epoch_loss = factorized_training_loop(epoch, model, None, optimizer, scheduler, smiles2graph_dic, train_proc, batch_size=batch_size, sigma=1.0,
                                              training_algo=train_strategy, gpu=gpu, smiles_list=smiles_list, target_name='std' + target_name,
                                              add_features_name=add_features_name)
        model.eval()
        metrics = [0.0, 0.0, 0.0]
        if main:                                         # rank 0 validates and checkpoints; the others receive the three metrics
            metrics = list(evaluate_top_scores(model, gpu, val_proc, smiles2graph_dic, ratio=0.25, show_info=True, smiles_list=smiles_list,
                                               target_name='std' + target_name, add_features_name=add_features_name))
        if world > 1:
            t = torch.tensor([float(v) for v in metrics], dtype=torch.float64, device=torch.device("cuda", gpu))
            torch.distributed.broadcast(t, src=0)
            metrics = t.tolist()
        average_score, average_pred_in_targ, average_top1_in_pred = metrics
        if save_metric is None or save_metric == 'average_score':
            if average_score >= score_old:
                score_old = average_score
                if main:
                    save_checkpoint(path_checkpoints, model, mean, std)
                    print('Note: the checkpint file is updated')
        elif save_metric == 'all':
            for slot, val in enumerate((average_score, average_pred_in_targ, average_top1_in_pred)):
                if val >= score_old[slot]:
                    score_old[slot] = val
                    if main:
                        save_checkpoint(path_checkpoints[slot], model, mean, std)
                        print('Note: the checkpint file is updated')
        if logger is not None and main:
            logger.info('Epoch [{}/{}],train_loss,{:.4f}, average_score_top1,{:.4f}, average_pred_in_targ_top25%,{:.4f}'.format(
                epoch + 1, epochs, epoch_loss, average_score, average_top1_in_pred))
        if main:
            print('Epoch [{}/{}], average score: {:.4f}'.format(epoch + 1, epochs, average_score))
            print('Epoch [{}/{}], average_pred_in_targ_top25%: {:.4f}'.format(epoch + 1, epochs, average_pred_in_targ))
            print('Epoch [{}/{}], average_targtop1_in_predtop25%: {:.4f}'.format(epoch + 1, epochs, average_top1_in_pred))
            print('Epoch [{}/{}], train loss: {:.4f}'.format(epoch + 1, epochs, epoch_loss))
            print('Epoch [{}/{}], epoch mean loss, full precision = {!r}'.format(epoch + 1, epochs, epoch_loss))
        if resume_path is not None and main:
            save_train_state(resume_path, model, optimizer, scheduler, epoch, mean, std, best=score_old)
        if world > 1:
            torch.distributed.barrier()
