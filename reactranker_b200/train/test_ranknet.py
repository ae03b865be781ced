"""RankNet ``test(...)`` with the reference's signature (train/test_ranknet.py:10-83).  Not a unit test."""
from logging import Logger

import torch.nn as nn
from pandas import DataFrame

from .. import _lib
from ..data.load_reactions import DataProcessor
from ..utils import load_checkpoint
from .eval import evaluate_top_scores


def test(model: nn.Module, test_data: DataFrame, path_checkpoints: str, smiles2graph_dic, batch_size: int, gpu: int, logger: Logger = None,
         smiles_list: list = None, target_name='ea', train_strategy='baseline', add_features_name=None):
    gpu = _lib.require_device(gpu)
    state = load_checkpoint(path_checkpoints)
    means, stds = state['data_scaler']['means'], state['data_scaler']['stds']
    print('means is: ', means)
    sign = 1.0 if target_name == 'lgk' else -1.0
    test_data['std' + target_name] = test_data[target_name].map(lambda x: sign * (x - means) / stds)
    model.load_state_dict(state['state_dict'])
    model = model.cuda(gpu).eval()
    score, pred_in_targ, top1_in_pred = evaluate_top_scores(model, gpu, DataProcessor(test_data), smiles2graph_dic, ratio=0.25, batch_size=batch_size,
                                                             smiles_list=smiles_list, target_name='std' + target_name, add_features_name=add_features_name)
    print('   Note：For test set average score is:   ', score)
    if logger is not None:
        logger.info('\n  Note：For test set average score is: {:.4f}\n'.format(score))
    return score, pred_in_targ, top1_in_pred
