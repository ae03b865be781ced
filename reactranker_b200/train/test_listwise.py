"""Checkpoint-load + evaluation ``test(...)`` with the reference's signature (train/test_listwise.py:10-86).  Not a unit test."""
from logging import Logger
from typing import Union

from .. import _lib
from ..data.load_reactions import DataProcessor
from ..utils import load_checkpoint
from .eval import calculate_ndcg, evaluate_top_scores


def test(model, test_data, path_checkpoints, batch_size, smiles2graph_dic, gpu: Union[int, str], logger: Logger = None, smiles_list=None,
         target_name: str = 'ea', cal_ngcd=False, is_order=True, return_order=True, show_info=None, add_features_name=None, task_type=None):
    gpu = _lib.require_device(gpu)
    print('==========================================')
    print('  Now, the test section is beginning!!!   ')
    print('==========================================')
    if logger is not None:
        logger.info('The path of checkpoints is:\n')
        logger.info(path_checkpoints)
        logger.info('the length of test data is: {}'.format(test_data.shape[0]))
    state = load_checkpoint(path_checkpoints)
    if state['data_scaler'] is not None and state['data_scaler']['means'] is not None:
        test_data['std' + target_name] = test_data[target_name] if target_name == 'lgk' else -test_data[target_name]
    model.load_state_dict(state['state_dict'])
    model = model.cuda(gpu)
    model.train() if task_type == 'MC_dropout' else model.eval()
    proc = DataProcessor(test_data)
    average_score, average_pred_in_targ, average_top1_in_pred = evaluate_top_scores(
        model, gpu=gpu, data_processor=proc, smiles2graph_dic=smiles2graph_dic, batch_size=batch_size, ratio=0.25, smiles_list=smiles_list,
        target_name='std' + target_name, show_info=show_info, add_features_name=add_features_name)
    print('   Note：For test set average score is:   ', average_score)
    print('   Note：For test set 0.25 average pred in targ is:{}'.format(average_pred_in_targ))
    print('   Note：For average target top1 in pred 0.25 is:{}'.format(average_top1_in_pred))
    if logger is not None:
        logger.info('\n  Note：For test set average score is: {:.4f}\n'.format(average_score))
    if cal_ngcd is True:                          # test_listwise.py:58-63
        scaler = state['data_scaler'] or {}
        ndcg, kl_div, order, smiles_and_index = calculate_ndcg(
            model, gpu=gpu, data_processor=proc, smiles2graph_dic=smiles2graph_dic, batch_size=batch_size, NDCG_cut=0.25, smiles_list=smiles_list,
            target_name='std' + target_name, is_order=is_order, means=scaler.get('means'), stds=scaler.get('stds'), add_features_name=add_features_name)
        print('   Note：For test set NDCG{} is: {}'.format(0.25, ndcg))
        print('   Note：For test set KL divergence is: {}'.format(kl_div))
        if logger is not None:
            logger.info('\n Note：For test set NDCG{} is: {}'.format(0.25, ndcg))
            logger.info('\n Note：For test set KL divergence is: {}'.format(kl_div))
        if return_order is True:
            return average_score, average_pred_in_targ, average_top1_in_pred, order, smiles_and_index
    return average_score, average_pred_in_targ, average_top1_in_pred
