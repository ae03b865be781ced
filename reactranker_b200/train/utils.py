"""``NoamLR`` / ``build_optimizer`` / ``build_lr_scheduler`` with the reference's signatures and
schedule (train/utils.py:7-133).  Adam itself is ``torch.optim.Adam`` (fused CUDA kernel)."""
from typing import List

import torch.nn as nn
from torch.optim import Adam, Optimizer
from torch.optim.lr_scheduler import _LRScheduler


class NoamLR(_LRScheduler):
    """Linear warm-up ``init_lr -> max_lr`` over ``warmup_epochs * steps_per_epoch`` steps, then
    exponential decay to ``final_lr`` at ``total_epochs * steps_per_epoch`` (train/utils.py:7-81).
    As in the reference the base-class constructor already performs one ``step()``."""

    def __init__(self, optimizer: Optimizer, warmup_epochs: int, total_epochs: int, steps_per_epoch: int,
                 init_lr: float, max_lr: float, final_lr: float):
        self.optimizer = optimizer
        self.warmup_epochs, self.total_epochs, self.steps_per_epoch = warmup_epochs, total_epochs, steps_per_epoch
        self.init_lr, self.max_lr, self.final_lr = init_lr, max_lr, final_lr
        self.current_step = 0
        self.lr = init_lr
        self.warmup_steps = int(self.warmup_epochs * self.steps_per_epoch)
        self.total_steps = self.total_epochs * self.steps_per_epoch
        self.linear_increment = (self.max_lr - self.init_lr) / self.warmup_steps
        self.exponential_gamma = (self.final_lr / self.max_lr) ** (1 / (self.total_steps - self.warmup_steps))
        super().__init__(optimizer)

    def get_lr(self) -> List[float]:
        return [self.lr]

    def step(self, current_step: int = None):
        if current_step is not None:
            self.current_step = current_step
        else:
            self.current_step += 1
        if self.current_step <= self.warmup_steps:
            self.lr = self.init_lr + self.current_step * self.linear_increment
        elif self.current_step <= self.total_steps:
            self.lr = self.max_lr * (self.exponential_gamma ** (self.current_step - self.warmup_steps))
        else:
            self.lr = self.final_lr
        self.optimizer.param_groups[0]['lr'] = self.lr


def param_count(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def build_optimizer(model: nn.Module, freeze=False) -> Optimizer:
    """Adam(lr=1e-4, weight_decay=0) on all parameters (train/utils.py:93-106)."""
    ps = model.parameters() if not freeze else filter(lambda p: p.requires_grad, model.parameters())
    ps = list(ps)
    fused = len(ps) > 0 and all(p.is_cuda for p in ps)
    params = [{'params': ps, 'lr': 0.0001, 'weight_decay': 0}]
    return Adam(params, fused=True) if fused else Adam(params)


def build_lr_scheduler(optimizer: Optimizer, warmup_epochs: int, total_epochs: int, train_data_size: int, batch_size: int,
                       init_lr: float, max_lr: float, final_lr: float):
    return NoamLR(optimizer=optimizer, warmup_epochs=warmup_epochs, total_epochs=total_epochs,
                  steps_per_epoch=train_data_size // batch_size, init_lr=init_lr, max_lr=max_lr, final_lr=final_lr)
