"""Parameter containers of the two encoders, with the reference's module/parameter names
(models/mpn.py:11-59, 127-168) so ``state_dict()`` keys, shapes and default initialisation
order are identical.  The arithmetic of ``MPN.forward`` / ``MPNDiff.forward``
(mpn.py:61-124, 170-240) runs inside ``ReactionModel.forward`` as one fused sequence of
sm_100a kernels (csrc/rr_model.cu); the sub-modules are not callable on their own.
"""
import torch
import torch.nn as nn


class MPN(nn.Module):
    """Bond-message D-MPNN (mpn.py:11-59): W_i [h, bond_fdim], W_h [h, h], W_o [h, atom_fdim + h]."""

    def __init__(self, bond_fdim: int, atom_fdim: int, MPN_hidden_size: int, MPN_bias: bool = True, MPN_depth: int = 6,
                 MPN_dropout: float = 0.2, return_atom_hiddens: bool = False):
        super().__init__()
        self.return_atom_hiddens = return_atom_hiddens
        self.bond_fdim, self.atom_fdim = bond_fdim, atom_fdim
        self.hidden_size, self.bias, self.depth, self.dropout = MPN_hidden_size, MPN_bias, MPN_depth, MPN_dropout
        self.layers_per_message = 1
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.act_func = nn.ReLU()
        self.cached_zero_vector = nn.Parameter(torch.zeros(self.hidden_size), requires_grad=False)
        self.W_i = nn.Linear(self.bond_fdim, self.hidden_size, bias=self.bias)
        if self.depth > 1:
            self.W_h = nn.Linear(self.hidden_size, self.hidden_size, bias=self.bias)
        self.W_o = nn.Linear(self.atom_fdim + self.hidden_size, self.hidden_size)   # always biased (mpn.py:59)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("MPN runs fused inside ReactionModel.forward (reactranker_b200/csrc/rr_model.cu)")


class MPNDiff(nn.Module):
    """Atom-message encoder of the difference features (mpn.py:127-168):
    W_i [h, h], W_h [h, h + bond_fdim], W_o [h, 2h]."""

    def __init__(self, atom_fdim: int, bond_fdim: int, MPNDiff_hidden_size: int, MPNDiff_bias: bool = True,
                 MPNDiff_depth: int = 3, MPNDiff_dropout: float = 0.2):
        super().__init__()
        self.atom_fdim, self.bond_fdim = atom_fdim, bond_fdim
        self.hidden_size, self.bias, self.depth, self.dropout = MPNDiff_hidden_size, MPNDiff_bias, MPNDiff_depth, MPNDiff_dropout
        self.layers_per_message = 1
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.act_func = nn.ReLU()
        self.cached_zero_vector = nn.Parameter(torch.zeros(self.hidden_size), requires_grad=False)
        self.W_i = nn.Linear(self.atom_fdim, self.hidden_size, bias=self.bias)
        if self.depth > 1:
            self.W_h = nn.Linear(self.hidden_size + self.bond_fdim, self.hidden_size, bias=self.bias)
        if self.depth > 0:
            self.W_o = nn.Linear(self.atom_fdim + self.hidden_size, self.hidden_size)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("MPNDiff runs fused inside ReactionModel.forward (reactranker_b200/csrc/rr_model.cu)")
