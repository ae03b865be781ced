"""One launch of selected message-passing kernels (for ncu captures)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reactranker_b200 import _lib, synthetic
from reactranker_b200.features.featurization import BatchMolGraph
L = _lib.lib()
dev = torch.device("cuda:0")
ds = synthetic.make_dataset(1000, [50] * 82)
b = BatchMolGraph([ds.mols[t] for t in ds.psmi])
dg = b.to_device(dev)
A, B, hp = b.n_atoms, b.n_bonds, 304
S = _lib.stream_ptr
g = ctypes.byref(dg.c)
mB, oB = (torch.randn(B, hp, device=dev) for _ in range(2))
mA = torch.randn(A, hp, device=dev)
for _ in range(2):
    _lib.check(L.rr_bond_message_fwd(g, mB.data_ptr(), oB.data_ptr(), hp, 1, S()))
    _lib.check(L.rr_neighbor_sum_bwd(g, 0, mA.data_ptr(), oB.data_ptr(), hp, S()))
torch.cuda.synchronize()
